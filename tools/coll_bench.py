#!/usr/bin/env python3
"""Bus bandwidth of allgather / reduce_scatter / alltoall / MoE-skewed alltoallv on CUDA buffers, ours vs NCCL (torch).
Formulas: reference tools/perf/ucc_pt_coll_{allgather,reduce_scatter,alltoall}.cc  (S/t * (N-1)/N, S = total bytes per rank)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucc_b200.dist import Communicator, init_distributed  # noqa: E402

rank, N, _ = init_distributed("cpu:gloo,cuda:nccl")
dev = torch.device("cuda", torch.cuda.current_device())
# COLL_VARIANTS=default,push,ce,...: several tl/nvl settings in one launch (each gets its own communicator); NCCL is timed once
VARIANTS = {"default": [], "push": [("TUNE", "allgather:cuda:inf:@push#alltoall:cuda:inf:@push#alltoallv:cuda:inf:@push")],
            "push_nobulk": [("TUNE", "allgather:cuda:inf:@push#alltoall:cuda:inf:@push#alltoallv:cuda:inf:@push"), ("BULK", "n")],
            "push_bulk32": [("TUNE", "allgather:cuda:inf:@push#alltoall:cuda:inf:@push#alltoallv:cuda:inf:@push"), ("BULK_CTAS", "32")],
            "push_bulk64": [("TUNE", "allgather:cuda:inf:@push#alltoall:cuda:inf:@push#alltoallv:cuda:inf:@push"), ("BULK_CTAS", "64")],
            "ce": [("TUNE", "allgather:cuda:inf:@ce#alltoall:cuda:inf:@ce#alltoallv:cuda:inf:@ce")],
            "nvls_ag": [("TUNE", "allgather:cuda:inf:@nvls")], "rs_nvls": [("TUNE", "reduce_scatter:cuda:inf:@nvls")],
            "rs_oneshot": [("TUNE", "reduce_scatter:cuda:0-1M:@oneshot")]}
which = [v for v in os.environ.get("COLL_VARIANTS", "default").split(",") if v in VARIANTS]
comms = {v: Communicator(ctx_modify=[("tl/nvl", k, val) for k, val in VARIANTS[v]]) for v in which}
comm = None
stream = torch.cuda.Stream()
ITERS, WARM = 10, 3


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def timed(fn_post, fn_wait=None):
    with torch.cuda.stream(stream):
        for _ in range(WARM):
            fn_post()
        if fn_wait:
            fn_wait()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(ITERS):
            fn_post()
        if fn_wait:
            fn_wait()
        e1.record(stream)
        torch.cuda.synchronize()
    return maxr(e0.elapsed_time(e1) * 1e3 / ITERS)


def ours(coll, src, dst, **kw):
    reqs = []
    kinds.setdefault(coll, {})

    def post():
        r = comm.coll_init(coll, src, dst, **kw)
        r.post_on_stream(stream)
        reqs.append(r)

    def wait():
        for r in reqs:
            r.wait(); r.finalize()
        reqs.clear()
    t = timed(post, wait)
    kinds[coll] = comm.request_info_last()
    return t


kinds = {}


out = []
nccl_cache = {}


def nccl(key, fn, total):
    if (key, total) not in nccl_cache:
        nccl_cache[(key, total)] = timed(fn)
    return nccl_cache[(key, total)]


for vname in which:
  comm = comms[vname]
  for total in (1 << 20, 16 << 20, 256 << 20):          # bytes of the full (gathered / scattered) vector per rank
      n = total // 4
      blk = n // N
      full = torch.ones(blk * N, device=dev)
      part = torch.ones(blk, device=dev)
      full2 = torch.empty(blk * N, device=dev)
      part2 = torch.empty(blk, device=dev)
      f = (N - 1) / N
      row = {"variant": vname, "bytes": blk * N * 4}
      t = ours("allgather", part, full2); row["allgather"] = (round(t, 1), round(blk * N * 4 / t / 1e3 * f, 1))
      t = nccl('ag', lambda: dist.all_gather_into_tensor(full2, part), total); row["allgather_nccl"] = (round(t, 1), round(blk * N * 4 / t / 1e3 * f, 1))
      t = ours("reduce_scatter", full, part2); row["reduce_scatter"] = (round(t, 1), round(blk * N * 4 / t / 1e3 * f, 1))
      t = nccl('rs', lambda: dist.reduce_scatter_tensor(part2, full), total); row["reduce_scatter_nccl"] = (round(t, 1), round(blk * N * 4 / t / 1e3 * f, 1))
      t = ours("alltoall", full, full2); row["alltoall"] = (round(t, 1), round(blk * N * 4 / t / 1e3 * f, 1))
      t = nccl('a2a', lambda: dist.all_to_all_single(full2, full), total); row["alltoall_nccl"] = (round(t, 1), round(blk * N * 4 / t / 1e3 * f, 1))
      # MoE-shaped alltoallv: rank r sends 2x the average to expert rank 0, the rest evenly (counts in elements)
      hot = min(2 * blk, n // 2)
      cold = (n - hot) // max(1, N - 1)
      sc = [hot] + [cold] * (N - 1)                      # everybody's token count per destination
      rc = [hot if rank == 0 else cold] * N              # what this rank receives from each source
      sd = [sum(sc[:i]) for i in range(N)]
      rd = [sum(rc[:i]) for i in range(N)]
      src = torch.ones(sum(sc), device=dev); dst = torch.empty(sum(rc), device=dev)
      t = ours("alltoallv", src, dst, src_counts=sc, src_displs=sd, dst_counts=rc, dst_displs=rd)
      row["alltoallv_moe"] = (round(t, 1), round(sum(sc) * 4 / t / 1e3 * f, 1))
      t = nccl('a2av', lambda: dist.all_to_all_single(dst, src, rc, sc), total); row["alltoallv_moe_nccl"] = (round(t, 1), round(sum(sc) * 4 / t / 1e3 * f, 1))
      row["kernels"] = dict(kinds)
      out.append(row)
      del full, part, full2, part2, src, dst
if rank == 0:
    print(json.dumps({"n_gpus": N, "unit": "(us, busbw GB/s)", "rows": out}))
for c in comms.values():
    c.destroy()
dist.destroy_process_group()
