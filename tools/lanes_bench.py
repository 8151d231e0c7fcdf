#!/usr/bin/env python3
"""Do independent collectives of ONE team overlap?  K allreduces posted on K streams, UCC_TL_NVL_SLOTS = 1 (one lane: kernels of the team
run back to back) against SLOTS = K (each takes its own lane).  torchrun --nproc-per-node N tools/lanes_bench.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucc_b200.dist import Communicator, init_distributed  # noqa: E402

rank, N, _ = init_distributed("cpu:gloo,cuda:nccl")
dev = torch.device("cuda", torch.cuda.current_device())
K = int(os.environ.get("LANES_K", "4"))
ITERS = 10


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


out = []
for nbytes in (1 << 20, 16 << 20, 128 << 20):
    cnt = nbytes // 4
    for slots in (1, K):
        comm = Communicator(ctx_modify=[("tl/nvl", "SLOTS", str(slots))])
        streams = [torch.cuda.Stream() for _ in range(K)]
        src = [torch.ones(cnt, device=dev) for _ in range(K)]
        dst = [torch.empty(cnt, device=dev) for _ in range(K)]
        torch.cuda.synchronize()
        main = torch.cuda.Stream()

        def round_():
            reqs = []
            for k in range(K):           # same post order on every rank: collective k takes lane k % slots
                r = comm.allreduce_init(src[k], dst[k])
                r.post_on_stream(streams[k])
                reqs.append(r)
            return reqs

        for _ in range(3):
            for r in round_():
                r.wait(); r.finalize()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(main):
            e0.record(main)
        for s in streams:
            s.wait_stream(main)
        allreqs = []
        for _ in range(ITERS):
            allreqs += round_()
        for s in streams:
            main.wait_stream(s)
        with torch.cuda.stream(main):
            e1.record(main)
        for r in allreqs:
            r.wait(); r.finalize()
        torch.cuda.synchronize()
        us = maxr(e0.elapsed_time(e1) * 1e3 / ITERS)
        ok = all(bool((d == N).all()) for d in dst)
        out.append({"bytes": nbytes, "slots": slots, "K": K, "us_per_round": round(us, 1), "ok": ok, "kernel": comm.request_info_last()})
        comm.destroy()
if rank == 0:
    for row in out:
        print(json.dumps(row), flush=True)
dist.destroy_process_group()
