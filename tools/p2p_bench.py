#!/usr/bin/env python3
"""send / recv (two-member active-set bcast) between rank 0 and rank 1 on CUDA buffers: ours (tl/nvl heap channel kernel) vs NCCL send/recv.
torchrun --nproc-per-node 2 tools/p2p_bench.py      unidirectional bandwidth and half round-trip latency, device-timed"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucc_b200.dist import Communicator, init_distributed  # noqa: E402

rank, N, _ = init_distributed("cpu:gloo,cuda:nccl")
dev = torch.device("cuda", torch.cuda.current_device())
comm = Communicator()
stream = torch.cuda.Stream()
ITERS = 20


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def p2p_req(buf, src, dst):
    return comm.coll_init("bcast", buf, None, root=src, active_set=(src, dst - src, 2), tag=1)


rows = []
for nbytes in (4096, 65536, 1 << 20, 16 << 20, 128 << 20):
    buf = torch.full((nbytes // 4,), float(rank + 1), device=dev)
    torch.cuda.synchronize()
    row = {"bytes": nbytes}
    if rank < 2:
        # ---- ours: rank 0 -> rank 1, ITERS messages back to back
        with torch.cuda.stream(stream):
            for _ in range(3):
                r = p2p_req(buf, 0, 1); r.post_on_stream(stream, wait_posted=False); r.wait(); r.finalize()
            torch.cuda.synchronize()
    dist.barrier()
    if rank < 2:
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reqs = [p2p_req(buf, 0, 1) for _ in range(ITERS)]
            e0.record(stream)
            for r in reqs:
                r.post_on_stream(stream, wait_posted=False)
            for r in reqs:
                r.wait(); r.finalize()
            # large messages enter the streams from progress (rendezvous: the push after the receiver published its buffer, the
            # receiver's wait kernel after the push was launched), so the closing event is recorded after the host-side waits
            e1.record(stream)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / ITERS
        ok = bool((buf == 1.0).all()) if rank == 1 else True
    else:
        us, ok = 0.0, True
    row["ours_us"] = round(maxr(us), 2); row["ours_GBps"] = round(nbytes / row["ours_us"] / 1e3, 1); row["ok"] = ok
    row["kernel"] = comm.request_info_last() if rank < 2 else ""
    # ---- NCCL send / recv
    buf.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
    if rank < 2:
        with torch.cuda.stream(stream):
            for _ in range(3):
                dist.send(buf, 1) if rank == 0 else dist.recv(buf, 0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(ITERS):
                dist.send(buf, 1) if rank == 0 else dist.recv(buf, 0)
            e1.record(stream)
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / ITERS
    else:
        us = 0.0
    row["nccl_us"] = round(maxr(us), 2); row["nccl_GBps"] = round(nbytes / max(row["nccl_us"], 1e-3) / 1e3, 1)
    rows.append(row)
    del buf
if rank == 0:
    for r in rows:
        print(json.dumps(r), flush=True)
comm.destroy()
dist.destroy_process_group()
