#!/usr/bin/env python3
"""Allreduce algorithm / SM-budget matrix in ONE torchrun launch: several communicators with different tl/nvl settings, each timed
at a few sizes on cudaMalloc (or symmetric) float32 buffers, device time = CUDA events, max over ranks, whole-vector verification."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucc_b200.dist import Communicator, init_distributed  # noqa: E402
from bench import pattern, expected_sum  # noqa: E402

rank, N, _ = init_distributed("cpu:gloo,cuda:nccl")
dev = torch.device("cuda", torch.cuda.current_device())
stream = torch.cuda.Stream()
SIZES = [int(x) for x in os.environ.get("MATRIX_SIZES", f"{64 << 20},{256 << 20},{1 << 30}").split(",")]
ITERS = int(os.environ.get("MATRIX_ITERS", "10"))


def maxr(x):
    t = torch.tensor([x], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def variant(name, mods, symm=None):
    try:
        comm = Communicator(ctx_modify=[("tl/nvl", k, v) for k, v in mods], symm_size=symm)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"variant": name, "error": str(e)}), flush=True)
        return
    row = {"variant": name, "n": N}
    use_symm = symm and comm.symm_region() is not None
    for S in SIZES:
        cnt = S // 4
        if use_symm:
            comm.symm_reset()
            src, dst = comm.symm_empty(cnt), comm.symm_empty(cnt)
            pattern(torch, cnt, rank, torch.float32, dev, out=src)
        else:
            src, dst = pattern(torch, cnt, rank, torch.float32, dev), torch.empty(cnt, device=dev)
        dst.zero_()
        torch.cuda.synchronize()
        try:
            with torch.cuda.stream(stream):
                reqs = [comm.allreduce_init(src, dst) for _ in range(3 + ITERS)]
                for r in reqs[:3]:
                    r.post_on_stream(stream)
                for r in reqs[:3]:
                    r.wait()
                dist.barrier(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for r in reqs[3:]:
                    r.post_on_stream(stream, wait_posted=False)
                for r in reqs[3:]:
                    r.wait_posted()
                e1.record(stream)
                for r in reqs[3:]:
                    r.wait()
                torch.cuda.synchronize()
                for r in reqs:
                    r.finalize()
            us = maxr(e0.elapsed_time(e1) * 1e3 / ITERS)
            ok = bool(torch.equal(dst, expected_sum(torch, cnt, N, torch.float32, dev)))
            row[str(S)] = {"us": round(us, 1), "busbw": round(S / us / 1e3 * 2 * (N - 1) / N, 1), "ok": ok, "kernel": comm.request_info_last()}
        except Exception as e:  # noqa: BLE001
            row[str(S)] = {"error": str(e)}
        del src, dst
    comm.destroy()
    if rank == 0:
        print(json.dumps(row), flush=True)


FORCE = lambda alg: [("TUNE", f"allreduce:cuda:inf:@{alg}"), ("ALLREDUCE_ONESHOT_THRESH", "0")]  # noqa: E731
which = os.environ.get("MATRIX_VARIANTS", "default,twoshot,nvls,nvls_pipe,nvls_pipe384,symm,symm_nb32,symm_nb64,symm_nb128,default_nb64,default_nb128,twoshot_nb64,twoshot_nb128").split(",")
V = {
    "default": ([], None), "twoshot": (FORCE("twoshot"), None), "nvls": (FORCE("nvls"), None), "nvls_pipe": (FORCE("nvls_pipe"), None),
    "nvls_pipe384": (FORCE("nvls_pipe") + [("SYMMETRIC_SIZE", "384M")], None), "symm": ([], "3G"),
    "symm_nb32": ([("MAX_BLOCKS", "32")], "3G"), "symm_nb64": ([("MAX_BLOCKS", "64")], "3G"), "symm_nb128": ([("MAX_BLOCKS", "128")], "3G"),
    "default_nb64": ([("MAX_BLOCKS", "64")], None), "default_nb128": ([("MAX_BLOCKS", "128")], None),
    "twoshot_nb64": (FORCE("twoshot") + [("MAX_BLOCKS", "64")], None), "twoshot_nb128": (FORCE("twoshot") + [("MAX_BLOCKS", "128")], None),
    "nvls_nb64": (FORCE("nvls") + [("MAX_BLOCKS", "64")], None),
    "nvls512": (FORCE("nvls") + [("SYMMETRIC_SIZE", "512M")], None), "nvls1g": (FORCE("nvls") + [("SYMMETRIC_SIZE", "1G")], None),
    "nvls_pipe1g": (FORCE("nvls_pipe") + [("SYMMETRIC_SIZE", "1G")], None),
}
for name in which:
    if name in V:
        variant(name, *V[name])
dist.destroy_process_group()
