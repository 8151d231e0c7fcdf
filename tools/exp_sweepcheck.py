#!/usr/bin/env python3
"""Experiment: reproduce the N=1 sweep verification failure of bench.py and find out who is wrong (src / dst / exp)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
dist.init_process_group("gloo", rank=0, world_size=1)
from ucc_b200.dist import Communicator  # noqa: E402
from bench import pattern, expected_sum  # noqa: E402

comm = Communicator()
stream = torch.cuda.Stream()
dev = torch.device("cuda", 0)
SYNC_BEFORE_POST = len(sys.argv) > 1 and sys.argv[1] == "sync"
for nbytes, iters in ((1 << 30, 8), (64 << 20, 20), (64 << 20, 20), (256 << 20, 8), (1 << 30, 8), (64 << 20, 20), (16 << 20, 20), (64 << 20, 1)):
    cnt = nbytes // 4
    src = pattern(torch, cnt, 0, torch.float32, dev)
    dst = torch.empty(cnt, device=dev).zero_()
    if SYNC_BEFORE_POST:
        torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        hs = []
        for _ in range(5 + iters):
            r = comm.allreduce_init(src, dst)
            r.post_on_stream(stream, wait_posted=False)
            hs.append(r)
        for r in hs:
            r.wait_posted()
        for r in hs:
            r.wait(); r.finalize()
        torch.cuda.synchronize()
    exp = expected_sum(torch, cnt, 1, torch.float32, dev)
    src2 = pattern(torch, cnt, 0, torch.float32, dev)
    torch.cuda.synchronize()
    print(f"bytes {nbytes:>10} iters {iters:>2} sync {SYNC_BEFORE_POST}: src==exp {bool(torch.equal(src, exp))} src2==exp {bool(torch.equal(src2, exp))} dst==src {bool(torch.equal(dst, src))} "
          f"dst==exp {bool(torch.equal(dst, exp))} ptrs src {src.data_ptr():#x} dst {dst.data_ptr():#x} exp {exp.data_ptr():#x} kernel {comm.request_info_last()}", flush=True)
    hs = None
    del src, dst, exp, src2
comm.destroy()
