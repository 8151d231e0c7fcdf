#!/usr/bin/env python3
"""Experiment: device-to-device copy kernels of tl/nvl (thread copy vs TMA bulk copy) against cudaMemcpyAsync, with a full compare."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "ucc_b200", "lib", "libucc.so"), mode=C.RTLD_GLOBAL)
nvl = C.CDLL(os.path.join(ROOT, "ucc_b200", "lib", "ucc", "libucc_tl_nvl.so"))
nvl.nvl_launch_self_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
nvl.nvl_launch_self_copy_bulk.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
st = torch.cuda.Stream()


def timed(fn, iters=10):
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(iters):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for nbytes in (1 << 20, 64 << 20, 256 << 20, 1 << 30):
    n = nbytes // 4
    src = (torch.arange(n, device="cuda", dtype=torch.int64) % 13).float()
    torch.cuda.synchronize()
    row = {"bytes": nbytes}
    for name, fn in (("thread256x512", lambda d: nvl.nvl_launch_self_copy(d.data_ptr(), src.data_ptr(), nbytes, 256, 512, st.cuda_stream)),
                     ("thread296x512", lambda d: nvl.nvl_launch_self_copy(d.data_ptr(), src.data_ptr(), nbytes, 296, 512, st.cuda_stream)),
                     ("bulk148", lambda d: nvl.nvl_launch_self_copy_bulk(d.data_ptr(), src.data_ptr(), nbytes, 148, st.cuda_stream)),
                     ("bulk74", lambda d: nvl.nvl_launch_self_copy_bulk(d.data_ptr(), src.data_ptr(), nbytes, 74, st.cuda_stream)),
                     ("bulk32", lambda d: nvl.nvl_launch_self_copy_bulk(d.data_ptr(), src.data_ptr(), nbytes, 32, st.cuda_stream)),
                     ("bulk16", lambda d: nvl.nvl_launch_self_copy_bulk(d.data_ptr(), src.data_ptr(), nbytes, 16, st.cuda_stream)),
                     ("torch_copy", lambda d: d.copy_(src, non_blocking=True))):
        dst = torch.zeros(n, device="cuda")
        torch.cuda.synchronize()
        us = timed(lambda: fn(dst))
        bad = int((dst != src).sum().item())
        row[name] = f"{us:.1f}us {nbytes / us / 1e3:.0f}GB/s bad={bad}"
        del dst
    print(row, flush=True)
