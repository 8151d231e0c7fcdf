#!/usr/bin/env python3
"""Box probe: what the GPU host exposes for an NVLink-native collective library.

Run under torchrun (N ranks, one per GPU).  Everything is written to
gpurun_out/probe_rank<r>.txt; rank 0 also writes gpurun_out/probe_nccl.json
(tier-B NCCL allreduce sweep, device timed, max over ranks).
"""
import ctypes
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
lrank = int(os.environ.get("LOCAL_RANK", 0))
os.makedirs("gpurun_out", exist_ok=True)
out = open(f"gpurun_out/probe_rank{rank}.txt", "w")


def P(*a):
    print(*a, file=out, flush=True)


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout
    except Exception as e:  # noqa
        return f"<{e}>"


if rank == 0:
    P(sh("nvidia-smi -L"))
    P(sh("nvidia-smi topo -m"))
    P(sh("nvidia-smi nvlink -s -i 0 | head -30"))
    P(sh("uname -a; cat /proc/sys/kernel/yama/ptrace_scope 2>&1; ls -la /dev/shm | head; df -h /dev/shm | tail -1"))
    P(sh("ls /dev/nvidia* ; ls /dev/nvidia-caps* 2>&1 | head"))
    P(sh("ipcs -l | head -20"))

torch.cuda.set_device(lrank)
cu = ctypes.CDLL("libcuda.so.1")
cu.cuInit(0)
dev = ctypes.c_int()
cu.cuDeviceGet(ctypes.byref(dev), lrank)
for name, aid in [("VMM", 102), ("POSIX_FD", 103), ("FABRIC", 128), ("MULTICAST", 132),
                  ("RDMA_VMM", 110), ("WAIT_VALUE_NOR", 123), ("COOP_LAUNCH", 95),
                  ("MAX_SMEM_OPTIN", 97), ("SM_COUNT", 16), ("L2", 38), ("CLOCK_KHZ", 13),
                  ("MEM_CLOCK_KHZ", 36), ("MEMPOOLS", 115), ("CONCURRENT_MANAGED", 89),
                  ("HOST_NATIVE_ATOMIC", 86)]:
    v = ctypes.c_int(-1)
    r = cu.cuDeviceGetAttribute(ctypes.byref(v), aid, dev)
    P(f"attr {name}({aid}) = {v.value} (rc {r})")

# pidfd syscalls
libc = ctypes.CDLL(None, use_errno=True)
SYS_pidfd_open, SYS_pidfd_getfd = 434, 438
fd = libc.syscall(SYS_pidfd_open, os.getpid(), 0)
P("pidfd_open(self) ->", fd, "errno", ctypes.get_errno())
if fd >= 0:
    fd2 = libc.syscall(SYS_pidfd_getfd, fd, 0, 0)
    P("pidfd_getfd(self, 0) ->", fd2, "errno", ctypes.get_errno())

if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    P("peer access 0<->1:", torch.cuda.can_device_access_peer(0, 1) if torch.cuda.device_count() > 1 else None)

    # torch symmetric memory / multicast probe
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(1 << 20, dtype=torch.float32, device=f"cuda:{lrank}")
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        P("symm_mem ok; multicast_ptr =", getattr(hdl, "multicast_ptr", None))
    except Exception as e:  # noqa
        P("symm_mem failed:", repr(e)[:400])

    # NCCL sweep (tier B)
    res = []
    sizes = [1 << s for s in range(10, 31, 2)]
    for nbytes in sizes:
        n = nbytes // 4
        x = torch.ones(n, dtype=torch.float32, device="cuda")
        iters = 50 if nbytes <= (1 << 22) else (20 if nbytes <= (1 << 26) else 8)
        for _ in range(5):
            dist.all_reduce(x)
        torch.cuda.synchronize()
        dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(iters):
            dist.all_reduce(x)
        ev[1].record()
        torch.cuda.synchronize()
        us = ev[0].elapsed_time(ev[1]) * 1e3 / iters
        tt = torch.tensor([us], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        us = tt.item()
        bus = nbytes / us / 1e3 * 2 * (world - 1) / world
        res.append({"bytes": nbytes, "us": us, "busbw": bus})
        del x
    if rank == 0:
        json.dump({"n": world, "allreduce_f32": res}, open("gpurun_out/probe_nccl.json", "w"), indent=1)
        P(json.dumps(res))

    # peer copy bw
    if rank == 0 and torch.cuda.device_count() > 1:
        a = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0")
        b = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:1")
        for _ in range(3):
            b.copy_(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        P("peer copy 256MiB GB/s:", (1 << 28) * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    dist.barrier()
    dist.destroy_process_group()
P("done", time.time())
