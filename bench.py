#!/usr/bin/env python3
"""Headline benchmark: ucc_perftest-style allreduce bus bandwidth / latency on CUDA buffers.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric (BASELINE.json): allreduce bus GB/s = S/t * 2(N-1)/N (reference formula,
tools/perf/ucc_pt_coll_allreduce.cc:84-93) at 1 GiB float32, plus a 1 KB..1 GiB sweep with latency.
One rank per GPU (torchrun for N > 1).  A "step" is one allreduce of the headline size through the
library's public C API (ucc_collective_init / triggered_post / test / finalize) executed by the fused
tl/nvl NVLink kernel.  Device time = CUDA events on the launching stream, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=1 << 30, help="headline message size in bytes")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--e2e-chunks", type=int, default=1,
                    help="end-to-end step: > 1 cuts the step into chunks whose H2D copy overlaps the previous chunk's allreduce (not validated on hardware yet)")
    ap.add_argument("--symm", default="", help="opt-in (not yet measured): put the device buffers into a symmetric user region of this size, "
                    "e.g. 3G (Communicator(symm_size=...).symm_empty, the ncclMemAlloc analogue): in-place in-switch allreduce")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nccl", action="store_true")
    ap.add_argument("--out", default="")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def main():
    a = parse()
    if a.impl == "reference":
        # openucx/ucc is an autotools C library that needs UCX/UCS (+MPI for ucc_perftest); none of them is in the
        # image and `pip install /root/reference` fails ("Neither 'setup.py' nor 'pyproject.toml' found") - see DESIGN.md
        print(json.dumps({"impl": "reference", "unavailable": "openucx/ucc needs autotools+UCX+MPI to build; not pip-installable offline (see DESIGN.md)"}))
        return 0

    import torch
    import torch.distributed as dist
    from ucc_b200 import capi as U
    from ucc_b200.dist import Communicator, init_distributed

    rank, world, lrank = init_distributed()
    if world != a.gpus and rank == 0:
        print(f"warning: --gpus {a.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    N = world
    dev = torch.device("cuda", torch.cuda.current_device())
    dt = getattr(torch, a.dtype)
    esz = torch.empty(0, dtype=dt).element_size()
    comm = Communicator(symm_size=a.symm) if a.symm else Communicator()
    use_symm = bool(a.symm) and comm.symm_region() is not None

    def dev_empty(cnt):
        return comm.symm_empty(cnt, dt) if use_symm else torch.empty(cnt, dtype=dt, device=dev)
    stream = torch.cuda.Stream()
    launches = {"n": 0}

    def maxr(x):
        if N == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def busbw(nbytes, us):
        f = 2.0 * (N - 1) / N if N > 1 else 1.0  # N=1: no peers, report plain S/t
        return nbytes / us / 1e3 * f

    def time_ours(nbytes, iters, warm, inplace=False):
        """device time per allreduce (us, max over ranks): `iters` requests posted back to back on one stream"""
        cnt = nbytes // esz
        if use_symm:
            comm.symm_reset()           # earlier buffers are dead by now
        src = dev_empty(cnt).fill_(1) if use_symm else torch.ones(cnt, dtype=dt, device=dev)
        dst = src if inplace else dev_empty(cnt)
        reqs = [comm.allreduce_init(src, dst) for _ in range(warm + iters)]
        with torch.cuda.stream(stream):
            for r in reqs[:warm]:
                r.post_on_stream(stream)
            for r in reqs[:warm]:
                r.wait()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in reqs[warm:]:
                r.post_on_stream(stream, wait_posted=False)
            for r in reqs[warm:]:
                r.wait_posted()      # zero-copy kernels enter the stream once the peers' buffers are mapped
            e1.record(stream)
            for r in reqs[warm:]:
                r.wait()
            torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        for r in reqs:
            r.finalize()
        ok = bool(torch.allclose(dst[:16].float().cpu(), torch.full((16,), float(N)))) if not inplace else True
        del src, dst
        return maxr(us), ok

    def time_nccl(nbytes, iters, warm):
        if N == 1 or a.no_nccl:
            return None
        cnt = nbytes // esz
        x = torch.ones(cnt, dtype=dt, device=dev)
        with torch.cuda.stream(stream):
            for _ in range(warm):
                dist.all_reduce(x)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                dist.all_reduce(x)
            e1.record(stream)
            torch.cuda.synchronize()
        del x
        return maxr(e0.elapsed_time(e1) * 1e3 / iters)

    # ------------------------------------------------------------------ headline (kernel-only, device timed)
    S = a.size
    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    us, ok = time_ours(S, a.steps, a.warmup)
    launches["n"] = a.steps
    value = busbw(S, us)
    nccl_us = time_nccl(S, a.steps, a.warmup)

    # ------------------------------------------------------------------ end to end through the public API
    e2e = None
    if not a.no_e2e:
        cnt = S // esz
        host = torch.ones(cnt, dtype=dt).pin_memory()
        if use_symm:
            comm.symm_reset()
        src = dev_empty(cnt)
        dst = dev_empty(cnt)
        out = torch.empty(16, dtype=dt).pin_memory()

        # The step is cut into chunks so that the H2D copy of chunk c+1 (copy engine, its own stream) overlaps the
        # allreduce of chunk c (NVLink kernel): the whole input still crosses PCIe and the whole vector is reduced.
        C_ = a.e2e_chunks if cnt % a.e2e_chunks == 0 and S >= (64 << 20) else 1
        h2d = torch.cuda.Stream()
        evs = [torch.cuda.Event() for _ in range(C_)]
        per = cnt // C_

        def step_simple():                                         # the validated default (C_ == 1)
            with torch.cuda.stream(stream):
                src.copy_(host, non_blocking=True)                 # H2D of this step's input from pinned memory
                r = comm.allreduce_init(src, dst)                  # ucc_collective_init
                r.post_on_stream(stream)                           # ucc_collective_triggered_post
                r.wait()                                           # ucc_collective_test + ucc_context_progress
                r.finalize()                                       # ucc_collective_finalize
                out.copy_(dst[:16], non_blocking=True)             # D2H read of the result
                stream.synchronize()
            return out[0].item()

        def step_chunked():
            with torch.cuda.stream(h2d):
                for c in range(C_):                                # H2D of this step's input from pinned memory
                    src[c * per:(c + 1) * per].copy_(host[c * per:(c + 1) * per], non_blocking=True)
                    evs[c].record(h2d)
            reqs = []
            with torch.cuda.stream(stream):
                for c in range(C_):
                    stream.wait_event(evs[c])
                    r = comm.allreduce_init(src[c * per:(c + 1) * per], dst[c * per:(c + 1) * per])   # ucc_collective_init
                    r.post_on_stream(stream)                       # ucc_collective_triggered_post
                    reqs.append(r)
                for r in reqs:
                    r.wait()                                       # ucc_collective_test + ucc_context_progress
                    r.finalize()                                   # ucc_collective_finalize
                out.copy_(dst[:16], non_blocking=True)             # D2H read of the result
                stream.synchronize()
            return out[0].item()

        step = step_chunked if C_ > 1 else step_simple
        for _ in range(max(3, min(a.warmup, 5))):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
        e2e_us = maxr((time.perf_counter() - t0) * 1e6 / a.steps)
        e2e = {"value": busbw(S, e2e_us), "unit": "GB/s", "h2d_bytes_per_step": S, "d2h_bytes_per_step": 16 * esz, "us_per_step": e2e_us, "chunks": C_}
        del host, src, dst

    clocks = sampler.stop()   # sampled across the device-timed headline steps and the end-to-end steps

    # ------------------------------------------------------------------ sweep (latency + bus bandwidth vs size)
    sweep = []
    if not a.no_sweep:
        nb = 1 << 10
        while nb <= (1 << 30):
            iters = 50 if nb <= (1 << 22) else (20 if nb <= (1 << 26) else 8)
            u, okk = time_ours(nb, iters, 5)
            nu = time_nccl(nb, iters, 5)
            sweep.append({"bytes": nb, "us": round(u, 2), "busbw": round(busbw(nb, u), 2), "ok": okk,
                          "nccl_us": round(nu, 2) if nu else None, "nccl_busbw": round(busbw(nb, nu), 2) if nu else None})
            nb <<= 2

    res = {
        "metric": "allreduce_busbw_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": N, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(us / 1e3, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.dtype,
        "data": "synthetic (ones), CUDA device buffers", "impl": "ours",
        "config": {"model": "ucc_perftest allreduce", "collective": "allreduce", "op": "sum", "message_bytes": S, "global_batch": S * N,
                   "seq_len": S // esz, "parallelism": f"dp{N}", "timing": "cuda events on the posting stream, max over ranks",
                   "l2": "message (1 GiB) is larger than the 126 MB L2", "algorithm": "tl/nvl fused kernel chosen by coll_score",
                   "symmetric_memory": use_symm},
        "correct": bool(ok), "clocks": clocks, "gpu_launches": launches["n"],
        "latency_us": round(us, 2), "roofline_frac_of_900": round(value / 900.0, 4) if N > 1 else None, "roofline_frac_of_measured_770": round(value / 770.0, 4) if N > 1 else None,
        "roofline_frac_of_hbm_copy_6478": round(2.0 * value / 6478.3, 4) if N == 1 else None,
        "nccl_same_box": {"us": round(nccl_us, 2), "busbw": round(busbw(S, nccl_us), 2)} if nccl_us else None,
    }
    if e2e:
        res["e2e"] = e2e
    if sweep:
        res["sweep"] = sweep
    comm.destroy()
    if rank == 0:
        print(json.dumps(res), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
    if dist.is_initialized():
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
