"""Small helpers shared by the benchmarks and examples."""
import subprocess
import threading
import time

import torch


class CudaTimer:
    """CUDA-event timing on one stream (device time, not wall clock)."""

    def __init__(self, stream=None):
        self.stream = stream or torch.cuda.current_stream()
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self.e0.record(self.stream)
        return self

    def __exit__(self, *a):
        self.e1.record(self.stream)

    def ms(self):
        self.e1.synchronize()
        return self.e0.elapsed_time(self.e1)


class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons while a benchmark runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu=0, period_ms=100):
        self.gpu, self.period, self.rows, self.proc = gpu, period_ms, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.rows.append([x.strip() for x in line.split(",")]) for line in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None
        return self

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


def bus_bandwidth(coll, nbytes, seconds, n):
    """Bus bandwidth in GB/s with the reference's factors (tools/perf/ucc_pt_coll_*.cc)."""
    if n <= 1:
        return nbytes / seconds / 1e9
    f = {"allreduce": 2.0 * (n - 1) / n, "allgather": (n - 1) / n, "reduce_scatter": (n - 1) / n, "alltoall": (n - 1) / n}.get(coll, 1.0)
    return nbytes / seconds / 1e9 * f


def wall():
    return time.perf_counter()
