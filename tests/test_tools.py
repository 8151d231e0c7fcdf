"""ucc_info and ucc_perftest (reference tools/info/ucc_info.c, tools/perf/*): CLI behaviour on host memory."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "ucc_b200", "bin")


def _need(tool):
    p = os.path.join(BIN, tool)
    if not os.path.exists(p):
        subprocess.run(["make", "-C", ROOT, "-j8", "core", "tools"], capture_output=True)
    if not os.path.exists(p):
        pytest.skip(f"{tool} not built")
    return p


def test_ucc_info_flags():
    exe = _need("ucc_info")
    out = subprocess.run([exe, "-v"], capture_output=True, text=True, timeout=60).stdout
    assert re.search(r"1\.\d+", out)
    out = subprocess.run([exe, "-c", "-a"], capture_output=True, text=True, timeout=60).stdout
    assert "UCC_CLS" in out and "UCC_TL_SHM_TUNE" in out
    out = subprocess.run([exe, "-A"], capture_output=True, text=True, timeout=60).stdout
    assert "sra_knomial" in out and "allreduce" in out.lower()
    out = subprocess.run([exe, "-s"], capture_output=True, text=True, timeout=60).stdout
    assert "shm" in out.lower()


def _perftest(args, n=2, port=29500):
    exe = _need("ucc_perftest")
    procs = []
    for r in range(n):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(n), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([exe] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:] + o[-1000:]
    return outs[0][0]


def test_perftest_allreduce_host_sweep():
    out = _perftest(["-c", "allreduce", "-m", "host", "-b", "8", "-e", "4096", "-n", "20", "-w", "5", "-d", "float32", "-o", "sum"], port=29510)
    rows = [ln.split() for ln in out.splitlines() if re.match(r"^\s*\d+\s+\d+", ln)]
    assert len(rows) >= 8, out
    counts = [int(r[0]) for r in rows]
    assert counts[0] == 8 and counts[-1] >= 2048 and all(b == 2 * a for a, b in zip(counts, counts[1:]))
    full = _perftest(["-c", "allreduce", "-m", "host", "-b", "1024", "-e", "1024", "-n", "10", "-w", "2", "-F"], port=29511)
    assert "bandwidth" in full.lower() or "GB/s" in full, full


@pytest.mark.parametrize("coll", ["alltoall", "allgather", "bcast", "barrier", "reduce_scatter"])
def test_perftest_other_colls(coll):
    port = 29520 + hash(coll) % 50
    out = _perftest(["-c", coll, "-m", "host", "-b", "16", "-e", "256", "-n", "10", "-w", "2"], n=3, port=port)
    assert coll in out.lower()


def test_perftest_persistent_inplace_and_executor_ops():
    out = _perftest(["-c", "allreduce", "-m", "host", "-b", "64", "-e", "64", "-n", "10", "-w", "2", "-F", "-i", "-p"], port=29590)
    assert re.search(r"^\s*64\s", out, re.M)
    out = _perftest(["-c", "memcpy", "-m", "host", "-b", "1024", "-e", "1024", "-n", "10", "-w", "2"], n=1, port=29591)
    assert re.search(r"^\s*1024\s", out, re.M)
