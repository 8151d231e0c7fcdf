"""Multi-process (gloo bootstrap, world_size 2 and 3) host collectives through ucc_b200.dist."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [2, 3])
def test_torchrun_host(n):
    port = 29600 + n
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), "cpu"]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert "DIST_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.parametrize("n", [2, 3])
def test_parallel_helpers_host(n):
    """ucc_b200.parallel (DDP buckets, tensor-parallel layers, MoE alltoallv) on host tensors over tl/shm."""
    port = 29650 + n
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "parallel_worker.py"), "cpu"]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert "PARALLEL_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.parametrize("n", [2, 3])
def test_torch_backend_host(n):
    """torch.distributed backend "ucc_b200" (the ProcessGroupUCC role): c10d collectives + torch DDP on host tensors."""
    port = 29670 + n
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "pg_worker.py"), "cpu"]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert "PG_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_torchrun_host_onesided_cma():
    """one-sided alltoall between processes: blocks are read with process_vm_readv straight from the peers' buffers."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dist_worker.py"), "cpu"]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="", UCC_TL_SHM_TUNE="alltoall:@onesided#alltoallv:@onesided#allreduce:@sliding_window", UCC_TL_SHM_CMA="y")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert "DIST_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.parametrize("n,extra", [(3, []), (4, ["-c", "allreduce,alltoallv,reduce_scatter,bcast,gatherv", "-P", "2", "-i", "3", "-N", "8"])])
def test_ucc_test_dist_tool(n, extra):
    """tools/ucc_test_dist.py (the `ucc_test_mpi` role): team kinds world/half/odd_even/reverse x collectives x dtypes x ops x
    in-place x roots, every case checked against the locally computed oracle; the report must show no failures."""
    port = 29690 + n
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "ucc_test_dist.py"), "-t", "world,half,odd_even,reverse",
           "-I", "2", "-m", "8:70000:90", "-r", "all", "-d", "int32,bfloat16,uint8,float64", "-o", "sum,max,lor,band,avg", "-s", "17"] + extra
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert "UCC_TEST_DIST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    rep = out.stdout[out.stdout.index("TEST REPORT"):]
    nums = {ln.split(":")[0].strip(): int(ln.split(":")[1]) for ln in rep.splitlines() if ":" in ln and ln.split(":")[1].strip().isdigit()}
    assert nums["failed"] == 0 and nums["passed"] > 500, nums


@pytest.mark.parametrize("n,ppn,tune", [(4, 2, "allreduce:0-inf:@rab"), (6, 3, "allreduce:0-inf:@split_rail"), (5, 2, None)])
def test_cl_hier_multiprocess(n, ppn, tune):
    """cl/hier schedules with real processes: UCC_B200_FAKE_PPN spreads the ranks of one box over pretend nodes."""
    port = 29710 + n
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "ucc_test_dist.py"), "-t", "world,reverse",
           "-c", "allreduce,bcast,reduce,alltoall,alltoallv,allgatherv,barrier", "-I", "2", "-P", "2", "-i", "2",
           "-m", "8:200000:50", "-r", "all", "-d", "int32,float32", "-o", "sum,max", "-s", "3"]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="", UCC_B200_FAKE_PPN=str(ppn), UCC_CLS="hier,basic",
               UCC_CL_HIER_TLS="shm,self", UCC_CL_BASIC_TLS="shm,self", UCC_COLL_TRACE="info")
    if tune:
        env["UCC_CL_HIER_TUNE"] = tune
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert "UCC_TEST_DIST_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    assert "CL_HIER" in out.stdout + out.stderr      # the hierarchical CL really took collectives


def test_torchrun_host_registered_alltoall_cma():
    """Communicator.register + put-based one-sided alltoall between processes (tl/shm mem_map, process_vm_writev)"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "tests", "dist_worker.py"), "cpu"]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="", UCC_TL_SHM_TUNE="alltoall:@onesided", UCC_TL_SHM_CMA="y", DW_MEMH="1",
               UCC_TL_SHM_LOG_LEVEL="debug")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert "DIST_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
    assert "put into the members' registered destinations" in out.stdout + out.stderr   # (lines of the three processes interleave: no counting)


def test_example_long_context_pipeline():
    """examples/long_context_pipeline.py: ring attention + 1F1B pipeline over send / recv, 4 processes on host tensors"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4", "--master-addr", "127.0.0.1",
           "--master-port", "29621", os.path.join(ROOT, "examples", "long_context_pipeline.py")]
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert "EXAMPLE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
