"""tl/nvl collective kernels compiled as host C++ (NVL_HOST_EMU, tests/emu/nvl_emu.cpp): one OS thread per CUDA thread, N heaps
in one address space, the NVSwitch multicast window emulated.  Checks the indexing / phase / flag logic of every reduction kernel
without a GPU: the kernels validated on B200s (staged, one-shot, ring / rhd, zero-copy, the exchange kernel's pull / NVLS push / ring modes) as controls, and the ones written after
the GPU budget ran out (nvls_pipe, symmetric-memory allreduce / reduce_scatter / allgather, zero-copy push allgather / alltoall)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_fp16.h")):
        pytest.skip("CUDA headers not installed")
    exe = tmp_path_factory.mktemp("emu") / "nvl_emu"
    k = os.path.join(ROOT, "src", "components", "tl", "nvl", "kernels")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-pthread", f"-I{CUDA_INC}", f"-I{k}", f"-I{ROOT}/include", f"-I{ROOT}/src",
                         os.path.join(ROOT, "tests", "emu", "nvl_emu.cpp"), "-o", str(exe)], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-4000:]
    return str(exe)


@pytest.mark.parametrize("what", ["staged", "xchg", "pipe", "symm", "push", "oneshot_rs", "soak", "soak2"])
def test_nvl_kernels_host_emulation(emu, what):
    out = subprocess.run([emu, what], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "NVL_EMU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
