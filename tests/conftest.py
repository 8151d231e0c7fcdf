import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "multigpu: test needs >= 2 CUDA devices")


def _build_once():
    lib = os.path.join(ROOT, "ucc_b200", "lib", "libucc.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-C", ROOT, "core", "-j8"], stdout=subprocess.DEVNULL)


_build_once()


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
