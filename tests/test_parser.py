"""Config system (model: reference test/gtest/utils/test_parser.cc, core/test_lib_config.cc, core/test_context_config.cc):
typed fields (memunits, ranged uint, pipeline params, ternary, allow lists), env precedence, the UCC_TLS style
sub-prefix fall-back and context_config_modify addressing of CL/TL tables."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from ucc_b200 import capi as U
from ucc_b200.harness import UccJob, coll_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libc = C.CDLL(None)
libc.open_memstream.restype = C.c_void_p
libc.open_memstream.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]
libc.fclose.argtypes = [C.c_void_p]
U.lib.ucc_context_config_print.argtypes = [U.handle, C.c_void_p, C.c_char_p, C.c_int]
U.lib.ucc_context_config_print.restype = None


def ctx_cfg_dump(lib, modify=()):
    cfg = U.handle()
    assert U.ucc_context_config_read(lib, None, C.byref(cfg)) == U.UCC_OK
    sts = [U.ucc_context_config_modify(cfg, comp.encode() if comp else None, k.encode(), v.encode()) for comp, k, v in modify]
    buf, size = C.c_char_p(), C.c_size_t()
    f = libc.open_memstream(C.byref(buf), C.byref(size))
    U.lib.ucc_context_config_print(cfg, f, b"", 1)
    libc.fclose(f)
    U.ucc_context_config_release(cfg)
    return C.string_at(buf, size.value).decode(), sts


def test_typed_fields_roundtrip():
    with UccJob(1, with_ctx_oob=False) as j:
        lib = j.procs[0].lib
        out, sts = ctx_cfg_dump(lib, [("tl/shm", "CELL_SIZE", "16k"), ("tl/shm", "ALLREDUCE_KN_RADIX", "0-4k:host:8,4k-inf:2"),
                                      ("tl/shm", "ALLREDUCE_SRA_KN_PIPELINE", "thresh=64k:fragsize=32k:nfrags=4:pdepth=2:ordered"),
                                      ("tl/shm", "REDUCE_AVG_PRE_OP", "y")])
        assert all(s == U.UCC_OK for s in sts), sts
        assert "UCC_TL_SHM_CELL_SIZE=16K" in out
        assert "UCC_TL_SHM_ALLREDUCE_KN_RADIX=0-4K:host:8,4K-inf:2" in out or "0-4k:host:8" in out.lower()
        assert "thresh=64K:fragsize=32K:nfrags=4:pdepth=2:ordered" in out
        assert "UCC_TL_SHM_REDUCE_AVG_PRE_OP=y" in out
        # bad values / unknown fields are rejected, the table keeps its old value
        out2, sts = ctx_cfg_dump(lib, [("tl/shm", "CELL_SIZE", "banana"), ("tl/shm", "NO_SUCH", "1"), ("tl/nosuch", "CELL_SIZE", "1k")])
        assert all(s != U.UCC_OK for s in sts), sts
        assert "UCC_TL_SHM_CELL_SIZE=8K" in out2


def _run_snippet(code, env):
    e = dict(os.environ, PYTHONPATH=ROOT, **env)
    out = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


SNIP = (
    "import ctypes as C\n"
    "from ucc_b200 import capi as U\n"
    "from ucc_b200.harness import UccJob\n"
    "j = UccJob(1, with_ctx_oob=False)\n"
    "U.lib.ucc_context_config_print.argtypes = [U.handle, C.c_void_p, C.c_char_p, C.c_int]\n"
    "libc = C.CDLL(None); libc.fdopen.restype = C.c_void_p\n"
    "cfg = U.handle(); assert U.ucc_context_config_read(j.procs[0].lib, None, C.byref(cfg)) == 0\n"
    "fp = libc.fdopen(1, b'w'); U.lib.ucc_context_config_print(cfg, fp, b'', 1); libc.fflush(C.c_void_p(fp))\n")


def test_env_prefix_fallback_and_precedence():
    # UCC_TUNE-style short name reaches the TL table through the sub-prefix fall-back; the full name wins over it
    out = _run_snippet(SNIP, {"UCC_N_CELLS": "64"})
    assert "UCC_TL_SHM_N_CELLS=64" in out
    out = _run_snippet(SNIP, {"UCC_N_CELLS": "64", "UCC_TL_SHM_N_CELLS": "32"})
    assert "UCC_TL_SHM_N_CELLS=32" in out


def test_tls_allow_list_and_negation():
    code = (
        "import numpy as np\n"
        "from ucc_b200 import capi as U\n"
        "from ucc_b200.harness import UccJob, coll_args\n"
        "j = UccJob(2); t = j.create_team()\n"
        "s = [np.full(4, r + 1, np.int32) for r in range(2)]; d = [np.zeros(4, np.int32) for _ in range(2)]\n"
        "q = t.coll([coll_args('allreduce', s[r], d[r], dt='int32') for r in range(2)]); print('st', q.run()); q.finalize(); print(d[0][0])\n")
    out = _run_snippet(code, {"UCC_TLS": "shm,self", "UCC_COLL_TRACE": "info"})
    assert "st 0" in out and "\n3\n" in out and "TL_SHM" in out
    out = _run_snippet(code, {"UCC_TLS": "^nvl,nccl", "UCC_COLL_TRACE": "info"})
    assert "st 0" in out and "TL_SHM" in out


def test_unused_env_warning():
    out = subprocess.run([sys.executable, "-c", "from ucc_b200.harness import UccJob; UccJob(1, with_ctx_oob=False)"],
                         env=dict(os.environ, PYTHONPATH=ROOT, UCC_TL_SHM_NO_SUCH_KNOB="1", UCC_WARN_UNUSED_ENV_VARS="y"), capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    assert "UCC_TL_SHM_NO_SUCH_KNOB" in (out.stdout + out.stderr)
