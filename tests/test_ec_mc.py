"""EC reduce for every dtype/op + executor task kinds, MC alloc/memcpy/query — cpu everywhere, cuda under -m gpu
(reference: test/gtest/core/test_mc_reduce.cc, test_ec_cuda.cc, test_mc.cc)."""
import ctypes as C
import os

import numpy as np
import pytest

from ucc_b200 import capi as U
from ucc_b200 import internal as I
from ucc_b200.harness import UccJob, NP_DT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INT = ["int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64"]
FLT = ["float32", "float64", "float16"]
CPX = ["float32_complex", "float64_complex"]
OPS_INT = ["sum", "prod", "max", "min", "land", "lor", "lxor", "band", "bor", "bxor"]
OPS_FLT = ["sum", "prod", "max", "min"]


@pytest.fixture(scope="module")
def lib_alive():
    job = UccJob(1)  # keeps mc/ec initialised
    yield job
    job.cleanup()


def np_reduce(op, arrs):
    a = np.stack(arrs)
    f = {"sum": lambda: a.sum(0, dtype=a.dtype), "prod": lambda: a.prod(0, dtype=a.dtype), "max": lambda: a.max(0), "min": lambda: a.min(0),
         "band": lambda: np.bitwise_and.reduce(a, 0), "bor": lambda: np.bitwise_or.reduce(a, 0), "bxor": lambda: np.bitwise_xor.reduce(a, 0),
         "land": lambda: np.logical_and.reduce(a != 0, 0).astype(a.dtype), "lor": lambda: np.logical_or.reduce(a != 0, 0).astype(a.dtype),
         "lxor": lambda: np.logical_xor.reduce(a != 0, 0).astype(a.dtype)}
    return f[op]()


def mk(dt, n, seed):
    rng = np.random.default_rng(seed)
    t = NP_DT[dt]
    if np.issubdtype(t, np.integer):
        return rng.integers(1, 4, n).astype(t)
    if np.issubdtype(t, np.complexfloating):
        return (rng.random(n) + 1j * rng.random(n)).astype(t)
    return (rng.random(n) + 0.5).astype(t)


def reduce_args(dst, srcs, count, dt, op, alpha=None):
    a = I.eee_task_args()
    a.task_type = I.EE_TASK_REDUCE
    a.reduce.dst = dst
    for i, s in enumerate(srcs):
        a.reduce.srcs[i] = s
    a.reduce.n_srcs, a.reduce.count, a.reduce.dt, a.reduce.op = len(srcs), count, U.DT[dt], U.OP[op]
    if alpha is not None:
        a.flags = I.EEE_FLAG_ALPHA
        a.reduce.alpha = alpha
    return a


@pytest.mark.parametrize("dt,op", [(d, o) for d in INT for o in OPS_INT] + [(d, o) for d in FLT for o in OPS_FLT] + [(d, o) for d in CPX for o in ("sum", "prod")])
@pytest.mark.parametrize("nsrc", [2, 5])
def test_ec_cpu_reduce(lib_alive, dt, op, nsrc):
    n = 77
    srcs = [mk(dt, n, 3 * i + 1) for i in range(nsrc)]
    dst = np.zeros(n, NP_DT[dt])
    ex = I.Executor(U.UCC_EE_CPU_THREAD)
    ex.run(reduce_args(dst.ctypes.data, [s.ctypes.data for s in srcs], n, dt, op))
    ex.close()
    exp = np_reduce(op, srcs)
    if dt == "float16":
        assert np.allclose(dst.astype(np.float32), exp.astype(np.float32), rtol=2e-2)
    elif "float" in dt:
        assert np.allclose(dst, exp, rtol=1e-5)
    else:
        assert np.array_equal(dst, exp)


def test_ec_cpu_int_avg(lib_alive):
    """AVG on integers = truncated sum / N (reference: d[i] * (1/N) in double precision); casting 1/N to the integer type gave 0"""
    n = 1000
    ex = I.Executor(U.UCC_EE_CPU_THREAD)
    for dt in ("int32", "int64", "uint8", "int16"):
        s = [mk(dt, n, 5 * i + 2) for i in range(3)]
        d = np.zeros(n, NP_DT[dt])
        ex.run(reduce_args(d.ctypes.data, [x.ctypes.data for x in s], n, dt, "avg", alpha=1.0 / 3))
        tot = sum(x.astype(np.int64) for x in s).astype(NP_DT[dt]).astype(np.float64)   # the sum wraps in the storage type first
        assert np.array_equal(d, (tot * (1.0 / 3)).astype(NP_DT[dt])), dt
    ex.close()


def test_ec_cpu_alpha_strided_multi_copy(lib_alive):
    n = 100
    ex = I.Executor(U.UCC_EE_CPU_THREAD)
    s = [mk("float64", n, i) for i in range(3)]
    d = np.zeros(n)
    ex.run(reduce_args(d.ctypes.data, [x.ctypes.data for x in s], n, "float64", "avg", alpha=1.0 / 3))
    assert np.allclose(d, sum(s) / 3)
    # strided: src2 holds 4 vectors back to back
    s1 = mk("int32", n, 9)
    s2 = mk("int32", 4 * n, 10)
    d = np.zeros(n, np.int32)
    a = I.eee_task_args()
    a.task_type = I.EE_TASK_REDUCE_STRIDED
    rs = a.reduce_strided
    rs.dst, rs.src1, rs.src2, rs.stride, rs.count, rs.dt, rs.op, rs.n_src2 = d.ctypes.data, s1.ctypes.data, s2.ctypes.data, n * 4, n, U.DT["int32"], U.OP["sum"], 4
    ex.run(a)
    assert np.array_equal(d, s1 + s2.reshape(4, n).sum(0))
    # multi dst
    a = I.eee_task_args()
    a.task_type = I.EE_TASK_REDUCE_MULTI_DST
    xs = [mk("float32", 10 + j, j) for j in range(3)]
    ys = [mk("float32", 10 + j, 50 + j) for j in range(3)]
    ds = [np.zeros(10 + j, np.float32) for j in range(3)]
    for j in range(3):
        a.reduce_multi_dst.dst[j], a.reduce_multi_dst.src1[j], a.reduce_multi_dst.src2[j], a.reduce_multi_dst.counts[j] = ds[j].ctypes.data, xs[j].ctypes.data, ys[j].ctypes.data, 10 + j
    a.reduce_multi_dst.dt, a.reduce_multi_dst.op, a.reduce_multi_dst.n_bufs = U.DT["float32"], U.OP["sum"], 3
    ex.run(a)
    for j in range(3):
        assert np.allclose(ds[j], xs[j] + ys[j])
    # copy + copy multi
    src = mk("uint8", 1000, 1)
    dst = np.zeros(1000, np.uint8)
    a = I.eee_task_args()
    a.task_type = I.EE_TASK_COPY
    a.copy.dst, a.copy.src, a.copy.len = dst.ctypes.data, src.ctypes.data, 1000
    ex.run(a)
    assert np.array_equal(dst, src)
    a = I.eee_task_args()
    a.task_type = I.EE_TASK_COPY_MULTI
    dsts = [np.zeros(100 * (j + 1), np.uint8) for j in range(4)]
    srcs = [mk("uint8", 100 * (j + 1), j) for j in range(4)]
    for j in range(4):
        a.copy_multi.src[j], a.copy_multi.dst[j], a.copy_multi.counts[j] = srcs[j].ctypes.data, dsts[j].ctypes.data, 100 * (j + 1)
    a.copy_multi.num_vectors = 4
    ex.run(a)
    for j in range(4):
        assert np.array_equal(dsts[j], srcs[j])
    ex.close()


def test_mc_cpu(lib_alive):
    h = C.POINTER(I.mc_buffer_header)()
    for size in (64, 1 << 20, (1 << 20) + 1):  # pooled and plain
        assert I.lib.ucc_mc_alloc(C.byref(h), size, U.UCC_MEMORY_TYPE_HOST) == U.UCC_OK
        assert h.contents.addr and h.contents.addr % 64 == 0
        I.lib.ucc_mc_memset(h.contents.addr, 0x5a, size, U.UCC_MEMORY_TYPE_HOST)
        assert C.string_at(h.contents.addr, 4) == b"\x5a" * 4
        assert I.lib.ucc_mc_free(h) == U.UCC_OK
    buf = np.zeros(16)
    at = I.mem_attr()
    at.field_mask = 1
    assert I.lib.ucc_mc_get_mem_attr(buf.ctypes.data, C.byref(at)) == U.UCC_OK and at.mem_type == U.UCC_MEMORY_TYPE_HOST


# ------------------------------------------------------------------ CUDA
def _torch():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch


TDT = {"int8": "int8", "int16": "int16", "int32": "int32", "int64": "int64", "uint8": "uint8", "float32": "float32", "float64": "float64",
       "float16": "float16", "bfloat16": "bfloat16", "float32_complex": "complex64", "float64_complex": "complex128"}


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["interruptible", "persistent"])
@pytest.mark.parametrize("dt,op", [(d, o) for d in ["int8", "int16", "int32", "int64", "uint8"] for o in OPS_INT] +
                         [(d, o) for d in ["float32", "float64", "float16", "bfloat16"] for o in OPS_FLT] + [(d, o) for d in CPX for o in ("sum", "prod")])
def test_ec_cuda_reduce(lib_alive, mode, dt, op):
    torch = _torch()
    tdt = getattr(torch, TDT[dt])
    for n, off in ((1000, 0), (4099, 1)):  # aligned vector path and misaligned scalar path
        g = torch.Generator().manual_seed(n)
        if tdt.is_complex:
            srcs = [torch.view_as_complex(torch.rand(n + off, 2, generator=g).to(torch.float32 if tdt == torch.complex64 else torch.float64)).cuda()[off:] for _ in range(3)]
        elif tdt.is_floating_point:
            srcs = [(torch.rand(n + off, generator=g) + 0.5).to(tdt).cuda()[off:] for _ in range(3)]
        else:
            srcs = [torch.randint(1, 4, (n + off,), generator=g).to(tdt).cuda()[off:] for _ in range(3)]
        dst = torch.zeros(n + off, dtype=tdt, device="cuda")[off:]
        stream = torch.cuda.Stream() if mode == "persistent" else None
        ex = I.Executor(U.UCC_EE_CUDA_STREAM, stream.cuda_stream if stream else None)
        ex.run(reduce_args(dst.data_ptr(), [s.data_ptr() for s in srcs], n, dt, op))
        ex.close()
        torch.cuda.synchronize()
        ref = torch.stack([s.cpu().double() if tdt.is_floating_point else (s.cpu().to(torch.complex128) if tdt.is_complex else s.cpu().long()) for s in srcs])
        if op == "sum":
            exp = ref.sum(0)
        elif op == "prod":
            exp = ref.prod(0)
        elif op == "max":
            exp = ref.max(0).values
        elif op == "min":
            exp = ref.min(0).values
        else:
            exp = torch.from_numpy(np_reduce(op, [s.cpu().numpy() for s in srcs])).long()
        got = dst.cpu()
        if tdt.is_complex:
            assert torch.allclose(got.to(torch.complex128), exp, rtol=1e-4)
        elif tdt.is_floating_point:
            assert torch.allclose(got.double(), exp, rtol=3e-2 if dt in ("float16", "bfloat16") else 1e-5)
        else:
            assert torch.equal(got.long(), exp.to(tdt).long())


@pytest.mark.gpu
def test_ec_cuda_copy_alpha_events_mc(lib_alive):
    torch = _torch()
    ex = I.Executor(U.UCC_EE_CUDA_STREAM)
    s = [torch.rand(5000, device="cuda") for _ in range(4)]
    d = torch.zeros(5000, device="cuda")
    ex.run(reduce_args(d.data_ptr(), [x.data_ptr() for x in s], 5000, "float32", "avg", alpha=0.25))
    assert torch.allclose(d, sum(s) / 4)
    a = I.eee_task_args()
    a.task_type = I.EE_TASK_COPY_MULTI
    srcs = [torch.rand(1000 * (j + 1), device="cuda") for j in range(5)]
    dsts = [torch.zeros(1000 * (j + 1), device="cuda") for j in range(5)]
    for j in range(5):
        a.copy_multi.src[j], a.copy_multi.dst[j], a.copy_multi.counts[j] = srcs[j].data_ptr(), dsts[j].data_ptr(), 4000 * (j + 1)
    a.copy_multi.num_vectors = 5
    ex.run(a)
    torch.cuda.synchronize()
    for j in range(5):
        assert torch.equal(srcs[j], dsts[j])
    ex.close()
    # wait-only executor on a user stream must release the stream on stop
    st = torch.cuda.Stream()
    ex = I.Executor(U.UCC_EE_CUDA_STREAM, st.cuda_stream, task_types=0)
    ex.close()
    st.synchronize()
    # events
    ev = C.c_void_p()
    assert I.lib.ucc_ec_create_event(C.byref(ev), U.UCC_EE_CUDA_STREAM) == U.UCC_OK
    assert I.lib.ucc_ec_event_post(st.cuda_stream, ev, U.UCC_EE_CUDA_STREAM) == U.UCC_OK
    while I.lib.ucc_ec_event_test(ev, U.UCC_EE_CUDA_STREAM) == U.UCC_INPROGRESS:
        pass
    I.lib.ucc_ec_destroy_event(ev, U.UCC_EE_CUDA_STREAM)
    # mc/cuda: pooled alloc, query, memcpy, memset
    h = C.POINTER(I.mc_buffer_header)()
    assert I.lib.ucc_mc_alloc(C.byref(h), 4096, U.UCC_MEMORY_TYPE_CUDA) == U.UCC_OK
    at = I.mem_attr()
    at.field_mask = 7
    assert I.lib.ucc_mc_get_mem_attr(h.contents.addr, C.byref(at)) == U.UCC_OK and at.mem_type == U.UCC_MEMORY_TYPE_CUDA and at.alloc_length >= 4096
    host = np.arange(1024, dtype=np.float32)
    back = np.zeros(1024, np.float32)
    assert I.lib.ucc_mc_memcpy(h.contents.addr, host.ctypes.data, 4096, U.UCC_MEMORY_TYPE_CUDA, U.UCC_MEMORY_TYPE_HOST) == U.UCC_OK
    assert I.lib.ucc_mc_memcpy(back.ctypes.data, h.contents.addr, 4096, U.UCC_MEMORY_TYPE_HOST, U.UCC_MEMORY_TYPE_CUDA) == U.UCC_OK
    assert np.array_equal(host, back)
    assert I.lib.ucc_mc_free(h) == U.UCC_OK
    at.field_mask = 1
    t = torch.zeros(4, device="cuda")
    assert I.lib.ucc_mc_get_mem_attr(t.data_ptr(), C.byref(at)) == U.UCC_OK and at.mem_type == U.UCC_MEMORY_TYPE_CUDA


@pytest.mark.gpu
def test_ec_cuda_persistent_multi_worker():
    """reference ec_cuda_executor.cu:125-188: EXEC_NUM_WORKERS blocks share one task ring.  32 tasks posted back to back to a
    4-worker persistent executor (own process: the EC configuration is read once per library instance)."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from ucc_b200 import capi as U, internal as I
from test_ec_mc import reduce_args
cfg = U.handle(); U.check(U.ucc_lib_config_read(None, None, C.byref(cfg)), "cfg")
p = U.ucc_lib_params_t(); p.mask, p.thread_mode = U.UCC_LIB_PARAM_FIELD_THREAD_MODE, U.UCC_THREAD_SINGLE
lib = U.handle(); U.check(U.ucc_init_version(U.UCC_API_MAJOR, U.UCC_API_MINOR, C.byref(p), cfg, C.byref(lib)), "init")
st = torch.cuda.Stream()
n = 20000
srcs = [[torch.rand(n + 7 * k, device="cuda") for _ in range(3)] for k in range(32)]
dsts = [torch.zeros(n + 7 * k, device="cuda") for k in range(32)]
exp = [sum(srcs[k]) for k in range(32)]
torch.cuda.synchronize()     # (a device-wide synchronize while the persistent kernel runs would wait for it forever)
ex = I.Executor(U.UCC_EE_CUDA_STREAM, st.cuda_stream)
tasks = []
for k in range(32):
    a = reduce_args(dsts[k].data_ptr(), [s.data_ptr() for s in srcs[k]], n + 7 * k, "float32", "sum")
    t = C.c_void_p()
    U.check(I.lib.ucc_ee_executor_task_post(ex.h, C.byref(a), C.byref(t)), "post")
    tasks.append((a, t))
for a, t in tasks:
    while I.lib.ucc_ee_executor_task_test(t) == U.UCC_INPROGRESS:
        pass
    assert I.lib.ucc_ee_executor_task_test(t) == U.UCC_OK
    I.lib.ucc_ee_executor_task_finalize(t)
ex.close()
torch.cuda.synchronize()
for k in range(32):
    assert torch.allclose(dsts[k], exp[k]), k
print("MULTI_WORKER_OK")
''' % (ROOT, ROOT)
    env = dict(os.environ, UCC_EC_CUDA_EXEC_NUM_WORKERS="4", UCC_EC_CUDA_EXEC_MAX_TASKS="64", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=60)
    assert "MULTI_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
