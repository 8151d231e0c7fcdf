"""User-defined (generic) datatypes: contiguous element size + user reduction callback (reference: ucc_dt_create_generic users in
test/gtest/coll/test_allreduce.cc "UserDefinedDt", core/test_dt)."""
import ctypes as C

import numpy as np
import pytest

from ucc_b200 import capi as U
from ucc_b200.harness import UccJob, coll_args


class reduce_cb_params(C.Structure):
    _fields_ = [("mask", C.c_uint64), ("src1", C.c_void_p), ("src2", C.c_void_p), ("dst", C.c_void_p), ("n_vectors", C.c_size_t),
                ("count", C.c_size_t), ("stride", C.c_size_t), ("dt", C.c_void_p), ("cb_ctx", C.c_void_p)]


REDUCE_CB = C.CFUNCTYPE(C.c_int, C.POINTER(reduce_cb_params))


class _reduce(C.Structure):
    _fields_ = [("cb", REDUCE_CB), ("cb_ctx", C.c_void_p)]


class generic_dt_ops(C.Structure):
    _fields_ = [("mask", C.c_uint64), ("flags", C.c_uint64), ("contig_size", C.c_size_t), ("start_pack", C.c_void_p), ("start_unpack", C.c_void_p),
                ("packed_size", C.c_void_p), ("pack", C.c_void_p), ("unpack", C.c_void_p), ("finish", C.c_void_p), ("reduce", _reduce)]


ELEM = np.dtype([("key", np.int32), ("val", np.float64)], align=True)   # 16 bytes: {int32, pad, float64}
calls = {"n": 0}


@REDUCE_CB
def _sum_val_max_key(p):
    """dst[i] = {max(key), sum(val)} over src1 and the n_vectors strided src2 operands"""
    q = p.contents
    calls["n"] += 1
    acc = np.frombuffer((C.c_char * (q.count * ELEM.itemsize)).from_address(q.src1), dtype=ELEM).copy()
    for v in range(q.n_vectors):
        o = np.frombuffer((C.c_char * (q.count * ELEM.itemsize)).from_address(q.src2 + v * q.stride), dtype=ELEM)
        acc["key"] = np.maximum(acc["key"], o["key"])
        acc["val"] += o["val"]
    C.memmove(q.dst, acc.ctypes.data, q.count * ELEM.itemsize)
    return 0


@pytest.fixture(scope="module")
def gdt():
    ops = generic_dt_ops()
    ops.mask = 1
    ops.flags = 1 | 2                      # CONTIG | REDUCE
    ops.contig_size = ELEM.itemsize
    ops.reduce.cb = _sum_val_max_key
    dt = C.c_uint64()
    U.lib.ucc_dt_create_generic.argtypes = [C.POINTER(generic_dt_ops), C.c_void_p, C.POINTER(C.c_uint64)]
    assert U.lib.ucc_dt_create_generic(C.byref(ops), None, C.byref(dt)) == U.UCC_OK
    yield dt.value, ops
    U.ucc_dt_destroy(dt.value)


def _args(coll, src, dst, dt, **kw):
    a = coll_args(coll, dt="int8", src_ptr=src.ctypes.data if src is not None else None, dst_ptr=dst.ctypes.data if dst is not None else None,
                  count_src=len(src) if src is not None else 0, count_dst=len(dst) if dst is not None else 0, **kw)
    a.src.info.datatype = dt
    a.dst.info.datatype = dt
    return a


@pytest.mark.parametrize("n", [2, 3, 5, 8])
@pytest.mark.parametrize("alg", ["knomial", "sra_knomial", "ring", "dbt"])
def test_generic_dt_allreduce(gdt, n, alg):
    dt, _ = gdt
    with UccJob(n, env={"UCC_TL_SHM_TUNE": f"allreduce:@{alg}"}) as j:
        team = j.create_team()
        rng = np.random.default_rng(n)
        for count in (1, 17, 1000):
            src = []
            for r in range(n):
                a = np.zeros(count, ELEM)
                a["key"] = rng.integers(0, 1000, count)
                a["val"] = rng.random(count)
                src.append(a)
            dst = [np.zeros(count, ELEM) for _ in range(n)]
            before = calls["n"]
            req = team.coll([_args("allreduce", src[r], dst[r], dt) for r in range(n)])
            assert req.run() == U.UCC_OK
            req.finalize()
            assert calls["n"] > before                                   # the user callback did the arithmetic
            for r in range(n):
                assert np.array_equal(dst[r]["key"], np.max([s["key"] for s in src], 0))
                assert np.allclose(dst[r]["val"], np.sum([s["val"] for s in src], 0))


def test_generic_dt_data_movement_and_reduce(gdt):
    dt, _ = gdt
    n, count = 4, 33
    with UccJob(n) as j:
        team = j.create_team()
        src = []
        for r in range(n):
            a = np.zeros(count, ELEM)
            a["key"] = r
            a["val"] = np.arange(count) + 100 * r
            src.append(a)
        dst = [np.zeros(count * n, ELEM) for _ in range(n)]
        req = team.coll([_args("allgather", src[r], dst[r], dt) for r in range(n)])
        assert req.run() == U.UCC_OK
        req.finalize()
        for r in range(n):
            assert np.array_equal(dst[r], np.concatenate(src))
        b = [src[2].copy() if r == 2 else np.zeros(count, ELEM) for r in range(n)]
        req = team.coll([_args("bcast", b[r], None, dt, root=2) for r in range(n)])
        assert req.run() == U.UCC_OK
        req.finalize()
        assert all(np.array_equal(x, src[2]) for x in b)
        out = np.zeros(count, ELEM)
        req = team.coll([_args("reduce", src[r], out if r == 1 else None, dt, root=1) for r in range(n)])
        assert req.run() == U.UCC_OK
        req.finalize()
        assert np.array_equal(out["key"], np.full(count, n - 1)) and np.allclose(out["val"], np.sum([s["val"] for s in src], 0))


def test_generic_dt_without_reduce_cb_is_rejected_for_reductions():
    ops = generic_dt_ops()
    ops.mask, ops.flags, ops.contig_size = 1, 1, 8
    dt = C.c_uint64()
    U.lib.ucc_dt_create_generic.argtypes = [C.POINTER(generic_dt_ops), C.c_void_p, C.POINTER(C.c_uint64)]
    assert U.lib.ucc_dt_create_generic(C.byref(ops), None, C.byref(dt)) == U.UCC_OK
    with UccJob(2) as j:
        team = j.create_team()
        x = [np.zeros(4, np.int64) for _ in range(2)]
        a = _args("allreduce", x[0], x[0].copy(), dt.value)
        q = C.POINTER(U.ucc_coll_req_t)()
        assert U.ucc_collective_init(C.byref(a), C.byref(q), team.members[0].team) < 0
        # ... but plain data movement is fine
        g = [np.zeros(8, np.int64) for _ in range(2)]
        s = [np.full(4, r + 1, np.int64) for r in range(2)]
        req = team.coll([_args("allgather", s[r], g[r], dt.value) for r in range(2)])
        assert req.run() == U.UCC_OK
        req.finalize()
        assert np.array_equal(g[0], np.array([1] * 4 + [2] * 4))
    U.ucc_dt_destroy(dt.value)
