"""Host-memory collectives through the full stack (core -> cl/basic -> tl/shm | tl/self),
N emulated ranks in one process, numpy as the oracle.  Mirrors the reference's per-collective
typed gtests (test/gtest/coll/test_*.cc): dtype x op, in-place, persistent, several team sizes,
algorithm-forced variants through UCC_TL_SHM_TUNE."""
import numpy as np
import pytest

from ucc_b200 import capi as U
from ucc_b200.harness import UccJob, coll_args, NP_DT

SIZES = [1, 2, 3, 4, 7, 8]
COUNTS = [1, 5, 256, 4099]


@pytest.fixture(scope="module")
def job():
    j = UccJob(8)
    yield j
    j.cleanup()


@pytest.fixture(scope="module")
def teams(job):
    return {n: job.create_team(range(n)) for n in SIZES}


def rnd(dt, n, seed):
    rng = np.random.default_rng(seed)
    npd = NP_DT[dt]
    if np.issubdtype(npd, np.integer):
        return rng.integers(1, 5, size=n).astype(npd)
    if np.issubdtype(npd, np.complexfloating):
        return (rng.random(n) + 1j * rng.random(n)).astype(npd)
    return (rng.random(n) + 0.5).astype(npd)


def np_reduce(op, arrs):
    a = np.stack(arrs)
    if op == "sum":
        return a.sum(0, dtype=a.dtype)
    if op == "prod":
        return a.prod(0, dtype=a.dtype)
    if op == "max":
        return a.max(0)
    if op == "min":
        return a.min(0)
    if op == "avg":
        return (a.sum(0) / len(arrs)).astype(a.dtype)
    if op == "band":
        return np.bitwise_and.reduce(a, 0)
    if op == "bor":
        return np.bitwise_or.reduce(a, 0)
    if op == "bxor":
        return np.bitwise_xor.reduce(a, 0)
    if op == "land":
        return np.logical_and.reduce(a != 0, 0).astype(a.dtype)
    if op == "lor":
        return np.logical_or.reduce(a != 0, 0).astype(a.dtype)
    if op == "lxor":
        return np.logical_xor.reduce(a != 0, 0).astype(a.dtype)
    raise ValueError(op)


def close(a, b, dt):
    if dt in ("float16",):
        return np.allclose(a.astype(np.float32), b.astype(np.float32), rtol=2e-2, atol=1e-2)
    if "float" in dt:
        return np.allclose(a, b, rtol=1e-4, atol=1e-5)
    return np.array_equal(a, b)


def run(team, args):
    req = team.coll(args)
    st = req.run()
    req.finalize()
    assert st == U.UCC_OK


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("count", COUNTS)
@pytest.mark.parametrize("inplace", [False, True])
def test_allreduce_sum(teams, n, count, inplace):
    team = teams[n]
    src = [rnd("float32", count, r) for r in range(n)]
    exp = np_reduce("sum", src)
    if inplace:
        dst = [s.copy() for s in src]
        run(team, [coll_args("allreduce", None, dst[r], inplace=True) for r in range(n)])
    else:
        dst = [np.zeros(count, np.float32) for _ in range(n)]
        run(team, [coll_args("allreduce", src[r], dst[r]) for r in range(n)])
    for r in range(n):
        assert close(dst[r], exp, "float32"), r


@pytest.mark.parametrize("dt,op", [(d, o) for d in ["int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64"]
                                   for o in ["sum", "prod", "max", "min", "land", "lor", "lxor", "band", "bor", "bxor"]] +
                         [(d, o) for d in ["float32", "float64", "float16"] for o in ["sum", "prod", "max", "min", "avg"]] +
                         [(d, o) for d in ["float32_complex", "float64_complex"] for o in ["sum", "prod", "avg"]])
def test_allreduce_dt_op(teams, dt, op):
    n, count = 4, 33
    team = teams[n]
    src = [rnd(dt, count, r + 7) for r in range(n)]
    dst = [np.zeros(count, NP_DT[dt]) for _ in range(n)]
    run(team, [coll_args("allreduce", src[r], dst[r], dt=dt, op=op) for r in range(n)])
    exp = np_reduce(op, src)
    for r in range(n):
        assert close(dst[r], exp, dt), (dt, op, dst[r], exp)


def test_allreduce_bfloat16(teams):
    import torch
    n, count = 4, 64
    team = teams[n]
    src_t = [(torch.rand(count) + 0.5).to(torch.bfloat16) for _ in range(n)]
    src = [t.view(torch.int16).numpy().copy() for t in src_t]
    dst = [np.zeros(count, np.int16) for _ in range(n)]
    run(team, [coll_args("allreduce", src[r], dst[r], dt="bfloat16") for r in range(n)])
    exp = sum(t.float() for t in src_t)
    for r in range(n):
        got = torch.from_numpy(dst[r]).view(torch.bfloat16).float()
        assert torch.allclose(got, exp, rtol=3e-2, atol=3e-2)


def test_allreduce_persistent(teams):
    n, count = 4, 1000
    team = teams[n]
    src = [rnd("float32", count, r) for r in range(n)]
    dst = [np.zeros(count, np.float32) for _ in range(n)]
    req = team.coll([coll_args("allreduce", src[r], dst[r], persistent=True) for r in range(n)])
    for it in range(3):
        for r in range(n):
            src[r][:] = rnd("float32", count, 100 * it + r)
        exp = np_reduce("sum", src)
        assert req.run() == U.UCC_OK
        for r in range(n):
            assert close(dst[r], exp, "float32")
    req.finalize()


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("count", [1, 64, 3001])
@pytest.mark.parametrize("inplace", [False, True])
def test_allgather(teams, n, count, inplace):
    team = teams[n]
    src = [rnd("int32", count, r) for r in range(n)]
    exp = np.concatenate(src)
    dst = [np.zeros(count * n, np.int32) for _ in range(n)]
    if inplace:
        for r in range(n):
            dst[r][r * count:(r + 1) * count] = src[r]
        run(team, [coll_args("allgather", None, dst[r], dt="int32", inplace=True) for r in range(n)])
    else:
        run(team, [coll_args("allgather", src[r], dst[r], dt="int32") for r in range(n)])
    for r in range(n):
        assert np.array_equal(dst[r], exp)


@pytest.mark.parametrize("n", SIZES)
def test_allgatherv(teams, n):
    team = teams[n]
    counts = [3 + 5 * r for r in range(n)]
    displs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    src = [rnd("float64", counts[r], r) for r in range(n)]
    exp = np.concatenate(src)
    dst = [np.zeros(sum(counts), np.float64) for _ in range(n)]
    run(team, [coll_args("allgatherv", src[r], dst[r], dt="float64", dst_counts=counts, dst_displs=displs) for r in range(n)])
    for r in range(n):
        assert np.array_equal(dst[r], exp)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("count", [1, 17, 2048])
@pytest.mark.parametrize("inplace", [False, True])
def test_alltoall(teams, n, count, inplace):
    team = teams[n]
    src = [rnd("int64", count * n, r) for r in range(n)]
    exp = [np.concatenate([src[p][r * count:(r + 1) * count] for p in range(n)]) for r in range(n)]
    if inplace:
        dst = [s.copy() for s in src]
        run(team, [coll_args("alltoall", None, dst[r], dt="int64", inplace=True) for r in range(n)])
    else:
        dst = [np.zeros(count * n, np.int64) for _ in range(n)]
        run(team, [coll_args("alltoall", src[r], dst[r], dt="int64") for r in range(n)])
    for r in range(n):
        assert np.array_equal(dst[r], exp[r])


@pytest.mark.parametrize("n", SIZES)
def test_alltoallv(teams, n):
    team = teams[n]
    # rank r sends (r + p + 1) elements to peer p
    scounts = [[r + p + 1 for p in range(n)] for r in range(n)]
    rcounts = [[p + r + 1 for p in range(n)] for r in range(n)]
    sdis = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in scounts]
    rdis = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in rcounts]
    src = [rnd("float32", sum(scounts[r]), r) for r in range(n)]
    dst = [np.zeros(sum(rcounts[r]), np.float32) for r in range(n)]
    run(team, [coll_args("alltoallv", src[r], dst[r], src_counts=scounts[r], src_displs=sdis[r], dst_counts=rcounts[r],
                         dst_displs=rdis[r]) for r in range(n)])
    for r in range(n):
        exp = np.concatenate([src[p][sdis[p][r]:sdis[p][r] + scounts[p][r]] for p in range(n)])
        assert np.array_equal(dst[r], exp)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("count", [1, 100, 70000])
def test_bcast(teams, n, count):
    team = teams[n]
    for root in {0, n - 1, n // 2}:
        bufs = [rnd("float32", count, r) if r == root else np.zeros(count, np.float32) for r in range(n)]
        exp = bufs[root].copy()
        run(team, [coll_args("bcast", bufs[r], None, root=root) for r in range(n)])
        for r in range(n):
            assert np.array_equal(bufs[r], exp)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("count", [1, 100, 40000])
@pytest.mark.parametrize("op", ["sum", "max", "avg"])
def test_reduce(teams, n, count, op):
    team = teams[n]
    for root in {0, n - 1}:
        src = [rnd("float64", count, r) for r in range(n)]
        dst = [np.zeros(count, np.float64) for _ in range(n)]
        run(team, [coll_args("reduce", src[r], dst[r] if r == root else None, dt="float64", op=op, root=root,
                             count_dst=count) for r in range(n)])
        assert close(dst[root], np_reduce(op, src), "float64")


def test_reduce_inplace(teams):
    n, count, root = 4, 50, 2
    team = teams[n]
    src = [rnd("int32", count, r) for r in range(n)]
    bufs = [s.copy() for s in src]
    run(team, [coll_args("reduce", None if r == root else bufs[r], bufs[r] if r == root else None, dt="int32", root=root,
                         inplace=(r == root), count_src=count, count_dst=count) for r in range(n)])
    assert np.array_equal(bufs[root], np_reduce("sum", src))


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("inplace", [False, True])
def test_reduce_scatter(teams, n, inplace):
    team = teams[n]
    blk = 37
    src = [rnd("float32", blk * n, r) for r in range(n)]
    exp = np_reduce("sum", src)
    if inplace:
        dst = [s.copy() for s in src]
        run(team, [coll_args("reduce_scatter", None, dst[r], inplace=True) for r in range(n)])
        for r in range(n):
            assert close(dst[r][r * blk:(r + 1) * blk], exp[r * blk:(r + 1) * blk], "float32")
    else:
        dst = [np.zeros(blk, np.float32) for _ in range(n)]
        run(team, [coll_args("reduce_scatter", src[r], dst[r]) for r in range(n)])
        for r in range(n):
            assert close(dst[r], exp[r * blk:(r + 1) * blk], "float32")


@pytest.mark.parametrize("n", SIZES)
def test_reduce_scatterv(teams, n):
    team = teams[n]
    counts = [2 + 3 * r for r in range(n)]
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    src = [rnd("int32", sum(counts), r) for r in range(n)]
    exp = np_reduce("sum", src)
    dst = [np.zeros(counts[r], np.int32) for r in range(n)]
    run(team, [coll_args("reduce_scatterv", src[r], dst[r], dt="int32", dst_counts=counts, dst_displs=offs) for r in range(n)])
    for r in range(n):
        assert np.array_equal(dst[r], exp[offs[r]:offs[r] + counts[r]])


@pytest.mark.parametrize("n", SIZES)
def test_barrier_fanin_fanout(teams, n):
    team = teams[n]
    run(team, [coll_args("barrier") for _ in range(n)])
    run(team, [coll_args("fanin", root=n - 1) for _ in range(n)])
    run(team, [coll_args("fanout", root=0) for _ in range(n)])


@pytest.mark.parametrize("n", SIZES)
def test_gather_scatter(teams, n):
    team = teams[n]
    blk = 11
    for root in {0, n - 1}:
        src = [rnd("int32", blk, r) for r in range(n)]
        dst = [np.zeros(blk * n, np.int32) for _ in range(n)]
        run(team, [coll_args("gather", src[r], dst[r] if r == root else None, dt="int32", root=root, count_dst=blk * n) for r in range(n)])
        assert np.array_equal(dst[root], np.concatenate(src))
        big = rnd("int32", blk * n, 99)
        out = [np.zeros(blk, np.int32) for _ in range(n)]
        run(team, [coll_args("scatter", big if r == root else None, out[r], dt="int32", root=root, count_src=blk * n) for r in range(n)])
        for r in range(n):
            assert np.array_equal(out[r], big[r * blk:(r + 1) * blk])


@pytest.mark.parametrize("n", SIZES)
def test_gatherv_scatterv(teams, n):
    team = teams[n]
    counts = [1 + 2 * r for r in range(n)]
    displs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    root = n - 1
    src = [rnd("float32", counts[r], r) for r in range(n)]
    dst = np.zeros(sum(counts), np.float32)
    run(team, [coll_args("gatherv", src[r], dst if r == root else None, root=root,
                         dst_counts=counts if r == root else None, dst_displs=displs if r == root else None) for r in range(n)])
    assert np.array_equal(dst, np.concatenate(src))
    big = rnd("float32", sum(counts), 5)
    out = [np.zeros(counts[r], np.float32) for r in range(n)]
    run(team, [coll_args("scatterv", big if r == root else None, out[r], root=root,
                         src_counts=counts if r == root else None, src_displs=displs if r == root else None) for r in range(n)])
    for r in range(n):
        assert np.array_equal(out[r], big[displs[r]:displs[r] + counts[r]])


def test_zero_size(teams):
    team = teams[4]
    e = np.zeros(0, np.float32)
    run(team, [coll_args("allreduce", e, e) for _ in range(4)])
    run(team, [coll_args("bcast", e, None) for _ in range(4)])
