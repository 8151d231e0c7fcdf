"""cl/hier on synthetic multi-node placements (reference: test/gtest/core/test_topo.cc style fake proc info +
cl/hier algorithm schedules, cl_hier/allreduce/allreduce_rab.c, allreduce_split_rail.c, bcast_2step.c, reduce_2step.c)."""
import numpy as np
import pytest

from ucc_b200 import capi as U
from ucc_b200.harness import UccJob, coll_args


def run(team, args):
    req = team.coll(args)
    st = req.run()
    req.finalize()
    assert st == U.UCC_OK


def job(n, ppn, tune=None):
    env = {"UCC_CLS": "hier,basic", "UCC_CL_HIER_TLS": "shm,self", "UCC_CL_BASIC_TLS": "shm,self"}
    if tune:
        env["UCC_CL_HIER_TUNE"] = tune
    return UccJob(n, ppn=ppn, env=env, cls="hier,basic")


@pytest.mark.parametrize("n,ppn", [(4, 2), (8, 4), (6, 2), (6, 3), (5, 2), (8, 1)])
@pytest.mark.parametrize("alg", ["rab", "split_rail"])
def test_hier_allreduce(n, ppn, alg):
    with job(n, ppn, tune=f"allreduce:0-inf:@{alg}") as j:
        team = j.create_team()
        rng = np.random.default_rng(n * 10 + ppn)
        for count in (1, 7, 12 * ppn, 4096):
            for inplace in (False, True):
                src = [rng.integers(-50, 50, count).astype(np.int64) for _ in range(n)]
                exp = np.sum(src, 0)
                if inplace:
                    dst = [s.copy() for s in src]
                    run(team, [coll_args("allreduce", None, dst[r], dt="int64", op="sum", inplace=True) for r in range(n)])
                else:
                    dst = [np.zeros(count, np.int64) for _ in range(n)]
                    run(team, [coll_args("allreduce", src[r], dst[r], dt="int64", op="sum") for r in range(n)])
                for r in range(n):
                    assert np.array_equal(dst[r], exp), (alg, count, inplace, r)
        # avg is not hierarchical: must fall back to cl/basic and still be right
        src = [rng.random(33) for _ in range(n)]
        dst = [np.zeros(33) for _ in range(n)]
        run(team, [coll_args("allreduce", src[r], dst[r], dt="float64", op="avg") for r in range(n)])
        for r in range(n):
            assert np.allclose(dst[r], np.mean(src, 0))


@pytest.mark.parametrize("n,ppn", [(4, 2), (8, 4), (6, 3), (7, 3)])
def test_hier_bcast_reduce_barrier(n, ppn):
    with job(n, ppn) as j:
        team = j.create_team()
        rng = np.random.default_rng(n)
        for root in range(n):   # leader roots go 2step, others fall back to cl/basic
            bufs = [rng.random(100).astype(np.float32) if r == root else np.zeros(100, np.float32) for r in range(n)]
            exp = bufs[root].copy()
            run(team, [coll_args("bcast", bufs[r], None, root=root) for r in range(n)])
            for r in range(n):
                assert np.array_equal(bufs[r], exp), (root, r)
            src = [rng.integers(0, 100, 64).astype(np.int32) for _ in range(n)]
            keep = [s.copy() for s in src]
            dst = [np.zeros(64, np.int32) for _ in range(n)]
            run(team, [coll_args("reduce", src[r], dst[r] if r == root else None, dt="int32", op="sum", root=root) for r in range(n)])
            assert np.array_equal(dst[root], np.sum(keep, 0)), root
            for r in range(n):
                assert np.array_equal(src[r], keep[r])
        for _ in range(3):
            run(team, [coll_args("barrier") for _ in range(n)])


def test_hier_single_node_falls_back():
    with job(4, 4) as j:   # one node: cl/hier refuses the team, cl/basic serves it
        team = j.create_team()
        src = [np.full(10, r + 1, np.int32) for r in range(4)]
        dst = [np.zeros(10, np.int32) for _ in range(4)]
        run(team, [coll_args("allreduce", src[r], dst[r], dt="int32", op="sum") for r in range(4)])
        assert all(np.array_equal(d, np.full(10, 10)) for d in dst)


@pytest.mark.parametrize("n,ppn", [(4, 2), (6, 3), (8, 4), (5, 2)])
def test_hier_alltoall_node_split(n, ppn):
    with job(n, ppn) as j:
        team = j.create_team()
        rng = np.random.default_rng(n)
        count = 5
        src = [rng.integers(0, 1 << 30, count * n).astype(np.int32) for _ in range(n)]
        dst = [np.zeros(count * n, np.int32) for _ in range(n)]
        run(team, [coll_args("alltoall", src[r], dst[r], dt="int32") for r in range(n)])
        for r in range(n):
            assert np.array_equal(dst[r], np.concatenate([src[p][r * count:(r + 1) * count] for p in range(n)])), r
        sc = [[(r + 2 * p) % 4 for p in range(n)] for r in range(n)]
        rc = [[sc[p][r] for p in range(n)] for r in range(n)]
        sd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in sc]
        rd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in rc]
        src = [rng.random(max(1, sum(sc[r]))).astype(np.float32) for r in range(n)]
        dst = [np.zeros(max(1, sum(rc[r])), np.float32) for r in range(n)]
        run(team, [coll_args("alltoallv", src[r], dst[r], src_counts=sc[r], src_displs=sd[r], dst_counts=rc[r], dst_displs=rd[r]) for r in range(n)])
        for r in range(n):
            exp = np.concatenate([src[p][sd[p][r]:sd[p][r] + sc[p][r]] for p in range(n)])
            assert np.array_equal(dst[r][:len(exp)], exp), r


@pytest.mark.parametrize("n,ppn", [(4, 2), (6, 3), (8, 4), (7, 3)])
@pytest.mark.parametrize("inplace", [False, True])
def test_hier_allgatherv_gab(n, ppn, inplace):
    with job(n, ppn) as j:
        team = j.create_team()
        rng = np.random.default_rng(n)
        counts = [3 + (r % 4) for r in range(n)]
        displs = np.concatenate([[0], np.cumsum(counts)[:-1]])
        src = [rng.integers(0, 100, counts[r]).astype(np.int32) for r in range(n)]
        dst = [np.zeros(sum(counts), np.int32) for _ in range(n)]
        if inplace:
            for r in range(n):
                dst[r][displs[r]:displs[r] + counts[r]] = src[r]
            run(team, [coll_args("allgatherv", None, dst[r], dt="int32", dst_counts=counts, dst_displs=displs, inplace=True) for r in range(n)])
        else:
            run(team, [coll_args("allgatherv", src[r], dst[r], dt="int32", dst_counts=counts, dst_displs=displs) for r in range(n)])
        for r in range(n):
            assert np.array_equal(dst[r], np.concatenate(src)), r


@pytest.mark.parametrize("order", ["parallel", "ordered", "sequential"])
@pytest.mark.parametrize("n,ppn", [(4, 2), (6, 3)])
def test_hier_allreduce_rab_pipelined(n, ppn, order):
    env_extra = {"UCC_CL_HIER_ALLREDUCE_RAB_PIPELINE": f"thresh=1k:fragsize=4k:nfrags=3:pdepth=2:{order}"}
    env = {"UCC_CLS": "hier,basic", "UCC_CL_HIER_TLS": "shm,self", "UCC_CL_BASIC_TLS": "shm,self", "UCC_CL_HIER_TUNE": "allreduce:0-inf:@rab"}
    env.update(env_extra)
    with UccJob(n, ppn=ppn, env=env, cls="hier,basic") as j:
        team = j.create_team()
        rng = np.random.default_rng(5)
        for count in (100, 3000, 10001):   # below threshold, 6 fragments, ragged 20 fragments
            for inplace in (False, True):
                src = [rng.integers(-50, 50, count).astype(np.int64) for _ in range(n)]
                exp = np.sum(src, 0)
                dst = [s.copy() for s in src] if inplace else [np.zeros(count, np.int64) for _ in range(n)]
                args = [coll_args("allreduce", None if inplace else src[r], dst[r], dt="int64", op="sum", inplace=inplace, persistent=True) for r in range(n)]
                req = team.coll(args)
                for rep in range(2):    # persistent: the pipeline must re-arm cleanly
                    if inplace and rep:
                        for r in range(n):
                            dst[r][:] = src[r]
                    assert req.run() == U.UCC_OK
                    for r in range(n):
                        assert np.array_equal(dst[r], exp), (count, inplace, rep, r)
                req.finalize()


@pytest.mark.parametrize("order", ["parallel", "sequential"])
@pytest.mark.parametrize("n,ppn", [(4, 2), (6, 3), (8, 4)])
def test_hier_allreduce_split_rail_pipelined(n, ppn, order):
    """ALLREDUCE_SPLIT_RAIL_PIPELINE (reference cl_hier.c:68-71): fragments are multiples of the node size, the last one is shorter"""
    env = {"UCC_CLS": "hier,basic", "UCC_CL_HIER_TLS": "shm,self", "UCC_CL_BASIC_TLS": "shm,self", "UCC_CL_HIER_TUNE": "allreduce:0-inf:@split_rail",
           "UCC_CL_HIER_ALLREDUCE_SPLIT_RAIL_PIPELINE": f"thresh=1k:fragsize=4k:nfrags=3:pdepth=2:{order}"}
    with UccJob(n, ppn=ppn, env=env, cls="hier,basic") as j:
        team = j.create_team()
        rng = np.random.default_rng(9)
        for count in (8 * ppn, 300 * ppn, 1037 * ppn):   # below the threshold, a few fragments, ragged last fragment
            for inplace in (False, True):
                src = [rng.integers(-50, 50, count).astype(np.int64) for _ in range(n)]
                exp = np.sum(src, 0)
                dst = [s.copy() for s in src] if inplace else [np.zeros(count, np.int64) for _ in range(n)]
                req = team.coll([coll_args("allreduce", None if inplace else src[r], dst[r], dt="int64", op="sum", inplace=inplace, persistent=True) for r in range(n)])
                for rep in range(2):
                    if inplace and rep:
                        for r in range(n):
                            dst[r][:] = src[r]
                    assert req.run() == U.UCC_OK
                    for r in range(n):
                        assert np.array_equal(dst[r], exp), (count, inplace, rep, r)
                req.finalize()


@pytest.mark.parametrize("order", ["parallel", "ordered", "sequential"])
@pytest.mark.parametrize("n,ppn", [(4, 2), (7, 3)])
def test_hier_bcast_reduce_2step_pipelined(n, ppn, order):
    """BCAST_2STEP_PIPELINE / REDUCE_2STEP_PIPELINE (reference cl_hier.c:72-79): leader roots run the pipelined 2step chains (non-root
    leaders keep one scratch per fragment in flight), other roots fall back to cl/basic"""
    pipe = f"thresh=1k:fragsize=2k:nfrags=3:pdepth=2:{order}"
    env = {"UCC_CLS": "hier,basic", "UCC_CL_HIER_TLS": "shm,self", "UCC_CL_BASIC_TLS": "shm,self", "UCC_CL_HIER_BCAST_2STEP_PIPELINE": pipe, "UCC_CL_HIER_REDUCE_2STEP_PIPELINE": pipe}
    with UccJob(n, ppn=ppn, env=env, cls="hier,basic") as j:
        team = j.create_team()
        rng = np.random.default_rng(n)
        for count in (50, 1500, 4099):
            for root in (0, ppn, 1):
                bufs = [rng.random(count).astype(np.float32) if r == root else np.zeros(count, np.float32) for r in range(n)]
                exp = bufs[root].copy()
                run(team, [coll_args("bcast", bufs[r], None, root=root) for r in range(n)])
                for r in range(n):
                    assert np.array_equal(bufs[r], exp), ("bcast", count, root, r)
                for inplace in (False, True):
                    src = [rng.integers(0, 100, count).astype(np.int32) for _ in range(n)]
                    keep = [s.copy() for s in src]
                    dst = [np.zeros(count, np.int32) for _ in range(n)]
                    if inplace:
                        dst[root][:] = src[root]
                    args = [coll_args("reduce", None if (inplace and r == root) else src[r], dst[r] if r == root else None, dt="int32", op="sum", root=root,
                                      inplace=inplace and r == root) for r in range(n)]
                    run(team, args)
                    assert np.array_equal(dst[root], np.sum(keep, 0)), ("reduce", count, root, inplace)
                    for r in range(n):
                        assert np.array_equal(src[r], keep[r])


@pytest.mark.parametrize("thresh", ["0", "8", "inf"])
def test_hier_alltoallv_split_node_thresh(thresh):
    """ALLTOALLV_SPLIT_NODE_THRESH (reference cl_hier.c:63-66): only node-peer blocks above the threshold go through the NODE sub-team"""
    n, ppn = 6, 3
    env = {"UCC_CLS": "hier,basic", "UCC_CL_HIER_TLS": "shm,self", "UCC_CL_BASIC_TLS": "shm,self", "UCC_CL_HIER_ALLTOALLV_SPLIT_NODE_THRESH": thresh}
    with UccJob(n, ppn=ppn, env=env, cls="hier,basic") as j:
        team = j.create_team()
        rng = np.random.default_rng(3)
        sc = [[(r + 3 * p) % 5 for p in range(n)] for r in range(n)]     # 0..4 float32 = 0..16 bytes: both sides of the 8-byte threshold
        rc = [[sc[p][r] for p in range(n)] for r in range(n)]
        sd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in sc]
        rd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in rc]
        src = [rng.random(max(1, sum(sc[r]))).astype(np.float32) for r in range(n)]
        dst = [np.zeros(max(1, sum(rc[r])), np.float32) for r in range(n)]
        run(team, [coll_args("alltoallv", src[r], dst[r], src_counts=sc[r], src_displs=sd[r], dst_counts=rc[r], dst_displs=rd[r]) for r in range(n)])
        for r in range(n):
            exp = np.concatenate([src[p][sd[p][r]:sd[p][r] + sc[p][r]] for p in range(n)])
            assert np.array_equal(dst[r][:len(exp)], exp), r
