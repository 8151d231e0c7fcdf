"""Every tl/shm algorithm forced through UCC_TL_SHM_TUNE (fresh job per algorithm, like the
reference's algorithm-forced gtest variants, test/gtest/coll/test_allreduce.cc:290-660)."""
import numpy as np
import pytest

from ucc_b200 import capi as U
from ucc_b200.harness import UccJob, coll_args

ALGS = {
    "allreduce": ["knomial", "sra_knomial", "dbt", "ring", "sliding_window"],
    "allgather": ["knomial", "ring", "neighbor", "bruck", "sparbit", "linear", "batched"],
    "allgatherv": ["ring", "knomial", "linear"],
    "alltoall": ["pairwise", "bruck", "onesided"],
    "alltoallv": ["pairwise", "hybrid", "onesided"],
    "bcast": ["knomial", "sag_knomial", "dbt"],
    "reduce": ["knomial", "dbt", "srg"],
    "reduce_scatter": ["ring", "knomial"],
    "gather": ["knomial", "linear"],
    "scatter": ["knomial", "linear"],
}
SIZES = [2, 3, 4, 5, 6, 8, 11, 16]


def run(team, args):
    req = team.coll(args)
    st = req.run()
    req.finalize()
    assert st == U.UCC_OK


def check_coll(team, coll, n, count):
    rng = np.random.default_rng(n * 1000 + count)
    if coll == "allreduce":
        for op in ("sum", "avg"):
            src = [rng.random(count).astype(np.float64) for _ in range(n)]
            dst = [np.zeros(count) for _ in range(n)]
            run(team, [coll_args(coll, src[r], dst[r], dt="float64", op=op) for r in range(n)])
            exp = np.sum(src, 0) / (n if op == "avg" else 1)
            for r in range(n):
                assert np.allclose(dst[r], exp), (r, op)
    elif coll == "allgather":
        src = [rng.integers(0, 1 << 30, count).astype(np.int32) for _ in range(n)]
        dst = [np.zeros(count * n, np.int32) for _ in range(n)]
        run(team, [coll_args(coll, src[r], dst[r], dt="int32") for r in range(n)])
        for r in range(n):
            assert np.array_equal(dst[r], np.concatenate(src)), r
    elif coll == "allgatherv":
        counts = [(count // 2) + r for r in range(n)]
        displs = np.concatenate([[0], np.cumsum(counts)[:-1]])
        src = [rng.integers(0, 100, counts[r]).astype(np.int32) for r in range(n)]
        dst = [np.zeros(sum(counts), np.int32) for _ in range(n)]
        run(team, [coll_args(coll, src[r], dst[r], dt="int32", dst_counts=counts, dst_displs=displs) for r in range(n)])
        for r in range(n):
            assert np.array_equal(dst[r], np.concatenate(src))
    elif coll == "alltoall":
        src = [rng.integers(0, 1 << 30, count * n).astype(np.int32) for _ in range(n)]
        dst = [np.zeros(count * n, np.int32) for _ in range(n)]
        run(team, [coll_args(coll, src[r], dst[r], dt="int32") for r in range(n)])
        for r in range(n):
            assert np.array_equal(dst[r], np.concatenate([src[p][r * count:(r + 1) * count] for p in range(n)])), r
    elif coll == "alltoallv":
        sc = [[(r + p) % 3 + 1 for p in range(n)] for r in range(n)]
        rc = [[(p + r) % 3 + 1 for p in range(n)] for r in range(n)]
        sd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in sc]
        rd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in rc]
        src = [rng.random(sum(sc[r])).astype(np.float32) for r in range(n)]
        dst = [np.zeros(sum(rc[r]), np.float32) for r in range(n)]
        run(team, [coll_args(coll, src[r], dst[r], src_counts=sc[r], src_displs=sd[r], dst_counts=rc[r], dst_displs=rd[r]) for r in range(n)])
        for r in range(n):
            assert np.array_equal(dst[r], np.concatenate([src[p][sd[p][r]:sd[p][r] + sc[p][r]] for p in range(n)]))
    elif coll == "bcast":
        for root in {0, n - 1, n // 2}:
            bufs = [rng.random(count).astype(np.float32) if r == root else np.zeros(count, np.float32) for r in range(n)]
            exp = bufs[root].copy()
            run(team, [coll_args(coll, bufs[r], None, root=root) for r in range(n)])
            for r in range(n):
                assert np.array_equal(bufs[r], exp), (root, r)
    elif coll == "reduce":
        for root in {0, n - 1, n // 2}:
            for op in ("sum", "avg"):
                src = [rng.random(count) for _ in range(n)]
                dst = np.zeros(count)
                run(team, [coll_args(coll, src[r], dst if r == root else None, dt="float64", op=op, root=root, count_dst=count) for r in range(n)])
                assert np.allclose(dst, np.sum(src, 0) / (n if op == "avg" else 1)), (root, op)
    elif coll == "reduce_scatter":
        src = [rng.random(count * n) for _ in range(n)]
        dst = [np.zeros(count) for _ in range(n)]
        run(team, [coll_args(coll, src[r], dst[r], dt="float64") for r in range(n)])
        exp = np.sum(src, 0)
        for r in range(n):
            assert np.allclose(dst[r], exp[r * count:(r + 1) * count]), r
    elif coll == "gather":
        for root in {0, n - 1}:
            src = [rng.integers(0, 1000, count).astype(np.int32) for _ in range(n)]
            dst = np.zeros(count * n, np.int32)
            run(team, [coll_args(coll, src[r], dst if r == root else None, dt="int32", root=root, count_dst=count * n) for r in range(n)])
            assert np.array_equal(dst, np.concatenate(src)), root
    elif coll == "scatter":
        for root in {0, n - 1}:
            big = rng.integers(0, 1000, count * n).astype(np.int32)
            out = [np.zeros(count, np.int32) for _ in range(n)]
            run(team, [coll_args(coll, big if r == root else None, out[r], dt="int32", root=root, count_src=count * n) for r in range(n)])
            for r in range(n):
                assert np.array_equal(out[r], big[r * count:(r + 1) * count]), (root, r)


@pytest.mark.parametrize("coll,alg", [(c, a) for c, algs in ALGS.items() for a in algs])
def test_forced_alg(coll, alg):
    # score inf forces the algorithm; unsupported shapes fall back to the default via the fallback chain
    with UccJob(16, env={"UCC_TL_SHM_TUNE": f"{coll}:inf:@{alg}"}) as job:
        for n in SIZES:
            team = job.create_team(range(n))
            for count in (1, 24, 5000):
                check_coll(team, coll, n, count)
            team.destroy()


def test_tl_coll_plugin_example(monkeypatch):
    """tl/shm algorithm plugin (libucc_tlcp_shm_example.so): with a score above the TL's own it takes the allreduce ranges,
    what it declines (AVG) falls back to the built-in algorithms."""
    import ctypes as C
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(os.environ.get("UCC_B200_LIB") or os.path.join(root, "ucc_b200", "lib", "libucc.so"))
    plugin = os.path.join(libdir, "ucc", "libucc_tlcp_shm_example.so")
    if not os.path.exists(plugin):
        pytest.skip("plugin not built")
    code = (
        "import ctypes as C, numpy as np\n"
        "from ucc_b200 import capi as U\n"
        "from ucc_b200.harness import UccJob, coll_args\n"
        "j = UccJob(4); t = j.create_team()\n"
        f"calls = C.c_int.in_dll(C.CDLL({plugin!r}), 'ucc_tlcp_shm_example_calls')\n"
        "for op, exp in (('sum', 10), ('max', 4), ('avg', 2.5)):\n"
        "    s = [np.full(100, r + 1, np.float64) for r in range(4)]; d = [np.zeros(100) for _ in range(4)]\n"
        "    before = calls.value\n"
        "    q = t.coll([coll_args('allreduce', s[r], d[r], dt='float64', op=op) for r in range(4)]); assert q.run() == 0; q.finalize()\n"
        "    assert all(np.allclose(x, exp) for x in d), (op, d[0][:3])\n"
        "    print(op, calls.value - before)\n")
    for score, expect in (("100", {"sum": 4, "max": 4, "avg": 0}), ("0", {"sum": 0, "max": 0, "avg": 0})):
        env = dict(os.environ, PYTHONPATH=root, UCC_TLCP_SHM_EXAMPLE_SCORE=score)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr[-2000:]
        got = dict((ln.split()[0], int(ln.split()[1])) for ln in out.stdout.splitlines() if ln and ln.split()[0] in expect)
        assert got == expect, (score, out.stdout)


@pytest.mark.parametrize("win,count", [("512K", 400_003), ("4K", 20_011), ("1", 301)])
@pytest.mark.parametrize("n", [2, 3, 5])
def test_sliding_window_multi_window_inplace(n, win, count):
    """several windows per slice: float64 slices of > 1 MB with the default 512 KB window, small vectors with ALLREDUCE_SLIDING_WIN_BUF_SIZE
    turned down (a window below one element is one element)"""
    with UccJob(n, env={"UCC_TL_SHM_TUNE": "allreduce:inf:@sliding_window", "UCC_TL_SHM_ALLREDUCE_SLIDING_WIN_BUF_SIZE": win}) as job:
        team = job.create_team()
        rng = np.random.default_rng(n)
        for inplace in (False, True):
            src = [rng.random(count) for _ in range(n)]
            exp = np.sum(src, 0)
            dst = [s.copy() for s in src] if inplace else [np.zeros(count) for _ in range(n)]
            run(team, [coll_args("allreduce", None if inplace else src[r], dst[r], dt="float64", op="sum", inplace=inplace) for r in range(n)])
            for r in range(n):
                assert np.allclose(dst[r], exp), (inplace, r)


@pytest.mark.parametrize("n", [3, 4, 5, 8, 11])
def test_alltoallv_hybrid_mixed_sizes(n):
    """hybrid alltoallv: messages below ALLTOALLV_HYBRID_THRESH ride Bruck's rounds in fixed slots, bigger ones go pairwise; sizes
    on both sides of the threshold, empty messages, and a threshold that is not a multiple of the element size"""
    for thresh in ("256", "100"):
        env = {"UCC_TL_SHM_TUNE": "alltoallv:@hybrid", "UCC_TL_SHM_ALLTOALLV_HYBRID_THRESH": thresh, "UCC_TLS": "shm,self"}
        with UccJob(n, env=env) as j:
            team = j.create_team()
            rng = np.random.default_rng(n)

            def cnt(s, d):       # elements s sends to d: 0, a few, around the threshold, far above
                return [0, 3, 24, 25, 26, 64, 65, 500][(3 * s + 5 * d + s * d) % 8]
            sc = [[cnt(r, p) for p in range(n)] for r in range(n)]
            rc = [[cnt(p, r) for p in range(n)] for r in range(n)]
            sd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in sc]
            rd = [np.concatenate([[0], np.cumsum(c)[:-1]]) for c in rc]
            src = [rng.random(max(1, sum(sc[r]))).astype(np.float32) for r in range(n)]
            dst = [np.full(max(1, sum(rc[r])), -1, np.float32) for r in range(n)]
            run(team, [coll_args("alltoallv", src[r], dst[r], src_counts=sc[r], src_displs=sd[r], dst_counts=rc[r], dst_displs=rd[r]) for r in range(n)])
            for r in range(n):
                exp = np.concatenate([src[p][sd[p][r]:sd[p][r] + sc[p][r]] for p in range(n)])
                assert np.array_equal(dst[r][:len(exp)], exp), (n, thresh, r)


@pytest.mark.parametrize("n", [2, 3, 5, 6, 7, 8, 13])
def test_allgatherv_knomial_shuffled_displacements(n):
    """recursive-doubling allgatherv with extra ranks (non power-of-two teams) and blocks that are NOT laid out in rank order"""
    with UccJob(n, env={"UCC_TL_SHM_TUNE": "allgatherv:@knomial", "UCC_TLS": "shm,self"}) as j:
        team = j.create_team()
        rng = np.random.default_rng(7 * n)
        counts = [int(c) for c in rng.integers(0, 300, n)]
        order = list(rng.permutation(n))
        displs = [0] * n
        o = 0
        for r in order:                      # block of rank `r` sits where the permutation puts it, with gaps
            displs[r] = o
            o += counts[r] + 3
        src = [rng.integers(0, 1000, counts[r]).astype(np.int32) for r in range(n)]
        dst = [np.full(o, -7, np.int32) for _ in range(n)]
        run(team, [coll_args("allgatherv", src[r], dst[r], dt="int32", dst_counts=counts, dst_displs=displs) for r in range(n)])
        for r in range(n):
            for p in range(n):
                assert np.array_equal(dst[r][displs[p]:displs[p] + counts[p]], src[p]), (n, r, p)


def _mem_map_fns():
    import ctypes as C
    U.lib.ucc_mem_map.argtypes = [U.handle, C.c_int, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    U.lib.ucc_mem_map.restype = C.c_int
    U.lib.ucc_mem_unmap.argtypes = [C.POINTER(C.c_void_p)]
    U.lib.ucc_mem_unmap.restype = C.c_int


@pytest.mark.parametrize("n", [2, 4, 7])
def test_alltoall_onesided_registered_destinations(n, capfd):
    """ucc_mem_map on HOST buffers (tl/shm mem_map / memh_pack, the role of reference tl_ucp_context.c:506-577) + alltoall `onesided`
    with UCC_COLL_ARGS_FLAG_DST_MEMH_GLOBAL: the blocks are PUT straight into the members' registered destinations (the shape of
    reference alltoall_onesided.c), at the same offset inside every member's segment; without the handles the same algorithm takes
    the get-based variant.  The trace shows which one ran (the put variant sends 1-byte tokens only, no 8-byte addresses)."""
    import ctypes as C
    _mem_map_fns()
    count, pad = 257, 64
    with UccJob(n, env={"UCC_TLS": "shm,self", "UCC_TL_SHM_TUNE": "alltoall:@onesided:inf", "UCC_TL_SHM_LOG_LEVEL": "trace"}) as j:
        team = j.create_team(range(n))
        segs = [np.zeros(pad + count * n + 5, np.int32) for _ in range(n)]     # registered segment; the collective's dst starts at `pad`
        blobs, keep = [], []
        for r in range(n):
            seg = U.ucc_mem_map_t(segs[r].ctypes.data, segs[r].nbytes)
            params = U.ucc_mem_map_params_t()
            params.segments, params.n_segments = C.pointer(seg), 1
            memh, size = C.c_void_p(), C.c_size_t()
            assert U.lib.ucc_mem_map(j.procs[r].ctx, 0, C.byref(params), C.byref(size), C.byref(memh)) == U.UCC_OK
            blob = C.string_at(memh.value, size.value)
            assert b"shm" in blob                                             # the host TL contributed a record
            blobs.append(blob); keep.append(memh)
        glob = []
        for r in range(n):
            arr = (C.c_void_p * n)()
            for p in range(n):
                b = C.create_string_buffer(blobs[p], len(blobs[p]))
                keep.append(b)
                h = C.c_void_p(C.addressof(b))
                assert U.lib.ucc_mem_map(j.procs[r].ctx, 1, None, None, C.byref(h)) == U.UCC_OK
                arr[p] = h.value
            glob.append(arr)
        rng = np.random.default_rng(n)
        for registered in (True, False, True):
            src = [rng.integers(0, 1 << 30, count * n).astype(np.int32) for _ in range(n)]
            for s in segs:
                s[:] = -1
            args = []
            for r in range(n):
                a = coll_args("alltoall", src[r], segs[r][pad:pad + count * n], dt="int32")
                if registered:
                    a.mask |= U.UCC_COLL_ARGS_FIELD_MEM_MAP_DST_MEMH | U.UCC_COLL_ARGS_FIELD_FLAGS
                    a.flags |= U.UCC_COLL_ARGS_FLAG_DST_MEMH_GLOBAL
                    a.dst_memh.global_memh = C.cast(glob[r], C.POINTER(C.c_void_p))
                args.append(a)
            capfd.readouterr()
            run(team, args)
            log = capfd.readouterr()
            log = log.out + log.err
            for r in range(n):
                assert np.array_equal(segs[r][pad:pad + count * n], np.concatenate([src[p][r * count:(r + 1) * count] for p in range(n)])), (registered, r)
                assert np.all(segs[r][:pad] == -1) and np.all(segs[r][pad + count * n:] == -1)
            assert ("len 8" in log) == (not registered), "the put variant must not exchange addresses"
        for m in keep[:n]:
            assert U.lib.ucc_mem_unmap(C.byref(m)) == U.UCC_OK


def _random_program(teams, rng, tune, steps=50):
    """`steps` random collectives on randomly chosen teams of `teams`, up to three outstanding, each checked against numpy"""
    kinds = ["allreduce", "allgather", "alltoall", "reduce_scatter", "bcast", "reduce", "barrier", "gather", "scatter", "allgatherv", "alltoallv",
             "gatherv", "scatterv", "reduce_scatterv", "fanin", "fanout"]
    window = []

    def retire(k):
        while len(window) > k:
            q, check, what = window.pop(0)
            st = q.wait()
            assert st == U.UCC_OK, (tune, what, U.status_str(st))
            q.finalize()
            check()
    for step in range(steps):
        team = teams[int(rng.integers(0, 2)) % len(teams)]
        n = len(team.members)
        kind = kinds[int(rng.integers(0, len(kinds)))]
        blk = int(rng.choice([1, 2, 17, 256, 4099, 50001]))
        dt = ["float64", "int32", "float32"][int(rng.integers(0, 3))]
        npdt = np.dtype(dt)
        op = ["sum", "max", "min"][int(rng.integers(0, 3))]
        root = int(rng.integers(0, n))
        what = (step, kind, n, blk, dt, op, root)
        mk = lambda c: rng.integers(0, 50, c).astype(npdt)                                           # noqa: E731
        red = {"sum": lambda a: np.sum(a, 0), "max": lambda a: np.max(a, 0), "min": lambda a: np.min(a, 0)}[op]
        eq = lambda got, exp, what=what: np.testing.assert_allclose(got, exp, rtol=1e-6, err_msg=str((tune, what)))   # noqa: E731
        if rng.integers(0, 8) == 0 and n >= 3:
            # active-set broadcasts (the send / recv and sub-group bcast shape): up to three with different tags in flight, random
            # start / stride / size / root; only the members take part (reference test/gtest/active_set/test_active_set.cc)
            import ctypes as C
            retire(0)
            sets = []
            for k in range(int(rng.integers(1, 4))):
                size = int(rng.integers(2, n + 1))
                stride = int(rng.choice([1, 1, 2, -1])) if size * 2 <= n + 1 else int(rng.choice([1, -1]))
                span = (size - 1) * abs(stride)
                if span >= n:
                    stride, span = (1 if stride > 0 else -1), size - 1
                start = int(rng.integers(0, n - span)) + (span if stride < 0 else 0)
                members = [start + i * stride for i in range(size)]
                root = members[int(rng.integers(0, size))]
                cnt = int(rng.choice([1, 33, 5000, 70001]))
                bufs = {r: (mk(cnt) if r == root else np.zeros(cnt, npdt)) for r in members}
                exp = bufs[root].copy()
                reqs = []
                for r in members:
                    a = coll_args("bcast", bufs[r], None, dt=dt, root=root, active_set=(start, stride, size), tag=100 + k)
                    q = C.POINTER(U.ucc_coll_req_t)()
                    assert U.ucc_collective_init(C.byref(a), C.byref(q), team.members[r].team) == U.UCC_OK, (tune, what, members)
                    reqs.append((q, a))
                sets.append((members, bufs, exp, reqs))
            for _, _, _, reqs in sets:
                for q, _ in reqs:
                    assert U.ucc_collective_post(q) == U.UCC_OK
            import time
            t0 = time.time()
            while any(q.contents.status == U.UCC_INPROGRESS for *_, reqs in sets for q, _ in reqs):
                team.job.progress()
                assert time.time() - t0 < 60, (tune, what, "active set")
            for members, bufs, exp, reqs in sets:
                for q, _ in reqs:
                    assert q.contents.status == U.UCC_OK, (tune, what, members)
                    U.ucc_collective_finalize(q)
                for r in members:
                    eq(bufs[r], exp)
            continue
        if kind in ("allreduce", "allgather", "alltoall") and rng.integers(0, 5) == 0:
            # persistent request: posted three times, new input every time
            retire(0)
            src = [mk(blk * (n if kind == "alltoall" else 1)) for _ in range(n)]
            dst = [np.zeros(blk * (1 if kind == "allreduce" else n), npdt) for _ in range(n)]
            q = team.coll([coll_args(kind, src[r], dst[r], dt=dt, op=op, persistent=True) for r in range(n)])
            for rep in range(3):
                for r in range(n):
                    src[r][:] = mk(src[r].size)
                    dst[r][:] = 0
                q.post()
                assert q.wait() == U.UCC_OK, (tune, what, "persistent", rep)
                for r in range(n):
                    if kind == "allreduce":
                        eq(dst[r], red(src))
                    elif kind == "allgather":
                        eq(dst[r], np.concatenate(src))
                    else:
                        eq(dst[r], np.concatenate([src[p_][r * blk:(r + 1) * blk] for p_ in range(n)]))
            q.finalize()
            continue
        inplace = kind in ("allreduce", "allgather", "reduce_scatter") and rng.integers(0, 3) == 0
        what = what + (("inplace",) if inplace else ())
        if kind == "allreduce":
            src = [mk(blk) for _ in range(n)]; dst = [s_.copy() if inplace else np.zeros(blk, npdt) for s_ in src]
            args = [coll_args(kind, None if inplace else src[r], dst[r], dt=dt, op=op, inplace=inplace) for r in range(n)]
            exp = red(src)
            check = lambda dst=dst, exp=exp, eq=eq: [eq(d, exp) for d in dst]                        # noqa: E731
        elif kind == "allgather":
            src = [mk(blk) for _ in range(n)]; dst = [np.zeros(blk * n, npdt) for _ in range(n)]
            if inplace:
                for r in range(n):
                    dst[r][r * blk:(r + 1) * blk] = src[r]
            args = [coll_args(kind, None if inplace else src[r], dst[r], dt=dt, inplace=inplace) for r in range(n)]
            exp = np.concatenate(src)
            check = lambda dst=dst, exp=exp, eq=eq: [eq(d, exp) for d in dst]                        # noqa: E731
        elif kind == "allgatherv":
            counts = [int(rng.integers(0, blk + 1)) for _ in range(n)]
            displs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int)
            src = [mk(counts[r]) if counts[r] else np.zeros(1, npdt)[:0] for r in range(n)]
            dst = [np.zeros(max(sum(counts), 1), npdt) for _ in range(n)]
            args = [coll_args(kind, src[r] if counts[r] else None, dst[r], dt=dt, count_src=counts[r], dst_counts=counts, dst_displs=displs) for r in range(n)]
            exp = np.concatenate(src) if sum(counts) else np.zeros(0, npdt)
            check = lambda dst=dst, exp=exp, eq=eq, tot=sum(counts): [eq(d[:tot], exp) for d in dst]  # noqa: E731
        elif kind == "alltoall":
            src = [mk(blk * n) for _ in range(n)]; dst = [np.zeros(blk * n, npdt) for _ in range(n)]
            args = [coll_args(kind, src[r], dst[r], dt=dt) for r in range(n)]
            check = lambda src=src, dst=dst, n=n, blk=blk, eq=eq: [eq(dst[r], np.concatenate([src[p][r * blk:(r + 1) * blk] for p in range(n)])) for r in range(n)]  # noqa: E731
        elif kind == "alltoallv":
            m = rng.integers(0, min(blk, 3000) + 1, (n, n))
            sd = [np.concatenate([[0], np.cumsum(m[r])[:-1]]).astype(int) for r in range(n)]
            rd = [np.concatenate([[0], np.cumsum(m[:, r])[:-1]]).astype(int) for r in range(n)]
            src = [mk(max(int(m[r].sum()), 1)) for r in range(n)]
            dst = [np.zeros(max(int(m[:, r].sum()), 1), npdt) for r in range(n)]
            args = [coll_args(kind, src[r], dst[r], dt=dt, src_counts=m[r], src_displs=sd[r], dst_counts=m[:, r], dst_displs=rd[r]) for r in range(n)]
            check = lambda src=src, dst=dst, m=m, sd=sd, n=n, eq=eq: [eq(dst[r][:int(m[:, r].sum())], np.concatenate([src[p][sd[p][r]:sd[p][r] + m[p][r]] for p in range(n)])) for r in range(n)]  # noqa: E731
        elif kind == "reduce_scatter":
            src = [mk(blk * n) for _ in range(n)]; dst = [np.zeros(blk, npdt) for _ in range(n)]
            exp = red(src)
            if inplace:   # the total vector is in dst, the result block stays at its offset
                dst = [s_.copy() for s_ in src]
                args = [coll_args(kind, None, dst[r], dt=dt, op=op, inplace=True) for r in range(n)]
                check = lambda dst=dst, exp=exp, blk=blk, n=n, eq=eq: [eq(dst[r][r * blk:(r + 1) * blk], exp[r * blk:(r + 1) * blk]) for r in range(n)]  # noqa: E731
            else:
                args = [coll_args(kind, src[r], dst[r], dt=dt, op=op) for r in range(n)]
                check = lambda dst=dst, exp=exp, blk=blk, n=n, eq=eq: [eq(dst[r], exp[r * blk:(r + 1) * blk]) for r in range(n)]  # noqa: E731
        elif kind == "bcast":
            b = [mk(blk) if r == root else np.zeros(blk, npdt) for r in range(n)]
            exp = b[root].copy()
            args = [coll_args(kind, b[r], None, dt=dt, root=root) for r in range(n)]
            check = lambda b=b, exp=exp, eq=eq: [eq(x, exp) for x in b]                              # noqa: E731
        elif kind == "reduce":
            src = [mk(blk) for _ in range(n)]; out = np.zeros(blk, npdt)
            args = [coll_args(kind, src[r], out if r == root else None, dt=dt, op=op, root=root, count_dst=blk) for r in range(n)]
            exp = red(src)
            check = lambda out=out, exp=exp, eq=eq: eq(out, exp)                                     # noqa: E731
        elif kind == "gather":
            src = [mk(blk) for _ in range(n)]; g = np.zeros(blk * n, npdt)
            args = [coll_args(kind, src[r], g if r == root else None, dt=dt, root=root, count_dst=blk * n) for r in range(n)]
            exp = np.concatenate(src)
            check = lambda g=g, exp=exp, eq=eq: eq(g, exp)                                           # noqa: E731
        elif kind == "scatter":
            big = mk(blk * n); outs = [np.zeros(blk, npdt) for _ in range(n)]
            args = [coll_args(kind, big if r == root else None, outs[r], dt=dt, root=root, count_src=blk * n) for r in range(n)]
            check = lambda big=big, outs=outs, blk=blk, n=n, eq=eq: [eq(outs[r], big[r * blk:(r + 1) * blk]) for r in range(n)]  # noqa: E731
        elif kind in ("gatherv", "scatterv"):
            counts = [int(rng.integers(0, blk + 1)) for _ in range(n)]
            displs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int)
            tot = max(sum(counts), 1)
            if kind == "gatherv":
                src = [mk(max(counts[r], 1)) for r in range(n)]; g = np.zeros(tot, npdt)
                args = [coll_args(kind, src[r], g if r == root else None, dt=dt, root=root, count_src=counts[r], dst_counts=counts, dst_displs=displs) for r in range(n)]
                exp = np.concatenate([src[r][:counts[r]] for r in range(n)]) if sum(counts) else np.zeros(0, npdt)
                check = lambda g=g, exp=exp, eq=eq, t_=sum(counts): eq(g[:t_], exp)                       # noqa: E731
            else:
                big = mk(tot); outs = [np.zeros(max(counts[r], 1), npdt) for r in range(n)]
                args = [coll_args(kind, big if r == root else None, outs[r], dt=dt, root=root, count_dst=counts[r], src_counts=counts, src_displs=displs) for r in range(n)]
                check = lambda big=big, outs=outs, counts=counts, displs=displs, n=n, eq=eq: [eq(outs[r][:counts[r]], big[displs[r]:displs[r] + counts[r]]) for r in range(n)]  # noqa: E731
        elif kind == "reduce_scatterv":
            counts = [int(rng.integers(0, blk + 1)) for _ in range(n)]
            displs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int)
            tot = max(sum(counts), 1)
            src = [mk(tot) for _ in range(n)]; dst = [np.zeros(max(counts[r], 1), npdt) for r in range(n)]
            args = [coll_args(kind, src[r], dst[r], dt=dt, op=op, count_src=sum(counts), dst_counts=counts, dst_displs=displs) for r in range(n)]
            exp = red(src)
            check = lambda dst=dst, exp=exp, counts=counts, displs=displs, n=n, eq=eq: [eq(dst[r][:counts[r]], exp[displs[r]:displs[r] + counts[r]]) for r in range(n)]  # noqa: E731
        elif kind in ("fanin", "fanout"):
            args = [coll_args(kind, root=root) for _ in range(n)]
            check = lambda: None                                                                     # noqa: E731
        else:
            args = [coll_args("barrier") for _ in range(n)]
            check = lambda: None                                                                     # noqa: E731
        q = team.coll(args)
        q.post()
        window.append((q, check, what))
        retire(int(rng.integers(0, 3)))
    retire(0)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("B200_FUZZ_SEEDS", "10"))))
def test_random_programs_with_random_algorithms(seed):
    """Fuzz of the host transport: a random algorithm is forced for every collective type (UCC_TL_SHM_TUNE, unsupported shapes fall
    back through the score chain), then 50 collectives of random kind / count / datatype / operator / root run on a team of random
    size and on a sub-team of it, up to three outstanding at a time (same post order on every member), each checked against numpy."""
    rng = np.random.default_rng(4000 + seed)
    tune = "#".join(f"{c}:inf:@{algs[int(rng.integers(0, len(algs)))]}" for c, algs in ALGS.items())
    n_all = int(rng.integers(3, 10))
    with UccJob(n_all, env={"UCC_TL_SHM_TUNE": tune, "UCC_TLS": "shm,self"}) as job:
        sub = sorted(rng.choice(n_all, size=int(rng.integers(2, n_all)), replace=False).tolist())
        _random_program([job.create_team(range(n_all)), job.create_team(sub)], rng, tune)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("B200_FUZZ_SEEDS", "8"))))
def test_random_programs_on_cl_hier(seed):
    """The same random programs through the hierarchical CL on a synthetic multi-node placement (random node size, random cl/hier
    algorithm per collective: allreduce rab / split_rail, bcast / reduce 2step, alltoall(v) node_split, allgatherv gab; everything the
    hierarchy declines falls back to cl/basic)."""
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(4, 10))
    ppn = int(rng.integers(1, n))
    tune = "allreduce:0-inf:@" + ["rab", "split_rail"][int(rng.integers(0, 2))]
    env = {"UCC_CLS": "hier,basic", "UCC_CL_HIER_TLS": "shm,self", "UCC_CL_BASIC_TLS": "shm,self", "UCC_CL_HIER_TUNE": tune}
    with UccJob(n, ppn=ppn, env=env, cls="hier,basic") as job:
        _random_program([job.create_team()], rng, (tune, n, ppn), steps=40)


KNOBS = {
    "UCC_TL_SHM_KN_RADIX": ["0", "2", "3", "8"],
    "UCC_TL_SHM_FANIN_KN_RADIX": ["auto", "2", "5"],
    "UCC_TL_SHM_FANOUT_KN_RADIX": ["auto", "3"],
    "UCC_TL_SHM_SCATTER_KN_RADIX": ["auto", "2", "3"],
    "UCC_TL_SHM_BCAST_SAG_KN_RADIX": ["auto", "3", "0-1k:2,1k-inf:4"],
    "UCC_TL_SHM_ALLTOALLV_PAIRWISE_NUM_POSTS": ["auto", "1", "3"],
    "UCC_TL_SHM_ALLTOALL_PAIRWISE_NUM_POSTS": ["auto", "1", "2"],
    "UCC_TL_SHM_ALLGATHER_BATCHED_NUM_POSTS": ["auto", "0", "1", "3"],
    "UCC_TL_SHM_GATHERV_LINEAR_NUM_POSTS": ["0", "1", "2"],
    "UCC_TL_SHM_SCATTERV_LINEAR_NUM_POSTS": ["0", "1", "3"],
    "UCC_TL_SHM_REDUCE_SCATTER_RING_BIDIRECTIONAL": ["y", "n"],
    "UCC_TL_SHM_REDUCE_SCATTERV_RING_BIDIRECTIONAL": ["y", "n"],
}


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("B200_FUZZ_SEEDS", "6"))))
def test_random_programs_with_random_knobs(seed):
    """The tuning knobs of the host transport that mirror tl/ucp's (reference tl_ucp.c:55-250: radixes, outstanding-message limits,
    bidirectional rings): a random value for each, the algorithms they act on forced, then a random program checked against numpy"""
    rng = np.random.default_rng(9000 + seed)
    env = {k: v[int(rng.integers(0, len(v)))] for k, v in KNOBS.items()}
    tune = "allgather:inf:@batched#bcast:inf:@sag_knomial#scatter:inf:@knomial#gather:inf:@knomial#reduce_scatter:inf:@ring#alltoall:inf:@pairwise#alltoallv:inf:@pairwise"
    n_all = int(rng.integers(3, 10))
    with UccJob(n_all, env=dict(env, UCC_TL_SHM_TUNE=tune, UCC_TLS="shm,self")) as job:
        sub = sorted(rng.choice(n_all, size=int(rng.integers(2, n_all)), replace=False).tolist())
        _random_program([job.create_team(range(n_all)), job.create_team(sub)], rng, str(env))


@pytest.mark.parametrize("bidir", ["y", "n"])
@pytest.mark.parametrize("n", [2, 3, 4, 7])
def test_reduce_scatterv_ring_bidirectional(n, bidir):
    """two inverted rings, each carrying one half of every block (reference reduce_scatterv_ring.c); blocks of 0, 1 and odd counts"""
    with UccJob(n, env={"UCC_TL_SHM_REDUCE_SCATTERV_RING_BIDIRECTIONAL": bidir, "UCC_TL_SHM_REDUCE_SCATTER_RING_BIDIRECTIONAL": bidir, "UCC_TLS": "shm,self"}) as j:
        team = j.create_team()
        rng = np.random.default_rng(n)
        for counts in ([0, 1, 7, 2, 9, 0, 33][:n], [1] * n, [1001 + 2 * r for r in range(n)]):
            for inplace in (False, True):
                src = [rng.integers(-100, 100, sum(counts)).astype(np.int64) for _ in range(n)]
                keep = [s.copy() for s in src]
                displs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int)
                dst = [np.zeros(max(1, counts[r]), np.int64) for r in range(n)]
                if inplace:
                    args = [coll_args("reduce_scatterv", None, src[r], dt="int64", op="sum", dst_counts=counts, dst_displs=displs, inplace=True) for r in range(n)]
                else:
                    args = [coll_args("reduce_scatterv", src[r], dst[r], dt="int64", op="sum", dst_counts=counts, dst_displs=displs) for r in range(n)]
                run(team, args)
                exp = np.sum(keep, 0)
                for r in range(n):
                    got = src[r][displs[r]:displs[r] + counts[r]] if inplace else dst[r][:counts[r]]
                    assert np.array_equal(got, exp[displs[r]:displs[r] + counts[r]]), (counts, inplace, r)
                    # the ring reads the contributions where they are: a non-inplace source must come back untouched
                    assert inplace or np.array_equal(src[r], keep[r]), (counts, r)
        for count in (1, 5, 4096):
            src = [rng.random(count * n) for _ in range(n)]
            dst = [np.zeros(count) for _ in range(n)]
            keep = [s_.copy() for s_ in src]
            run(team, [coll_args("reduce_scatter", src[r], dst[r], dt="float64", op="avg") for r in range(n)])
            for r in range(n):
                assert np.allclose(dst[r], np.mean(keep, 0)[r * count:(r + 1) * count])
                assert np.array_equal(src[r], keep[r])


@pytest.mark.parametrize("reorder", ["y", "n"])
def test_ring_ranks_reordering_by_host(reorder, capfd):
    """RANKS_REORDERING (reference tl_ucp.c:246-249): on a team whose members alternate between two (synthetic) nodes the ring algorithms walk
    node 0's members first, then node 1's - two node crossings per lap instead of eight; results do not depend on the order"""
    n = 8
    env = {"UCC_TL_SHM_RANKS_REORDERING": reorder, "UCC_TLS": "shm,self", "UCC_TL_SHM_LOG_LEVEL": "debug",
           "UCC_TL_SHM_TUNE": "allgather:inf:@ring#allgatherv:inf:@ring#reduce_scatter:inf:@ring#reduce_scatterv:inf:@ring"}
    with UccJob(n, ppn=4, env=env) as j:
        capfd.readouterr()
        team = j.create_team([0, 4, 1, 5, 2, 6, 3, 7])       # team rank -> context rank: nodes 0 1 0 1 ...
        log = capfd.readouterr()
        assert ("ring order by host: 0 2 4 6 1 3 5 7" in log.out + log.err) == (reorder == "y")
        for count in (1, 37, 3000):
            check_coll(team, "allgather", n, count)
            check_coll(team, "allgatherv", n, count)
            check_coll(team, "reduce_scatter", n, count)
        rng = np.random.default_rng(1)
        counts = [5, 0, 11, 2, 8, 1, 64, 3]
        displs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(int)
        src = [rng.integers(-9, 9, sum(counts)).astype(np.int32) for _ in range(n)]
        dst = [np.zeros(max(1, counts[r]), np.int32) for r in range(n)]
        run(team, [coll_args("reduce_scatterv", src[r], dst[r], dt="int32", op="sum", dst_counts=counts, dst_displs=displs) for r in range(n)])
        for r in range(n):
            assert np.array_equal(dst[r][:counts[r]], np.sum(src, 0)[displs[r]:displs[r] + counts[r]]), r
        # a sub-team keeps rank order (the host order belongs to the whole team)
        sub = j.create_team([0, 4, 1])
        check_coll(sub, "allgather", 3, 10)


@pytest.mark.parametrize("radix", ["2", "3", "4", "7", "0-1k:2,1k-inf:5"])
def test_allgather_knomial_radix(radix):
    """ALLGATHER_KN_RADIX: recursive k-ing on every team size (full powers, one or several extras per proxy), allgather and allgatherv"""
    with UccJob(13, env={"UCC_TL_SHM_TUNE": "allgather:inf:@knomial#allgatherv:inf:@knomial", "UCC_TL_SHM_ALLGATHER_KN_RADIX": radix, "UCC_TLS": "shm,self"}) as job:
        for n in (2, 3, 4, 5, 6, 7, 8, 9, 11, 13):
            team = job.create_team(range(n))
            for count in (1, 24, 2000):
                check_coll(team, "allgather", n, count)
                check_coll(team, "allgatherv", n, count)
            team.destroy()


def test_reduce_avg_pre_op():
    """REDUCE_AVG_PRE_OP=y (reference tl_ucp.c: scale before the reduction): int8 contributions of 100 on 4 ranks overflow as a sum (400 -> -112, / 4 = -28)
    but not as a sum of pre-scaled values (4 x 25 = 100)"""
    for pre, exp in (("y", 100), ("n", -28)):
        with UccJob(4, env={"UCC_TL_SHM_TUNE": "allreduce:inf:@sra_knomial", "UCC_TL_SHM_REDUCE_AVG_PRE_OP": pre, "UCC_TLS": "shm,self"}) as job:
            team = job.create_team()
            src = [np.full(64, 100, np.int8) for _ in range(4)]
            dst = [np.zeros(64, np.int8) for _ in range(4)]
            run(team, [coll_args("allreduce", src[r], dst[r], dt="int8", op="avg") for r in range(4)])
            for r in range(4):
                assert np.all(dst[r] == exp), (pre, dst[r][:4])


@pytest.mark.parametrize("radix", ["2", "3", "4", "8", "0-4k:4,4k-inf:2"])
def test_allreduce_sra_knomial_radix(radix):
    """ALLREDUCE_SRA_KN_RADIX: scatter-reduce / allgather over the digits of the rank in base k; every team size (extras through proxies),
    counts that do not divide, in place, AVG, persistent re-posts"""
    with UccJob(13, env={"UCC_TL_SHM_TUNE": "allreduce:inf:@sra_knomial", "UCC_TL_SHM_ALLREDUCE_SRA_KN_RADIX": radix, "UCC_TLS": "shm,self"}) as job:
        rng = np.random.default_rng(5)
        for n in (2, 3, 4, 5, 7, 8, 9, 12, 13):
            team = job.create_team(range(n))
            for count in (n, 17, 1000, 4099):
                for op in ("sum", "avg"):
                    for inplace in (False, True):
                        src = [rng.integers(-1000, 1000, count).astype(np.float64) for _ in range(n)]
                        exp = np.sum(src, 0) / (n if op == "avg" else 1)
                        dst = [s.copy() for s in src] if inplace else [np.zeros(count) for _ in range(n)]
                        req = team.coll([coll_args("allreduce", None if inplace else src[r], dst[r], dt="float64", op=op, inplace=inplace, persistent=not inplace) for r in range(n)])
                        for rep in range(1 if inplace else 2):
                            assert req.run() == U.UCC_OK
                            for r in range(n):
                                assert np.allclose(dst[r], exp), (n, count, op, inplace, rep, r)
                        req.finalize()
            team.destroy()


@pytest.mark.parametrize("radix", ["2", "3", "4", "0-4k:5,4k-inf:2"])
def test_reduce_srg_knomial_radix(radix):
    """REDUCE_SRG_KN_RADIX: k-nomial scatter-reduce + k-nomial gather towards the root (ranks rotated so that the root has all-zero digits);
    every root, team sizes with extras, in place at the root, AVG, sources untouched"""
    with UccJob(11, env={"UCC_TL_SHM_TUNE": "reduce:inf:@srg", "UCC_TL_SHM_REDUCE_SRG_KN_RADIX": radix, "UCC_TLS": "shm,self"}) as job:
        rng = np.random.default_rng(6)
        for n in (2, 3, 4, 5, 7, 8, 9, 11):
            team = job.create_team(range(n))
            for count in (n, 29, 3001):
                for root in sorted({0, 1, n // 2, n - 1}):
                    for op in ("sum", "avg"):
                        for inplace in (False, True):
                            src = [rng.integers(-1000, 1000, count).astype(np.float64) for _ in range(n)]
                            keep = [s_.copy() for s_ in src]
                            dst = src[root].copy() if inplace else np.zeros(count)
                            run(team, [coll_args("reduce", None if (inplace and r == root) else src[r], dst if r == root else None, dt="float64", op=op, root=root,
                                                 count_dst=count, inplace=inplace and r == root) for r in range(n)])
                            assert np.allclose(dst, np.sum(keep, 0) / (n if op == "avg" else 1)), (n, count, root, op, inplace)
                            for r in range(n):
                                assert np.array_equal(src[r], keep[r])
            team.destroy()


@pytest.mark.parametrize("radix", ["2", "3", "4", "8"])
def test_reduce_scatter_knomial_radix(radix):
    """REDUCE_SCATTER_KN_RADIX: k-nomial scatter-reduce where the team is a power of the radix, recursive halving on other power-of-two teams, ring
    (fallback) elsewhere; in place too"""
    with UccJob(9, env={"UCC_TL_SHM_TUNE": "reduce_scatter:inf:@knomial", "UCC_TL_SHM_REDUCE_SCATTER_KN_RADIX": radix, "UCC_TLS": "shm,self"}) as job:
        rng = np.random.default_rng(8)
        for n in (2, 3, 4, 6, 8, 9):
            team = job.create_team(range(n))
            for count in (1, 7, 1025):
                for op in ("sum", "avg"):
                    check_coll(team, "reduce_scatter", n, count)
                    buf = [rng.integers(-99, 99, count * n).astype(np.float64) for _ in range(n)]
                    exp = np.sum(buf, 0) / (n if op == "avg" else 1)
                    run(team, [coll_args("reduce_scatter", None, buf[r], dt="float64", op=op, inplace=True) for r in range(n)])
                    for r in range(n):
                        assert np.allclose(buf[r][r * count:(r + 1) * count], exp[r * count:(r + 1) * count]), (n, count, op, r)
            team.destroy()


@pytest.mark.parametrize("order", ["parallel", "ordered", "sequential"])
@pytest.mark.parametrize("n", [2, 5, 8])
def test_sra_srg_pipelines(n, order, capfd):
    """ALLREDUCE_SRA_KN_PIPELINE / REDUCE_SRG_KN_PIPELINE: fragments of the vector run as re-armed SRA / SRG tasks (ragged last fragment,
    vectors below the threshold unpipelined, persistent re-posts, in place)"""
    pipe = f"thresh=2k:fragsize=4k:nfrags=4:pdepth=2:{order}"
    env = {"UCC_TL_SHM_TUNE": "allreduce:inf:@sra_knomial#reduce:inf:@srg", "UCC_TL_SHM_ALLREDUCE_SRA_KN_PIPELINE": pipe, "UCC_TL_SHM_REDUCE_SRG_KN_PIPELINE": pipe,
           "UCC_TL_SHM_ALLREDUCE_SRA_KN_RADIX": "3", "UCC_TLS": "shm,self", "UCC_TL_SHM_LOG_LEVEL": "debug"}
    with UccJob(n, env=env) as job:
        team = job.create_team()
        rng = np.random.default_rng(n)
        for count in (100, 1024, 2600, 5003):
            capfd.readouterr()
            for op in ("sum", "avg"):
                for inplace in (False, True):
                    src = [rng.integers(-1000, 1000, count).astype(np.float64) for _ in range(n)]
                    exp = np.sum(src, 0) / (n if op == "avg" else 1)
                    dst = [s_.copy() for s_ in src] if inplace else [np.zeros(count) for _ in range(n)]
                    req = team.coll([coll_args("allreduce", None if inplace else src[r], dst[r], dt="float64", op=op, inplace=inplace, persistent=True) for r in range(n)])
                    for rep in range(2):
                        if inplace and rep:
                            for r in range(n):
                                dst[r][:] = src[r]
                        assert req.run() == U.UCC_OK
                        for r in range(n):
                            assert np.allclose(dst[r], exp), ("allreduce", count, op, inplace, rep, r)
                    req.finalize()
            for root in sorted({0, n - 1}):
                src = [rng.integers(-1000, 1000, count).astype(np.float64) for _ in range(n)]
                keep = [s_.copy() for s_ in src]
                dst = np.zeros(count)
                run(team, [coll_args("reduce", src[r], dst if r == root else None, dt="float64", op="sum", root=root, count_dst=count) for r in range(n)])
                assert np.allclose(dst, np.sum(keep, 0)), ("reduce", count, root)
            log = capfd.readouterr()
            log = log.out + log.err
            nf = -(-count // 512)                                        # 4 KB fragments of float64
            assert (f"allreduce pipelined: {nf} fragments of 512 elements, 2 in flight" in log) == (count * 8 >= 2048), count
            assert (f"reduce pipelined: {nf} fragments" in log) == (count * 8 >= 2048), count
