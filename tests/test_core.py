"""Core object model (model: reference test/gtest/core/test_lib.cc, test_lib_config.cc, test_context.cc, test_team.cc,
test_timeout.cc, active_set/test_active_set.cc, asym_mem/test_asymmetric_memory.cc, core/test_mem_map.cc)."""
import ctypes as C
import os
import time

import numpy as np
import pytest

from ucc_b200 import capi as U
from ucc_b200.harness import UccJob, coll_args

libc = C.CDLL(None)
libc.open_memstream.restype = C.c_void_p
libc.open_memstream.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]
libc.fclose.argtypes = [C.c_void_p]
U.lib.ucc_lib_config_print.argtypes = [U.handle, C.c_void_p, C.c_char_p, C.c_int]
U.lib.ucc_lib_config_print.restype = None
U.lib.ucc_context_config_print.argtypes = [U.handle, C.c_void_p, C.c_char_p, C.c_int]
U.lib.ucc_context_config_print.restype = None


def run(team, args):
    req = team.coll(args)
    st = req.run()
    req.finalize()
    return st


def _print_cfg(fn, cfg, flags=1 | 2 | 4):
    buf, size = C.c_char_p(), C.c_size_t()
    f = libc.open_memstream(C.byref(buf), C.byref(size))
    fn(cfg, f, b"title", flags)
    libc.fclose(f)
    return C.string_at(buf, size.value).decode()


def test_version_and_status_strings():
    assert U.ucc_get_version_string().decode().startswith("1.")
    assert U.ucc_status_string(U.UCC_OK) == b"Success"
    assert b"progress" in U.ucc_status_string(U.UCC_INPROGRESS).lower()
    assert U.ucc_status_string(-12345) is not None


def test_lib_config_read_modify_print(monkeypatch):
    monkeypatch.setenv("UCC_CLS", "basic")
    cfg = U.handle()
    assert U.ucc_lib_config_read(None, None, C.byref(cfg)) == U.UCC_OK
    out = _print_cfg(U.lib.ucc_lib_config_print, cfg)
    assert "UCC_CLS=basic" in out and "title" in out
    assert U.ucc_lib_config_modify(cfg, b"CLS", b"all") == U.UCC_OK
    assert "UCC_CLS=all" in _print_cfg(U.lib.ucc_lib_config_print, cfg)
    assert U.ucc_lib_config_modify(cfg, b"NO_SUCH_FIELD", b"1") != U.UCC_OK
    U.ucc_lib_config_release(cfg)


def test_lib_config_env_prefix(monkeypatch):
    # "<PREFIX>_UCC_<NAME>" wins over the plain "UCC_<NAME>" (reference ucc_lib_config_read semantics)
    monkeypatch.setenv("UCC_CLS", "basic")
    monkeypatch.setenv("MYAPP_UCC_CLS", "basic,hier")
    cfg = U.handle()
    assert U.ucc_lib_config_read(b"MYAPP", None, C.byref(cfg)) == U.UCC_OK
    assert "CLS=basic,hier" in _print_cfg(U.lib.ucc_lib_config_print, cfg)
    U.ucc_lib_config_release(cfg)


def test_config_file(tmp_path):
    # the file is read once, when the library is loaded -> needs a fresh process
    import subprocess
    import sys
    f = tmp_path / "ucc.conf"
    f.write_text("# comment\nUCC_TL_SHM_TUNE = allreduce:@ring\n\n[section]\nUCC_CLS=basic\n")
    code = (
        "import ctypes as C, numpy as np\n"
        "from ucc_b200 import capi as U\n"
        "from ucc_b200.harness import UccJob, coll_args\n"
        "j = UccJob(2); team = j.create_team()\n"
        "src = [np.full(64, r + 1.0, np.float32) for r in range(2)]; dst = [np.zeros(64, np.float32) for _ in range(2)]\n"
        "req = team.coll([coll_args('allreduce', src[r], dst[r]) for r in range(2)]); assert req.run() == 0; req.finalize()\n"
        "assert np.all(dst[0] == 3)\n"
        "U.lib.ucc_context_config_print.argtypes = [U.handle, C.c_void_p, C.c_char_p, C.c_int]\n"
        "libc = C.CDLL(None); libc.fdopen.restype = C.c_void_p\n"
        "cfg = U.handle(); assert U.ucc_context_config_read(j.procs[0].lib, None, C.byref(cfg)) == 0\n"
        "fp = libc.fdopen(1, b'w'); U.lib.ucc_context_config_print(cfg, fp, b't', 1); libc.fflush(C.c_void_p(fp))\n"
        "j.cleanup()\n")
    env = dict(os.environ, UCC_CONFIG_FILE=str(f), PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env.pop("UCC_TL_SHM_TUNE", None)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "UCC_TL_SHM_TUNE=allreduce:@ring" in out.stdout


def test_lib_attr_thread_mode():
    for tm in (U.UCC_THREAD_SINGLE, U.UCC_THREAD_MULTIPLE):
        with UccJob(1, thread_mode=tm, with_ctx_oob=False) as j:
            attr = U.ucc_lib_attr_t()
            attr.mask = U.UCC_LIB_PARAM_FIELD_THREAD_MODE | (1 << 1)
            assert U.ucc_lib_get_attr(j.procs[0].lib, C.byref(attr)) == U.UCC_OK
            assert attr.thread_mode == tm
            assert attr.coll_types != 0


def test_context_attr_addr():
    with UccJob(2) as j:
        a = U.ucc_context_attr_t()
        a.mask = U.UCC_CONTEXT_ATTR_FIELD_CTX_ADDR | U.UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN | U.UCC_CONTEXT_ATTR_FIELD_WORK_BUFFER_SIZE
        assert U.ucc_context_get_attr(j.procs[0].ctx, C.byref(a)) == U.UCC_OK
        assert a.ctx_addr_len > 0 and a.ctx_addr


def test_local_context_team_with_oob_only():
    # contexts created without a context OOB: the team's OOB carries the address exchange
    with UccJob(3, with_ctx_oob=False) as j:
        team = j.create_team()
        src = [np.full(10, r, np.int32) for r in range(3)]
        dst = [np.zeros(10, np.int32) for _ in range(3)]
        assert run(team, [coll_args("allreduce", src[r], dst[r], dt="int32") for r in range(3)]) == U.UCC_OK
        assert np.all(dst[2] == 3)


def test_many_teams_subsets_and_ids():
    with UccJob(6) as j:
        teams = [j.create_team(), j.create_team([0, 2, 4]), j.create_team([1, 3, 5]), j.create_team([5, 0]), j.create_team(range(6), use_ep_map=True)]
        for t in teams:
            n = len(t.ranks)
            attr = U.ucc_team_attr_t()
            attr.mask = U.UCC_TEAM_ATTR_FIELD_SIZE | U.UCC_TEAM_ATTR_FIELD_EP
            assert U.ucc_team_get_attr(t.members[n - 1].team, C.byref(attr)) == U.UCC_OK
            assert attr.size == n and attr.ep == n - 1
            src = [np.full(33, i + 1, np.int64) for i in range(n)]
            dst = [np.zeros(33, np.int64) for _ in range(n)]
            assert run(t, [coll_args("allreduce", src[i], dst[i], dt="int64") for i in range(n)]) == U.UCC_OK
            assert all(np.all(d == n * (n + 1) // 2) for d in dst)
        # destroy / re-create: team ids are recycled without clashes
        for _ in range(5):
            t = j.create_team([1, 2, 3])
            src = [np.full(5, 1, np.int32) for _ in range(3)]
            dst = [np.zeros(5, np.int32) for _ in range(3)]
            assert run(t, [coll_args("allreduce", src[i], dst[i], dt="int32") for i in range(3)]) == U.UCC_OK
            t.destroy()
            j.teams.remove(t)


def test_interleaved_collectives_two_teams():
    with UccJob(4) as j:
        ta, tb = j.create_team(), j.create_team([0, 1, 2, 3])
        n = 4
        bufs = []
        reqs = []
        for k in range(6):
            t = ta if k % 2 == 0 else tb
            src = [np.full(1000, (k + 1) * (r + 1), np.int64) for r in range(n)]
            dst = [np.zeros(1000, np.int64) for _ in range(n)]
            bufs.append((src, dst, (k + 1) * 10))
            reqs.append(t.coll([coll_args("allreduce", src[r], dst[r], dt="int64") for r in range(n)]))
        for r in reqs:
            r.post()
        for r in reversed(reqs):
            assert r.wait() == U.UCC_OK
        for r in reqs:
            r.finalize()
        for src, dst, exp in bufs:
            assert all(np.all(d == exp) for d in dst)


def test_timeout_fires_when_a_rank_is_missing():
    with UccJob(2) as j:
        team = j.create_team()
        src = np.ones(16, np.float32)
        dst = np.zeros(16, np.float32)
        args = coll_args("allreduce", src, dst, timeout=0.2)
        req = C.POINTER(U.ucc_coll_req_t)()
        assert U.ucc_collective_init(C.byref(args), C.byref(req), team.members[0].team) == U.UCC_OK
        assert U.ucc_collective_post(req) == U.UCC_OK
        t0 = time.time()
        while req.contents.status == U.UCC_INPROGRESS and time.time() - t0 < 5:
            j.progress()
        assert req.contents.status == U.UCC_ERR_TIMED_OUT
        assert time.time() - t0 >= 0.19
        U.ucc_collective_finalize(req)
        # the late rank still has to run its side so that sequence numbers stay aligned
        args1 = coll_args("allreduce", src, np.zeros(16, np.float32), timeout=0.2)
        req1 = C.POINTER(U.ucc_coll_req_t)()
        assert U.ucc_collective_init(C.byref(args1), C.byref(req1), team.members[1].team) == U.UCC_OK
        assert U.ucc_collective_post(req1) == U.UCC_OK
        t0 = time.time()
        while req1.contents.status == U.UCC_INPROGRESS and time.time() - t0 < 5:
            j.progress()
        assert req1.contents.status in (U.UCC_OK, U.UCC_ERR_TIMED_OUT)
        U.ucc_collective_finalize(req1)


def test_active_set_bcast():
    # reference test_active_set.cc: point-to-point style bcast between a strided pair inside a bigger team
    n = 8
    with UccJob(n) as j:
        team = j.create_team()
        for start, stride, size in ((0, 1, 2), (1, 3, 2), (2, 2, 3), (7, -3, 3)):
            members = [start + i * stride for i in range(size)]
            root = members[0]
            bufs = {r: (np.arange(100, dtype=np.int32) + 7 * root if r == root else np.zeros(100, np.int32)) for r in members}
            reqs = []
            for r in members:
                a = coll_args("bcast", bufs[r], None, dt="int32", root=root, active_set=(start, stride, size), tag=5)
                q = C.POINTER(U.ucc_coll_req_t)()
                assert U.ucc_collective_init(C.byref(a), C.byref(q), team.members[r].team) == U.UCC_OK
                reqs.append((q, a))
            for q, _ in reqs:
                assert U.ucc_collective_post(q) == U.UCC_OK
            t0 = time.time()
            while any(q.contents.status == U.UCC_INPROGRESS for q, _ in reqs) and time.time() - t0 < 10:
                j.progress()
            for q, _ in reqs:
                assert q.contents.status == U.UCC_OK
                U.ucc_collective_finalize(q)
            for r in members:
                assert np.array_equal(bufs[r], np.arange(100, dtype=np.int32) + 7 * root), (start, stride, r)
        # the full team still works afterwards (sequence numbers untouched)
        src = [np.full(4, 1, np.int32) for _ in range(n)]
        dst = [np.zeros(4, np.int32) for _ in range(n)]
        assert run(team, [coll_args("allreduce", src[r], dst[r], dt="int32") for r in range(n)]) == U.UCC_OK
        assert np.all(dst[3] == n)


def test_active_set_bcast_large_keeps_team_sequence():
    """An active-set bcast whose size selects an algorithm that declines active sets (sag_knomial, the default from 32 KB) falls back to
    the knomial tree - and must leave the team's collective sequence untouched on its members: the declined attempt used to take the
    sequence number back twice, after which the members' messages of the next full-team collectives no longer matched the others'
    (found by the random-program test).  Default algorithms, 100 KB, then ring allreduce + reduce_scatter in flight together."""
    n = 6
    with UccJob(n, env={"UCC_TLS": "shm,self"}) as j:
        team = j.create_team()
        for start, stride, size, root in ((1, 2, 3, 3), (0, 1, 2, 0)):
            members = [start + i * stride for i in range(size)]
            bufs = {r: (np.arange(25000, dtype=np.int32) + root if r == root else np.zeros(25000, np.int32)) for r in members}
            reqs = []
            for r in members:
                a = coll_args("bcast", bufs[r], None, dt="int32", root=root, active_set=(start, stride, size), tag=9)
                q = C.POINTER(U.ucc_coll_req_t)()
                assert U.ucc_collective_init(C.byref(a), C.byref(q), team.members[r].team) == U.UCC_OK
                reqs.append((q, a))
            for q, _ in reqs:
                assert U.ucc_collective_post(q) == U.UCC_OK
            t0 = time.time()
            while any(q.contents.status == U.UCC_INPROGRESS for q, _ in reqs) and time.time() - t0 < 20:
                j.progress()
            for q, _ in reqs:
                assert q.contents.status == U.UCC_OK
                U.ucc_collective_finalize(q)
            for r in members:
                assert np.array_equal(bufs[r], np.arange(25000, dtype=np.int32) + root)
            # two full-team collectives in flight: their messages are told apart by the sequence number only
            s1 = [np.full(30000, r + 1.0) for r in range(n)]; d1 = [np.zeros(30000) for _ in range(n)]
            s2 = [np.full(6000 * n, r + 2.0) for r in range(n)]; d2 = [np.zeros(6000) for _ in range(n)]
            q1 = team.coll([coll_args("allreduce", s1[r], d1[r], dt="float64") for r in range(n)])
            q2 = team.coll([coll_args("reduce_scatter", s2[r], d2[r], dt="float64") for r in range(n)])
            q1.post(); q2.post()
            assert q1.wait() == U.UCC_OK and q2.wait() == U.UCC_OK
            q1.finalize(); q2.finalize()
            assert all(np.all(d == n * (n + 1) / 2) for d in d1) and all(np.all(d == n * (n + 3) / 2) for d in d2)


def test_callback_on_completion():
    with UccJob(2) as j:
        team = j.create_team()
        fired = []
        CB = U.COLL_CB_FN

        def mk(i):
            return CB(lambda data, st: fired.append((i, st)))
        cbs = [mk(i) for i in range(2)]
        src = [np.full(8, r + 1.0, np.float32) for r in range(2)]
        dst = [np.zeros(8, np.float32) for _ in range(2)]
        args = []
        for r in range(2):
            a = coll_args("allreduce", src[r], dst[r])
            a.mask |= U.UCC_COLL_ARGS_FIELD_CB
            a.cb.cb = cbs[r]
            a.cb.data = None
            args.append(a)
        req = team.coll(args)
        assert req.run() == U.UCC_OK
        req.finalize()
        assert sorted(fired) == [(0, U.UCC_OK), (1, U.UCC_OK)]


def test_init_and_post_and_errors():
    with UccJob(2) as j:
        team = j.create_team()
        # invalid: unknown collective type
        a = coll_args("allreduce", np.zeros(4, np.float32), np.zeros(4, np.float32))
        a.coll_type = 1 << 20
        q = C.POINTER(U.ucc_coll_req_t)()
        assert U.ucc_collective_init(C.byref(a), C.byref(q), team.members[0].team) < 0
        # init_and_post
        src = [np.full(8, r + 1, np.int32) for r in range(2)]
        dst = [np.zeros(8, np.int32) for _ in range(2)]
        qs = []
        for r in range(2):
            a = coll_args("allreduce", src[r], dst[r], dt="int32")
            q = C.POINTER(U.ucc_coll_req_t)()
            assert U.ucc_collective_init_and_post(C.byref(a), C.byref(q), team.members[r].team) == U.UCC_OK
            qs.append((q, a))
        while any(q.contents.status == U.UCC_INPROGRESS for q, _ in qs):
            j.progress()
        for q, _ in qs:
            assert q.contents.status == U.UCC_OK
            U.ucc_collective_finalize(q)
        assert np.all(dst[0] == 3)


def test_thread_multiple_concurrent_progress():
    import threading
    n = 4
    with UccJob(n, thread_mode=U.UCC_THREAD_MULTIPLE) as j:
        team = j.create_team()
        errs = []

        def worker(r):
            try:
                for k in range(30):
                    src = np.full(2048, r + k, np.int64)
                    dst = np.zeros(2048, np.int64)
                    a = coll_args("allreduce", src, dst, dt="int64")
                    q = C.POINTER(U.ucc_coll_req_t)()
                    U.check(U.ucc_collective_init(C.byref(a), C.byref(q), team.members[r].team), "init")
                    U.check(U.ucc_collective_post(q), "post")
                    t0 = time.time()
                    while q.contents.status == U.UCC_INPROGRESS:
                        U.ucc_context_progress(j.procs[r].ctx)
                        if time.time() - t0 > 60:
                            raise TimeoutError()
                    assert q.contents.status == U.UCC_OK
                    U.ucc_collective_finalize(q)
                    assert np.all(dst == sum(range(n)) + n * k), (r, k)
            except Exception as e:  # noqa: BLE001
                errs.append(repr(e))
        ths = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert not errs, errs


def test_thread_multiple_shared_context_stress():
    """THREAD_MULTIPLE proper: several threads share one context per rank, each drives its own team (allreduce, alltoall, bcast,
    barrier mixes) and everybody calls ucc_context_progress on the shared context concurrently."""
    import threading
    n, nteams, iters = 3, 3, 25
    with UccJob(n, thread_mode=U.UCC_THREAD_MULTIPLE) as j:
        teams = [j.create_team() for _ in range(nteams)]
        errs = []

        def run_one(r, a):
            q = C.POINTER(U.ucc_coll_req_t)()
            U.check(U.ucc_collective_init(C.byref(a), C.byref(q), a._team), "init")
            U.check(U.ucc_collective_post(q), "post")
            t0 = time.time()
            while q.contents.status == U.UCC_INPROGRESS:
                U.ucc_context_progress(j.procs[r].ctx)
                if time.time() - t0 > 90:
                    raise TimeoutError()
            assert q.contents.status == U.UCC_OK, q.contents.status
            U.ucc_collective_finalize(q)

        def worker(r, t):
            try:
                th = teams[t].members[r].team
                for k in range(iters):
                    count = 64 * (1 + (k + t) % 5) * (64 if k % 7 == 0 else 1)
                    src = np.full(count, r + k + t, np.int32)
                    dst = np.zeros(count, np.int32)
                    a = coll_args("allreduce", src, dst, dt="int32")
                    a._team = th
                    run_one(r, a)
                    assert np.all(dst == sum(range(n)) + n * (k + t)), ("allreduce", r, t, k)
                    s2 = np.repeat(np.arange(n, dtype=np.int64) + 100 * r + k, 16)
                    d2 = np.zeros(n * 16, np.int64)
                    a = coll_args("alltoall", s2, d2, dt="int64")
                    a._team = th
                    run_one(r, a)
                    exp = np.repeat(np.array([r + 100 * p + k for p in range(n)], np.int64), 16)
                    assert np.array_equal(d2, exp), ("alltoall", r, t, k)
                    b = np.full(33, k if r == k % n else -1, np.int64)
                    a = coll_args("bcast", b, None, dt="int64", root=k % n)
                    a._team = th
                    run_one(r, a)
                    assert np.all(b == k), ("bcast", r, t, k)
                    a = coll_args("barrier")
                    a._team = th
                    run_one(r, a)
            except Exception as e:  # noqa: BLE001
                errs.append(repr(e))
        ths = [threading.Thread(target=worker, args=(r, t)) for r in range(n) for t in range(nteams)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        assert not errs, errs


def test_mem_map_export_import():
    U.lib.ucc_mem_map.argtypes = [U.handle, C.c_int, C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_void_p)]
    U.lib.ucc_mem_map.restype = C.c_int
    U.lib.ucc_mem_unmap.argtypes = [C.POINTER(C.c_void_p)]
    U.lib.ucc_mem_unmap.restype = C.c_int
    with UccJob(2) as j:
        buf = np.arange(1024, dtype=np.int32)
        seg = U.ucc_mem_map_t(buf.ctypes.data, buf.nbytes)
        params = U.ucc_mem_map_params_t()
        params.segments = C.pointer(seg)
        params.n_segments = 1
        memh, size = C.c_void_p(), C.c_size_t()
        assert U.lib.ucc_mem_map(j.procs[0].ctx, 0, C.byref(params), C.byref(size), C.byref(memh)) == U.UCC_OK
        assert size.value > 0 and memh.value
        # "send" the relocatable handle to the peer and import it there
        blob = C.create_string_buffer(C.string_at(memh.value, size.value), size.value)
        imp, isz = C.c_void_p(C.addressof(blob)), C.c_size_t()
        assert U.lib.ucc_mem_map(j.procs[1].ctx, 1, None, C.byref(isz), C.byref(imp)) == U.UCC_OK
        assert isz.value == size.value
        # garbage is rejected
        junk = C.create_string_buffer(64)
        jp = C.c_void_p(C.addressof(junk))
        assert U.lib.ucc_mem_map(j.procs[1].ctx, 1, None, None, C.byref(jp)) != U.UCC_OK
        # unsupported modes
        assert U.lib.ucc_mem_map(j.procs[0].ctx, 2, C.byref(params), C.byref(size), C.byref(C.c_void_p())) == U.UCC_ERR_NOT_SUPPORTED
        assert U.lib.ucc_mem_unmap(C.byref(memh)) == U.UCC_OK


def test_service_collectives_with_subsets():
    """Internal service collectives (reference test/gtest/core/test_service_coll.cc): allreduce / allgather / bcast over the
    service team restricted to SUBSETS of a team given as ep maps (full, strided, arbitrary array)."""
    class _Map(C.Structure):            # ucc_ep_map_t with the union flattened (ctypes cannot pass unions by value)
        _fields_ = [("type", C.c_int), ("ep_num", C.c_uint64), ("a", C.c_uint64), ("b", C.c_uint64)]

    class _Subset(C.Structure):
        _fields_ = [("map", _Map), ("myrank", C.c_uint32)]
    L = U.lib
    L.ucc_service_allreduce.argtypes = [U.handle, C.c_void_p, C.c_void_p, C.c_uint64, C.c_size_t, C.c_int, _Subset, C.POINTER(C.c_void_p)]
    L.ucc_service_allgather.argtypes = [U.handle, C.c_void_p, C.c_void_p, C.c_size_t, _Subset, C.POINTER(C.c_void_p)]
    L.ucc_service_bcast.argtypes = [U.handle, C.c_void_p, C.c_size_t, C.c_uint32, _Subset, C.POINTER(C.c_void_p)]
    for f in (L.ucc_service_allreduce, L.ucc_service_allgather, L.ucc_service_bcast, L.ucc_service_coll_test, L.ucc_service_coll_finalize):
        f.restype = C.c_int
    L.ucc_service_coll_test.argtypes = [C.c_void_p]
    L.ucc_service_coll_finalize.argtypes = [C.c_void_p]
    EP_FULL, EP_STRIDED, EP_ARRAY = 1, 2, 3
    assert (U.UCC_EP_MAP_FULL, U.UCC_EP_MAP_STRIDED, U.UCC_EP_MAP_ARRAY) == (EP_FULL, EP_STRIDED, EP_ARRAY)
    n = 6
    with UccJob(n) as j:
        team = j.create_team()

        def wait_all(reqs):
            t0 = time.time()
            pending = dict(reqs)
            while pending:
                for r in list(pending):
                    st = L.ucc_service_coll_test(pending[r])
                    assert st >= 0, st
                    if st == U.UCC_OK:
                        assert L.ucc_service_coll_finalize(pending.pop(r)) == U.UCC_OK
                for p in j.procs:
                    U.ucc_context_progress(p.ctx)
                assert time.time() - t0 < 60
        arr = np.array([4, 0, 3], dtype=np.uint32)         # arbitrary order: subset rank i = team rank arr[i]
        subsets = {
            "full": (list(range(n)), lambda: _Map(EP_FULL, n, 0, 0)),
            "odd": ([1, 3, 5], lambda: _Map(EP_STRIDED, 3, 1, 2)),
            "array": ([4, 0, 3], lambda: _Map(EP_ARRAY, 3, arr.ctypes.data, 4)),
        }
        for name, (members, mk) in subsets.items():
            m = len(members)
            # allreduce MAX + SUM on int64 pairs
            src = {r: np.array([r + 1, 10 * (r + 1)], np.int64) for r in members}
            dst = {r: np.zeros(2, np.int64) for r in members}
            reqs = {}
            for i, r in enumerate(members):
                q = C.c_void_p()
                st = L.ucc_service_allreduce(team.members[r].team, src[r].ctypes.data, dst[r].ctypes.data, U.DT["int64"], 2, U.OP["sum"], _Subset(mk(), i), C.byref(q))
                assert st == U.UCC_OK, (name, st)
                reqs[r] = q
            wait_all(reqs)
            for r in members:
                assert dst[r].tolist() == [sum(x + 1 for x in members), 10 * sum(x + 1 for x in members)], (name, r, dst[r])
            # allgather of 3 bytes per member, in subset-rank order
            sb = {r: np.array([r, r + 100, r + 200], np.uint8) for r in members}
            rb = {r: np.zeros(3 * m, np.uint8) for r in members}
            reqs = {}
            for i, r in enumerate(members):
                q = C.c_void_p()
                assert L.ucc_service_allgather(team.members[r].team, sb[r].ctypes.data, rb[r].ctypes.data, 3, _Subset(mk(), i), C.byref(q)) == U.UCC_OK
                reqs[r] = q
            wait_all(reqs)
            exp = np.concatenate([sb[r] for r in members])
            for r in members:
                assert np.array_equal(rb[r], exp), (name, r)
            # bcast from subset root 1
            bb = {r: (np.arange(17, dtype=np.uint8) + 5 if i == 1 else np.zeros(17, np.uint8)) for i, r in enumerate(members)}
            reqs = {}
            for i, r in enumerate(members):
                q = C.c_void_p()
                assert L.ucc_service_bcast(team.members[r].team, bb[r].ctypes.data, 17, 1, _Subset(mk(), i), C.byref(q)) == U.UCC_OK
                reqs[r] = q
            wait_all(reqs)
            for r in members:
                assert np.array_equal(bb[r], np.arange(17, dtype=np.uint8) + 5), (name, r)


def test_config_file_team_sections(tmp_path):
    """[sections] predicated on team facts carry per-team TUNE strings (reference ucc_add_team_sections, tl_ucp_team.c:82-88):
    one process creates teams of 5 and of 3 ranks and each gets the algorithm its section names."""
    import subprocess
    import sys
    f = tmp_path / "ucc.conf"
    f.write_text("[big team_size=4-64]\nUCC_TL_SHM_TUNE = allreduce:0-inf:@ring\n\n[small team_size=2-3]\nUCC_TL_SHM_TUNE = allreduce:0-inf:@dbt\n\n"
                 "[never team_size=2-64 nnodes=5-9]\nUCC_TL_SHM_TUNE = allreduce:0-inf:@sliding_window\n")
    code = (
        "import numpy as np, sys\n"
        "from ucc_b200 import capi as U\n"
        "from ucc_b200.harness import UccJob, coll_args\n"
        "for n in (5, 3):\n"
        "    j = UccJob(n); t = j.create_team()\n"
        "    s = [np.full(4096, r + 1.0, np.float32) for r in range(n)]; d = [np.zeros(4096, np.float32) for _ in range(n)]\n"
        "    q = t.coll([coll_args('allreduce', s[r], d[r]) for r in range(n)]); assert q.run() == 0; q.finalize()\n"
        "    assert np.all(d[0] == n * (n + 1) / 2)\n"
        "    sys.stdout.flush(); sys.stderr.flush(); print('TEAM_DONE', n, flush=True)\n"
        "    j.cleanup()\n")
    env = dict(os.environ, UCC_CONFIG_FILE=str(f), UCC_LOG_LEVEL="info", PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env.pop("UCC_TL_SHM_TUNE", None)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    txt = out.stdout + out.stderr
    first, second = txt.split("TEAM_DONE 5")[0], txt.split("TEAM_DONE 5")[1]
    assert "sliding_window" not in txt
    # at INFO level cl/basic prints the TL-level map: "allreduce host: {0..4K}:TL_SHM:10:<algorithm> ..."
    sel = lambda t: [ln for ln in t.splitlines() if "allreduce host:" in ln and "TL_SHM" in ln]  # noqa: E731
    assert sel(first) and all(":ring" in ln and ":dbt" not in ln for ln in sel(first)), first[-1500:]
    assert sel(second) and all(":dbt" in ln for ln in sel(second)), second[-1500:]


def test_stale_shared_memory_segments_are_reaped():
    """A process killed with SIGKILL cannot unlink its named POSIX segments.  The first context of a later process removes the
    segments whose owner is dead (pid gone AND the owner's flock gone), and never those of a live process (src/utils/ucc_sys.c
    ucc_shm_reap_stale; the reference avoids the problem with SysV segments marked IPC_RMID at creation)."""
    import glob
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    child = ("from ucc_b200.harness import UccJob\nimport time\nj = UccJob(2, env={'UCC_TLS': 'shm,self'}); t = j.create_team(range(2))\n"
             "print('UP', flush=True); time.sleep(120)")
    other = "from ucc_b200.harness import UccJob\nj = UccJob(2, env={'UCC_TLS': 'shm,self'}); j.cleanup()"
    p = subprocess.Popen([sys.executable, "-c", child], stdout=subprocess.PIPE, text=True, env=env)
    try:
        for line in p.stdout:                      # (library warnings may precede it)
            if line.strip() == "UP":
                break
        else:
            raise AssertionError("the child did not come up")
        mine = lambda: glob.glob(f"/dev/shm/ucc_b200.{p.pid}.*")   # noqa: E731
        assert len(mine()) == 2
        subprocess.run([sys.executable, "-c", other], check=True, env=env, timeout=120)
        assert len(mine()) == 2, "segments of a LIVE process were removed"
    finally:
        p.kill()
        p.wait()
    assert len(mine()) == 2          # SIGKILL: nobody cleaned up
    subprocess.run([sys.executable, "-c", other], check=True, env=env, timeout=120)
    assert mine() == []
