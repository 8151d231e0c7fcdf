"""Single-process multi-rank harness (the gtest-style ``UccJob`` of the reference,
test/gtest/common/test_ucc.{h,cc}, re-done in Python on top of the C API).

N contexts live in one OS process.  Context creation blocks inside the
library on the OOB allgather, so contexts are created from N short-lived
threads; everything else (team creation, collectives) is driven
round-robin from the calling thread, exactly like a single-threaded MPI-less
test driver would.  A synthetic process placement (host / socket / numa)
can be injected per rank to exercise topology code without a cluster.
"""
import ctypes as C
import itertools
import os
import threading

import numpy as np

from . import capi as U

NP_DT = {"int8": np.int8, "int16": np.int16, "int32": np.int32, "int64": np.int64, "uint8": np.uint8,
         "uint16": np.uint16, "uint32": np.uint32, "uint64": np.uint64, "float16": np.float16, "float32": np.float32,
         "float64": np.float64, "float32_complex": np.complex64, "float64_complex": np.complex128,
         "float128": np.longdouble, "float128_complex": np.clongdouble}


class OobGroup:
    """In-memory allgather shared by `n` endpoints living in this process."""
    _ids = itertools.count(1)

    def __init__(self, n):
        self.n = n
        self.lock = threading.Lock()
        self.calls = {}
        self.count = [0] * n
        self.reqs = {}
        self._keep = []

    def oob(self, rank):
        grp = self

        def allgather(src, recv, size, info, req_pp):
            with grp.lock:
                k = grp.count[rank]
                grp.count[rank] += 1
                e = grp.calls.setdefault(k, {"data": [None] * grp.n, "n": 0, "left": grp.n})
                e["data"][rank] = C.string_at(src, size)
                e["n"] += 1
                rid = next(OobGroup._ids)
                grp.reqs[rid] = (k, recv, size)
            req_pp[0] = rid
            return U.UCC_OK

        def req_test(req):
            with grp.lock:
                k, recv, size = grp.reqs[req]
                e = grp.calls[k]
                if e["n"] < grp.n:
                    return U.UCC_INPROGRESS
                C.memmove(recv, b"".join(e["data"]), size * grp.n)
            return U.UCC_OK

        def req_free(req):
            with grp.lock:
                k, _, _ = grp.reqs.pop(req)
                e = grp.calls[k]
                e["left"] -= 1
                if e["left"] == 0:
                    del grp.calls[k]
            return U.UCC_OK

        o = U.ucc_oob_coll_t()
        o.allgather = U.OOB_ALLGATHER_FN(allgather)
        o.req_test = U.OOB_REQ_FN(req_test)
        o.req_free = U.OOB_REQ_FN(req_free)
        o.coll_info = None
        o.n_oob_eps = self.n
        o.oob_ep = rank
        self._keep.append(o)
        return o


def _read_lib_config(env=None, cls=None):
    cfg = U.handle()
    U.check(U.ucc_lib_config_read(None, None, C.byref(cfg)), "lib_config_read")
    if cls:
        U.check(U.ucc_lib_config_modify(cfg, b"CLS", cls.encode()), "lib_config_modify")
    return cfg


class UccProcess:
    """One (lib, context) pair = one emulated rank."""

    def __init__(self, job, rank, thread_mode=U.UCC_THREAD_SINGLE, cls=None):
        self.job, self.rank = job, rank
        cfg = _read_lib_config(cls=cls)
        p = U.ucc_lib_params_t()
        p.mask = U.UCC_LIB_PARAM_FIELD_THREAD_MODE
        p.thread_mode = thread_mode
        self.lib = U.handle()
        st = U.ucc_init_version(U.UCC_API_MAJOR, U.UCC_API_MINOR, C.byref(p), cfg, C.byref(self.lib))
        U.ucc_lib_config_release(cfg)
        U.check(st, "ucc_init")
        self.ctx = U.handle()

    def create_context(self, oob, proc_info=None, ctx_modify=()):
        ccfg = U.handle()
        U.check(U.ucc_context_config_read(self.lib, None, C.byref(ccfg)), "context_config_read")
        for comp, name, val in ctx_modify:
            U.check(U.ucc_context_config_modify(ccfg, comp.encode() if comp else None, name.encode(), val.encode()), "ctx_config_modify")
        params = U.ucc_context_params_t()
        if oob is not None:
            params.mask = U.UCC_CONTEXT_PARAM_FIELD_OOB
            params.oob = oob
        if proc_info is not None:
            st = U.ucc_context_create_proc_info(self.lib, C.byref(params), ccfg, C.byref(self.ctx), C.byref(proc_info))
        else:
            st = U.ucc_context_create(self.lib, C.byref(params), ccfg, C.byref(self.ctx))
        U.ucc_context_config_release(ccfg)
        self.ctx_status = st

    def progress(self):
        return U.ucc_context_progress(self.ctx)

    def destroy(self):
        if self.ctx:
            U.ucc_context_destroy(self.ctx)
            self.ctx = U.handle()
        if self.lib:
            U.ucc_finalize(self.lib)
            self.lib = U.handle()


def fake_proc_info(rank, ppn=None, sockets_per_node=2, numas_per_socket=1, base_pid=None):
    """Synthetic placement: `ppn` ranks per node, spread over sockets / numas."""
    pi = U.ucc_proc_info_t()
    if ppn is None:
        pi.host_hash = 0xC0FFEE
        pi.socket_id = 0xFF
        pi.numa_id = 0xFF
    else:
        local = rank % ppn
        per_sock = max(1, ppn // sockets_per_node)
        pi.host_hash = 1000 + rank // ppn
        pi.socket_id = min(local // per_sock, sockets_per_node - 1)
        per_numa = max(1, per_sock // numas_per_socket)
        pi.numa_id = pi.socket_id * numas_per_socket + min((local % per_sock) // per_numa, numas_per_socket - 1)
    pi.host_id = 0
    pi.pid = (base_pid if base_pid is not None else os.getpid())
    return pi


class UccReq:
    """A collective instantiated on every member of a team."""

    def __init__(self, team, args_list):
        self.team = team
        self.args = args_list
        self.reqs = []
        self.status = []
        for m, a in zip(team.members, args_list):
            r = C.POINTER(U.ucc_coll_req_t)()
            st = U.ucc_collective_init(C.byref(a), C.byref(r), m.team)
            self.status.append(st)
            self.reqs.append(r if st == U.UCC_OK else None)
        bad = [s for s in self.status if s != U.UCC_OK]
        if bad:
            for r in self.reqs:
                if r:
                    U.ucc_collective_finalize(r)
            self.reqs = []
            raise U.UccError(bad[0], "ucc_collective_init")

    def post(self):
        for r in self.reqs:
            U.check(U.ucc_collective_post(r), "ucc_collective_post")

    def test(self):
        worst = U.UCC_OK
        for r in self.reqs:
            st = r.contents.status
            if st < 0:
                return st
            if st != U.UCC_OK:
                worst = U.UCC_INPROGRESS
        return worst

    def wait(self, max_iters=20_000_000, max_seconds=120.0):
        import time
        procs = self.team.job.procs
        it = 0
        t0 = time.monotonic()
        while True:
            st = self.test()
            if st != U.UCC_INPROGRESS:
                return st
            for p in procs:
                p.progress()
            it += 1
            if it > max_iters or ((it & 0x3ff) == 0 and time.monotonic() - t0 > max_seconds):
                raise TimeoutError("collective did not complete")

    def finalize(self):
        for r in self.reqs:
            U.check(U.ucc_collective_finalize(r), "ucc_collective_finalize")
        self.reqs = []

    def run(self):
        self.post()
        st = self.wait()
        return st


class _Member:
    def __init__(self, proc, rank):
        self.proc, self.rank, self.team = proc, rank, U.handle()


class UccTeam:
    def __init__(self, job, ranks, use_ep_map=False, team_id=None):
        self.job = job
        self.ranks = list(ranks)
        n = len(self.ranks)
        self.members = [_Member(job.procs[r], i) for i, r in enumerate(self.ranks)]
        self.oob = OobGroup(n)
        self._keep = []
        for i, m in enumerate(self.members):
            p = U.ucc_team_params_t()
            p.mask = U.UCC_TEAM_PARAM_FIELD_EP | U.UCC_TEAM_PARAM_FIELD_EP_RANGE | U.UCC_TEAM_PARAM_FIELD_OOB
            p.ep = i
            p.ep_range = U.UCC_COLLECTIVE_EP_RANGE_CONTIG
            p.oob = self.oob.oob(i)
            if use_ep_map:
                arr = (C.c_uint64 * n)(*self.ranks)
                self._keep.append(arr)
                p.mask |= U.UCC_TEAM_PARAM_FIELD_EP_MAP
                p.ep_map.type = U.UCC_EP_MAP_ARRAY
                p.ep_map.ep_num = n
                p.ep_map.array.map = C.cast(arr, C.c_void_p)
                p.ep_map.array.elem_size = 8
            if team_id is not None:
                p.mask |= U.UCC_TEAM_PARAM_FIELD_ID
                p.id = team_id
            ctxs = (U.handle * 1)(m.proc.ctx)
            U.check(U.ucc_team_create_post(ctxs, 1, C.byref(p), C.byref(m.team)), "team_create_post")
        self._wait_create()

    def _wait_create(self, max_iters=5_000_000, max_seconds=180.0):
        import time
        it = 0
        t0 = time.monotonic()
        while True:
            pending = False
            for m in self.members:
                st = U.ucc_team_create_test(m.team)
                if st < 0:
                    raise U.UccError(st, "team_create_test")
                if st == U.UCC_INPROGRESS:
                    pending = True
            if not pending:
                return
            for p in self.job.procs:
                p.progress()
            it += 1
            if it > max_iters or ((it & 0xff) == 0 and time.monotonic() - t0 > max_seconds):
                raise TimeoutError("team creation did not complete")

    @property
    def size(self):
        return len(self.ranks)

    def coll(self, args_list):
        return UccReq(self, args_list)

    def destroy(self):
        pending = list(self.members)
        it = 0
        while pending:
            nxt = []
            for m in pending:
                st = U.ucc_team_destroy(m.team)
                if st == U.UCC_INPROGRESS:
                    nxt.append(m)
                elif st < 0:
                    raise U.UccError(st, "team_destroy")
            pending = nxt
            for p in self.job.procs:
                p.progress()
            it += 1
            if it > 1_000_000:
                raise TimeoutError("team destroy did not complete")
        self.members = []


class UccJob:
    """N emulated ranks in this process."""

    def __init__(self, n_procs, ppn=None, thread_mode=U.UCC_THREAD_SINGLE, env=None, cls=None, ctx_modify=(),
                 sockets_per_node=2, numas_per_socket=1, with_ctx_oob=True):
        self.n = n_procs
        self._saved_env = {}
        for k, v in (env or {}).items():
            self._saved_env[k] = os.environ.get(k)
            os.environ[k] = v
        self.procs = [UccProcess(self, r, thread_mode, cls=cls) for r in range(n_procs)]
        self.ctx_oob = OobGroup(n_procs)
        self.teams = []
        # fake pids keep endpoints distinct when a synthetic placement is injected
        pis = [fake_proc_info(r, ppn, sockets_per_node, numas_per_socket, base_pid=None) for r in range(n_procs)]

        def mk(r):
            self.procs[r].create_context(self.ctx_oob.oob(r) if (with_ctx_oob and n_procs > 1) else None,
                                         pis[r] if ppn is not None else None, ctx_modify)

        threads = [threading.Thread(target=mk, args=(r,)) for r in range(n_procs)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        bad = [p.ctx_status for p in self.procs if p.ctx_status != U.UCC_OK]
        if bad:
            raise U.UccError(bad[0], "ucc_context_create")

    def create_team(self, ranks=None, **kw):
        t = UccTeam(self, ranks if ranks is not None else range(self.n), **kw)
        self.teams.append(t)
        return t

    def progress(self):
        for p in self.procs:
            p.progress()

    def cleanup(self):
        for t in self.teams:
            if t.members:
                t.destroy()
        self.teams = []
        for p in self.procs:
            p.destroy()
        for k, v in self._saved_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        # after a failed test requests may still be in flight: tearing teams down under them
        # would turn the assertion into a crash, so leak the job instead
        if exc_type is None:
            self.cleanup()


# ----------------------------------------------------------------- args builders
def ptr(a):
    return a.ctypes.data if a is not None else None


def coll_args(coll, src=None, dst=None, dt="float32", op="sum", root=0, count_src=None, count_dst=None, inplace=False,
              persistent=False, mem_type=U.UCC_MEMORY_TYPE_HOST, src_counts=None, src_displs=None, dst_counts=None,
              dst_displs=None, flags=0, timeout=None, active_set=None, tag=None, src_ptr=None, dst_ptr=None,
              src_dt=None, dst_dt=None, src_mem_type=None, dst_mem_type=None):
    """Build ucc_coll_args_t from numpy buffers (or raw pointers + counts)."""
    a = U.ucc_coll_args_t()
    a.coll_type = U.COLL[coll]
    a.op = U.OP[op]
    a.root = root
    f = flags
    if inplace:
        f |= U.UCC_COLL_ARGS_FLAG_IN_PLACE
    if persistent:
        f |= U.UCC_COLL_ARGS_FLAG_PERSISTENT
    if timeout is not None:
        f |= U.UCC_COLL_ARGS_FLAG_TIMEOUT
        a.timeout = timeout
    if f:
        a.mask |= U.UCC_COLL_ARGS_FIELD_FLAGS
        a.flags = f
    if active_set is not None:
        a.mask |= U.UCC_COLL_ARGS_FIELD_ACTIVE_SET
        a.active_set.start, a.active_set.stride, a.active_set.size = active_set
    if tag is not None:
        a.mask |= U.UCC_COLL_ARGS_FIELD_TAG
        a.tag = tag
    sdt = U.DT[src_dt or dt]
    ddt = U.DT[dst_dt or dt]
    smt = mem_type if src_mem_type is None else src_mem_type
    dmt = mem_type if dst_mem_type is None else dst_mem_type
    sp = src_ptr if src_ptr is not None else ptr(src)
    dp = dst_ptr if dst_ptr is not None else ptr(dst)
    keep = []
    if src_counts is not None:
        c = np.ascontiguousarray(src_counts, dtype=np.uint64)
        d = np.ascontiguousarray(src_displs, dtype=np.uint64)
        keep += [c, d]
        a.src.info_v.buffer, a.src.info_v.counts, a.src.info_v.displacements = sp, ptr(c), ptr(d)
        a.src.info_v.datatype, a.src.info_v.mem_type = sdt, smt
        a.mask |= U.UCC_COLL_ARGS_FIELD_FLAGS
        a.flags |= U.UCC_COLL_ARGS_FLAG_COUNT_64BIT | U.UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT
    else:
        a.src.info.buffer = sp
        a.src.info.count = count_src if count_src is not None else (src.size if src is not None else 0)
        a.src.info.datatype, a.src.info.mem_type = sdt, smt
    if dst_counts is not None:
        c = np.ascontiguousarray(dst_counts, dtype=np.uint64)
        d = np.ascontiguousarray(dst_displs, dtype=np.uint64)
        keep += [c, d]
        a.dst.info_v.buffer, a.dst.info_v.counts, a.dst.info_v.displacements = dp, ptr(c), ptr(d)
        a.dst.info_v.datatype, a.dst.info_v.mem_type = ddt, dmt
        a.mask |= U.UCC_COLL_ARGS_FIELD_FLAGS
        a.flags |= U.UCC_COLL_ARGS_FLAG_COUNT_64BIT | U.UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT
    else:
        a.dst.info.buffer = dp
        a.dst.info.count = count_dst if count_dst is not None else (dst.size if dst is not None else 0)
        a.dst.info.datatype, a.dst.info.mem_type = ddt, dmt
    a._keep = keep + [src, dst]
    return a
