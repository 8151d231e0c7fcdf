"""ctypes access to *internal* entry points of libucc.so (mc / ec / coll_score / parser / topo).
Used by the unit tests the same way the reference's gtests link against internal symbols
(test/gtest/core/test_mc_reduce.cc, test_ec_cuda.cc, coll_score/*.cc, utils/*.cc)."""
import ctypes as C

from . import capi as U

lib = U.lib

# ------------------------------------------------------------------ MC
class mc_buffer_header(C.Structure):
    _fields_ = [("mt", C.c_int), ("from_pool", C.c_int), ("addr", C.c_void_p)]


class mem_attr(C.Structure):
    _fields_ = [("field_mask", C.c_uint64), ("mem_type", C.c_int), ("base_address", C.c_void_p), ("alloc_length", C.c_size_t)]


lib.ucc_mc_alloc.restype = C.c_int
lib.ucc_mc_alloc.argtypes = [C.POINTER(C.POINTER(mc_buffer_header)), C.c_size_t, C.c_int]
lib.ucc_mc_free.restype = C.c_int
lib.ucc_mc_free.argtypes = [C.POINTER(mc_buffer_header)]
lib.ucc_mc_memcpy.restype = C.c_int
lib.ucc_mc_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
lib.ucc_mc_memset.restype = C.c_int
lib.ucc_mc_memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_int]
lib.ucc_mc_get_mem_attr.restype = C.c_int
lib.ucc_mc_get_mem_attr.argtypes = [C.c_void_p, C.POINTER(mem_attr)]
lib.ucc_mc_available.restype = C.c_int
lib.ucc_mc_available.argtypes = [C.c_int]

# ------------------------------------------------------------------ EC
EE_TASK_REDUCE, EE_TASK_REDUCE_STRIDED, EE_TASK_REDUCE_MULTI_DST, EE_TASK_COPY, EE_TASK_COPY_MULTI = 1, 2, 4, 8, 16
EEE_FLAG_ALPHA, EEE_FLAG_SRCS_EXT = 1, 2
NUM_BUFS, MULTI_BUFS = 9, 7


class eee_reduce(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("srcs", C.c_void_p * NUM_BUFS), ("count", C.c_size_t), ("alpha", C.c_double),
                ("dt", C.c_uint64), ("op", C.c_int), ("n_srcs", C.c_uint16)]


class eee_reduce_strided(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src1", C.c_void_p), ("src2", C.c_void_p), ("stride", C.c_size_t), ("count", C.c_size_t),
                ("alpha", C.c_double), ("dt", C.c_uint64), ("op", C.c_int), ("n_src2", C.c_uint16)]


class eee_reduce_multi_dst(C.Structure):
    _fields_ = [("dst", C.c_void_p * MULTI_BUFS), ("src1", C.c_void_p * MULTI_BUFS), ("src2", C.c_void_p * MULTI_BUFS),
                ("counts", C.c_size_t * MULTI_BUFS), ("dt", C.c_uint64), ("op", C.c_int), ("n_bufs", C.c_uint16)]


class eee_copy(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("src", C.c_void_p), ("len", C.c_size_t)]


class eee_copy_multi(C.Structure):
    _fields_ = [("src", C.c_void_p * MULTI_BUFS), ("dst", C.c_void_p * MULTI_BUFS), ("counts", C.c_size_t * MULTI_BUFS), ("num_vectors", C.c_size_t)]


class _eee_u(C.Union):
    _fields_ = [("reduce", eee_reduce), ("reduce_strided", eee_reduce_strided), ("reduce_multi_dst", eee_reduce_multi_dst),
                ("copy", eee_copy), ("copy_multi", eee_copy_multi)]


class eee_task_args(C.Structure):
    _anonymous_ = ("u",)
    _fields_ = [("task_type", C.c_uint16), ("flags", C.c_uint16), ("u", _eee_u)]


class eee_params(C.Structure):
    _fields_ = [("mask", C.c_uint64), ("ee_type", C.c_int), ("task_types", C.c_uint64)]


for name, args in (("ucc_ee_executor_init", [C.POINTER(eee_params), C.POINTER(C.c_void_p)]), ("ucc_ee_executor_start", [C.c_void_p, C.c_void_p]),
                   ("ucc_ee_executor_status", [C.c_void_p]), ("ucc_ee_executor_stop", [C.c_void_p]), ("ucc_ee_executor_finalize", [C.c_void_p]),
                   ("ucc_ee_executor_task_post", [C.c_void_p, C.POINTER(eee_task_args), C.POINTER(C.c_void_p)]),
                   ("ucc_ee_executor_task_test", [C.c_void_p]), ("ucc_ee_executor_task_finalize", [C.c_void_p]),
                   ("ucc_ec_available", [C.c_int]), ("ucc_ec_create_event", [C.POINTER(C.c_void_p), C.c_int]),
                   ("ucc_ec_destroy_event", [C.c_void_p, C.c_int]), ("ucc_ec_event_post", [C.c_void_p, C.c_void_p, C.c_int]),
                   ("ucc_ec_event_test", [C.c_void_p, C.c_int])):
    f = getattr(lib, name)
    f.restype, f.argtypes = C.c_int, args


class Executor:
    def __init__(self, ee_type, stream=None, task_types=None):
        p = eee_params()
        p.mask, p.ee_type = 1, ee_type
        if task_types is not None:
            p.mask |= 2
            p.task_types = task_types
        self.h = C.c_void_p()
        U.check(lib.ucc_ee_executor_init(C.byref(p), C.byref(self.h)), "executor_init")
        U.check(lib.ucc_ee_executor_start(self.h, stream), "executor_start")
        while lib.ucc_ee_executor_status(self.h) == U.UCC_INPROGRESS:
            pass

    def run(self, args):
        t = C.c_void_p()
        U.check(lib.ucc_ee_executor_task_post(self.h, C.byref(args), C.byref(t)), "task_post")
        while True:
            st = lib.ucc_ee_executor_task_test(t)
            if st != U.UCC_INPROGRESS:
                break
        lib.ucc_ee_executor_task_finalize(t)
        U.check(st, "task")

    def close(self):
        lib.ucc_ee_executor_stop(self.h)
        lib.ucc_ee_executor_finalize(self.h)


# ------------------------------------------------------------------ coll_score
lib.ucc_coll_score_alloc.restype = C.c_int
lib.ucc_coll_score_alloc.argtypes = [C.POINTER(C.c_void_p)]
lib.ucc_coll_score_free.restype = None
lib.ucc_coll_score_free.argtypes = [C.c_void_p]
INIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p))
ALG_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(INIT_FN))
lib.ucc_coll_score_add_range.restype = C.c_int
lib.ucc_coll_score_add_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_uint32, INIT_FN, C.c_void_p]
lib.ucc_coll_score_merge.restype = C.c_int
lib.ucc_coll_score_merge.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int]
lib.ucc_coll_score_alloc_from_str.restype = C.c_int
lib.ucc_coll_score_alloc_from_str.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.c_uint32, INIT_FN, C.c_void_p, ALG_FN]
lib.ucc_coll_score_update.restype = C.c_int
lib.ucc_coll_score_update.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_uint64]
lib.ucc_coll_score_build_map.restype = C.c_int
lib.ucc_coll_score_build_map.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
lib.ucc_coll_score_free_map.restype = None
lib.ucc_coll_score_free_map.argtypes = [C.c_void_p]


class list_link(C.Structure):
    pass


list_link._fields_ = [("prev", C.POINTER(list_link)), ("next", C.POINTER(list_link))]


class coll_entry(C.Structure):
    _fields_ = [("list_elem", list_link), ("score", C.c_uint32), ("init", C.c_void_p), ("team", C.c_void_p)]


class msg_range(C.Structure):
    _fields_ = [("super", coll_entry), ("fallback", list_link), ("start", C.c_size_t), ("end", C.c_size_t)]


class coll_score(C.Structure):
    _fields_ = [("scores", (list_link * 5) * 16)]


def score_ranges(score_ptr, coll_idx, mt):
    """-> [(start, end, score, init_ptr, [fallback (score, init)])]"""
    sc = C.cast(score_ptr, C.POINTER(coll_score)).contents
    head = sc.scores[coll_idx][mt]
    head_addr = C.addressof(head)
    out = []
    cur = head.next
    while C.addressof(cur.contents) != head_addr:
        r = C.cast(cur, C.POINTER(msg_range)).contents
        fbs = []
        fh = C.addressof(r.fallback)
        f = r.fallback.next
        while C.addressof(f.contents) != fh:
            e = C.cast(f, C.POINTER(coll_entry)).contents
            fbs.append((e.score, e.init))
            f = e.list_elem.next
        out.append((r.start, r.end, r.super.score, r.super.init, fbs))
        cur = r.super.list_elem.next
    return out


# ------------------------------------------------------------------ misc utils
lib.ucc_str_to_memunits.restype = C.c_int
lib.ucc_str_to_memunits.argtypes = [C.c_char_p, C.POINTER(C.c_size_t)]
lib.ucc_ilog2 if hasattr(lib, "ucc_ilog2") else None
