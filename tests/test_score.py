"""coll_score unit tests (model: reference test/gtest/coll_score/test_score.cc, test_score_str.cc, test_score_update.cc):
range insertion, merge with fallbacks, the TUNE grammar, update semantics."""
import ctypes as C

import pytest

from ucc_b200 import capi as U
from ucc_b200 import internal as I

lib = I.lib
INF = (1 << 64) - 1
AR = 2          # index of allreduce (UCC_COLL_TYPE_ALLREDUCE = bit 2)
BCAST = 6
HOST, CUDA = 0, 1


def _mk_init(tag):
    @I.INIT_FN
    def fn(a, b, c):
        return 0
    fn.tag = tag
    return fn


INITS = [_mk_init(i) for i in range(6)]


def addr(fn):
    return C.cast(fn, C.c_void_p).value


def alloc():
    p = C.c_void_p()
    assert lib.ucc_coll_score_alloc(C.byref(p)) == 0
    return p


def add(s, coll_idx, mt, start, end, score, init, team=None):
    assert lib.ucc_coll_score_add_range(s, 1 << coll_idx, mt, start, end, score, init, team) == 0


def ranges(s, coll_idx=AR, mt=HOST):
    return [(a, b, sc, ini, fb) for (a, b, sc, ini, fb) in I.score_ranges(s, coll_idx, mt)]


def test_add_range_sorted_and_disjoint():
    s = alloc()
    add(s, AR, HOST, 100, 200, 10, INITS[0])
    add(s, AR, HOST, 0, 50, 5, INITS[1])
    add(s, AR, HOST, 300, INF, 7, INITS[2])
    r = ranges(s)
    assert [(a, b, sc) for a, b, sc, _, _ in r] == [(0, 50, 5), (100, 200, 10), (300, INF, 7)]
    lib.ucc_coll_score_free(s)


def merged(specs1, specs2):
    s1, s2 = alloc(), alloc()
    for sp in specs1:
        add(s1, AR, HOST, *sp)
    for sp in specs2:
        add(s2, AR, HOST, *sp)
    out = C.c_void_p()
    assert lib.ucc_coll_score_merge(s1, s2, C.byref(out), 1) == 0
    return out


def test_merge_non_overlap():
    m = merged([(0, 10, 1, INITS[0])], [(10, 20, 2, INITS[1])])
    assert [(a, b, sc) for a, b, sc, _, _ in ranges(m)] == [(0, 10, 1), (10, 20, 2)]
    lib.ucc_coll_score_free(m)


def test_merge_overlap_higher_wins_loser_is_fallback():
    m = merged([(0, 100, 10, INITS[0])], [(50, 150, 20, INITS[1])])
    r = ranges(m)
    assert [(a, b, sc) for a, b, sc, _, _ in r] == [(0, 50, 10), (50, 100, 20), (100, 150, 20)] or \
        [(a, b, sc) for a, b, sc, _, _ in r] == [(0, 50, 10), (50, 150, 20)]
    mid = [x for x in r if x[0] == 50][0]
    assert mid[3] == addr(INITS[1])
    assert [f[0] for f in mid[4]] == [10] and mid[4][0][1] == addr(INITS[0])
    lib.ucc_coll_score_free(m)


def test_merge_inside_and_same_score():
    m = merged([(0, 100, 10, INITS[0])], [(20, 40, 30, INITS[1])])
    r = ranges(m)
    assert [(a, b, sc) for a, b, sc, _, _ in r] == [(0, 20, 10), (20, 40, 30), (40, 100, 10)]
    lib.ucc_coll_score_free(m)
    # adjacent pieces with the same score / init are glued back together
    m = merged([(0, 50, 10, INITS[0])], [(50, 100, 10, INITS[0])])
    assert [(a, b, sc) for a, b, sc, _, _ in ranges(m)] == [(0, 100, 10)]
    lib.ucc_coll_score_free(m)


def from_str(s, size=8, init=INITS[0], alg_fn=None):
    p = C.c_void_p()
    st = lib.ucc_coll_score_alloc_from_str(s.encode(), C.byref(p), size, init, None, alg_fn if alg_fn else C.cast(None, I.ALG_FN))
    return st, p


def test_tune_string_grammar():
    st, p = from_str("allreduce:cuda:0-4k:10#bcast:host:1M-inf:inf")
    assert st == 0
    r = ranges(p, AR, CUDA)
    assert [(a, b, sc) for a, b, sc, _, _ in r] == [(0, 4096, 10)]
    assert ranges(p, AR, HOST) == []
    r = ranges(p, BCAST, HOST)
    assert r[0][0] == 1 << 20 and r[0][1] == INF and r[0][2] == 0x7fffffff
    lib.ucc_coll_score_free(p)
    # several colls / ranges in one token, all memory types when none is given
    st, p = from_str("allreduce,bcast:0-1k,4k-8k:3")
    assert st == 0
    for c in (AR, BCAST):
        for mt in (HOST, CUDA):
            assert [(a, b, sc) for a, b, sc, _, _ in ranges(p, c, mt)] == [(0, 1024, 3), (4096, 8192, 3)]
    lib.ucc_coll_score_free(p)


def test_tune_string_team_size_filter():
    st, p = from_str("allreduce:0-inf:[2-4,16]:5", size=8)   # team size 8 is outside -> token ignored
    assert st == 0 and ranges(p) == []
    lib.ucc_coll_score_free(p)
    st, p = from_str("allreduce:0-inf:[2-4,8]:5", size=8)
    assert st == 0 and [(a, b, sc) for a, b, sc, _, _ in ranges(p)] == [(0, INF, 5)]
    lib.ucc_coll_score_free(p)


@pytest.mark.parametrize("bad", ["allreduce:abc", "nosuchcoll:10", "allreduce:5-1:3", "allreduce:@", "allreduce:[a-b]:3"])
def test_tune_string_errors(bad):
    st, p = from_str(bad)
    assert st != 0


def test_tune_alg_token():
    seen = {}

    @I.ALG_FN
    def alg_fn(alg_id, alg_str, coll, mt, out):
        seen["call"] = (alg_id, alg_str, coll, mt)
        if alg_str == b"ring" or alg_id == 1:
            out[0] = INITS[3]
            return 0
        return -6  # not found

    st, p = from_str("allreduce:host:0-inf:@ring", alg_fn=alg_fn)
    assert st == 0
    r = ranges(p)
    assert r[0][3] == addr(INITS[3])
    lib.ucc_coll_score_free(p)
    st, p = from_str("allreduce:host:0-inf:@1:77", alg_fn=alg_fn)
    assert st == 0 and ranges(p)[0][2] == 77 and ranges(p)[0][3] == addr(INITS[3])
    lib.ucc_coll_score_free(p)
    st, p = from_str("allreduce:host:0-inf:@nosuch", alg_fn=alg_fn)
    assert st != 0


def test_update_overrides_and_keeps():
    base = alloc()
    add(base, AR, HOST, 0, INF, 10, INITS[0])
    st, upd = from_str("allreduce:host:1k-2k:50", init=INITS[1])
    assert st == 0
    mts = (C.c_int * 1)(HOST)
    assert lib.ucc_coll_score_update(base, upd, 10, C.cast(mts, C.c_void_p), 1, 1 << AR) == 0
    r = ranges(base)
    assert [(a, b, sc) for a, b, sc, _, _ in r] == [(0, 1024, 10), (1024, 2048, 50), (2048, INF, 10)]
    # score-only token keeps the original init function
    assert all(x[3] == addr(INITS[0]) for x in r)
    lib.ucc_coll_score_free(upd)
    lib.ucc_coll_score_free(base)


def test_update_score_zero_disables():
    base = alloc()
    add(base, AR, HOST, 0, INF, 10, INITS[0])
    st, upd = from_str("allreduce:host:0-4k:0")
    assert st == 0
    mts = (C.c_int * 1)(HOST)
    assert lib.ucc_coll_score_update(base, upd, 10, C.cast(mts, C.c_void_p), 1, 1 << AR) == 0
    r = ranges(base)
    assert [(a, b, sc) for a, b, sc, _, _ in r] == [(4096, INF, 10)]
    lib.ucc_coll_score_free(upd)
    lib.ucc_coll_score_free(base)


def test_memunits_parser():
    v = C.c_size_t()
    for s, exp in (("4k", 4096), ("4K", 4096), ("1m", 1 << 20), ("2G", 2 << 30), ("17", 17), ("8b", 8), ("1kb", 1024), ("inf", INF)):
        assert lib.ucc_str_to_memunits(s.encode(), C.byref(v)) == 0, s
        assert v.value == exp, (s, v.value)
    assert lib.ucc_str_to_memunits(b"12q", C.byref(v)) != 0


def test_tl_nvl_default_selection_strings_parse():
    """The built-in selection strings of tl/nvl (tl_nvl_coll.c get_scores), incl. the NVLS branch that only real multi-GPU NVSwitch
    teams take: every token must parse and every algorithm name must resolve through the TL's own alg_id_to_init - a typo there
    would make team creation fail on exactly the machines the CPU suite cannot emulate."""
    import os
    path = os.path.join(os.path.dirname(U.LIB_PATH), "ucc", "libucc_tl_nvl.so")
    if not os.path.exists(path):
        pytest.skip("tl/nvl plugin not built")
    try:
        nvl = C.CDLL(path)
    except OSError as e:
        pytest.skip(f"tl/nvl plugin not loadable here: {e}")
    alg_fn = C.cast(nvl.ucc_tl_nvl_alg_id_to_init, I.ALG_FN)
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "src", "components", "tl", "nvl", "tl_nvl_coll.c")).read()
    import re
    fmts = re.findall(r'snprintf\(sel[^;]*?"([^"]*)"', src)
    assert len(fmts) >= 3, fmts
    for f in fmts:
        s = f.replace("%s-inf:@%s", "512M-inf:@nvls_pipe").replace("0-%s", "0-1048577").replace("%s-inf", "512M-inf").lstrip("#")
        assert "%" not in s, (f, s)
        st, p = from_str(s, size=8, alg_fn=alg_fn)
        assert st == 0, (s, st)
        lib.ucc_coll_score_free(p)
    for alg in ("nvls", "nvls_pipe", "twoshot", "oneshot", "ring", "rhd"):
        st, p = from_str(f"allreduce:cuda:512M-inf:@{alg}", size=8, alg_fn=alg_fn)
        assert st == 0, alg
        lib.ucc_coll_score_free(p)


def test_merge_random_maps_against_a_pointwise_oracle():
    """Property test (hypothesis): for random disjoint range lists in two score maps, the merged map is sorted and disjoint and, at
    every probe point, its winner is the covering range with the highest score (map 1 on ties) and its fallback list holds the
    other covering range - the boundary-sweep merge (ucc_coll_score.c) against a point-by-point oracle."""
    from hypothesis import given, settings, strategies as st

    def disjoint(cuts, scores, inits):
        cuts = sorted(set(cuts))
        out = []
        for i in range(0, len(cuts) - 1, 2):                   # every other gap is a range
            out.append((cuts[i], cuts[i + 1], scores[i % len(scores)], INITS[inits[i % len(inits)]]))
        return out

    lists = st.lists(st.integers(0, 200), min_size=2, max_size=12)
    sc = st.lists(st.integers(1, 50), min_size=1, max_size=6)

    @settings(max_examples=150, deadline=None)
    @given(lists, sc, lists, sc)
    def check(c1, s1, c2, s2):
        r1, r2 = disjoint(c1, s1, [0, 1, 2]), disjoint(c2, s2, [3, 4, 5])
        m = merged(r1, r2)
        got = ranges(m)
        lib.ucc_coll_score_free(m)
        assert all(a < b for a, b, *_ in got)
        assert all(got[i][1] <= got[i + 1][0] for i in range(len(got) - 1))
        for x in range(0, 201):
            cover = [(s_, addr(f), k) for k, rs in enumerate((r1, r2)) for (a, b, s_, f) in rs if a <= x < b]
            hit = [g for g in got if g[0] <= x < g[1]]
            assert len(hit) == (1 if cover else 0), (x, cover, got)
            if cover:
                best = max(cover, key=lambda c: (c[0], -c[2]))
                assert hit[0][2] == best[0], (x, cover, hit)
                if len(cover) == 2 and cover[0][0] != cover[1][0]:
                    assert hit[0][3] == best[1]
                    loser = min(cover, key=lambda c: (c[0], -c[2]))
                    assert (loser[0], loser[1]) in [(f[0], f[1]) for f in hit[0][4]], (x, cover, hit)
    check()
