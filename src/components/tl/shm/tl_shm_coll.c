/* tl/shm collective algorithms, written as step programs (tl_shm_prog.h).
 * Algorithm catalogue follows what the reference's host TL offers
 * (tl_ucp: knomial / sra_knomial / ring / dbt / bruck / pairwise / linear /
 * neighbor, reference tl/ucp/<coll>/<coll>.c alg tables); every algorithm
 * here is a new formulation on top of the program builder. */
#include "tl_shm_coll.h"
#include "coll_patterns/knomial_tree.h"
#include "coll_patterns/sra_knomial.h"
#include "coll_patterns/bruck_alltoall.h"
#include "coll_patterns/ring.h"
#include "coll_patterns/double_binary_tree.h"

#define CHK(_x) do { st = (_x); if (ucc_unlikely(st != UCC_OK)) goto err; } while (0)
#define ARGS(_t) (&(_t)->super.bargs.args)
#define OFF(_p, _bytes) PTR_OFFSET(_p, _bytes)

static unsigned cfg_radix(const ucc_mrange_uint_t *r, size_t msg, ucc_memory_type_t mt, unsigned dflt, ucc_rank_t size)
{
    unsigned v = ucc_mrange_uint_get(r, msg, mt);
    if (v == UCC_UUNITS_AUTO || v < 2) v = dflt;
    if (v > size) v = size;
    if (v > 64) v = 64;
    return v < 2 ? 2 : v;
}

#define KN_DFLT(_t) (SHM_CTX((_t)->team)->cfg.kn_radix >= 2 ? SHM_CTX((_t)->team)->cfg.kn_radix : 4u) /* KN_RADIX: default of the radixes left on auto */
static unsigned cfg_radix_u(unsigned v, unsigned dflt, ucc_rank_t size)
{
    if (v == UCC_UUNITS_AUTO || v < 2) v = dflt;
    if (v > size) v = size;
    if (v > 64) v = 64;
    return v < 2 ? 2 : v;
}
/* ring position -> rank: the team's host order (RANKS_REORDERING) for collectives of the whole team, rank order otherwise */
static inline const ucc_rank_t *ring_order(const ucc_tl_shm_task_t *t)
{ return (t->team->ring_order && t->vsize == UCC_TL_TEAM_SIZE(t->team) && t->vrank == UCC_TL_TEAM_RANK(t->team) && !UCC_COLL_ARGS_ACTIVE_SET(ARGS(t))) ? t->team->ring_order : NULL; }
#define RING_AT(_ord, _i, _N) ((_ord) ? (_ord)[(_i) % (_N)] : (_i) % (_N))
#define NPOSTS(_v) (((_v) == UCC_UUNITS_AUTO) ? 0u : (_v))

/* ================================================================== */
/* barrier / fanin / fanout                                            */
/* ================================================================== */
static ucc_status_t prog_fanin(ucc_tl_shm_task_t *t, ucc_rank_t root, unsigned radix, unsigned step)
{
    ucc_kn_tree_t tr; ucc_status_t st = UCC_OK;
    ucc_kn_tree_init(&tr, t->vrank, t->vsize, root, radix);
    for (unsigned i = 0; i < tr.n_children; i++) CHK(shm_prog_recv(t, tr.children[i], NULL, 0, UCC_MEMORY_TYPE_HOST, step));
    CHK(shm_prog_wait(t));
    if (tr.parent != UCC_RANK_INVALID) { CHK(shm_prog_send(t, tr.parent, NULL, 0, UCC_MEMORY_TYPE_HOST, step)); CHK(shm_prog_wait(t)); }
err:
    return st;
}
static ucc_status_t prog_fanout(ucc_tl_shm_task_t *t, ucc_rank_t root, unsigned radix, unsigned step)
{
    ucc_kn_tree_t tr; ucc_status_t st = UCC_OK;
    ucc_kn_tree_init(&tr, t->vrank, t->vsize, root, radix);
    if (tr.parent != UCC_RANK_INVALID) { CHK(shm_prog_recv(t, tr.parent, NULL, 0, UCC_MEMORY_TYPE_HOST, step)); CHK(shm_prog_wait(t)); }
    for (unsigned i = 0; i < tr.n_children; i++) CHK(shm_prog_send(t, tr.children[i], NULL, 0, UCC_MEMORY_TYPE_HOST, step));
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_barrier_knomial(ucc_tl_shm_task_t *t)
{
    ucc_tl_shm_context_t *ctx = SHM_CTX(t->team);
    unsigned radix = cfg_radix(&ctx->cfg.barrier_kn_radix, 0, UCC_MEMORY_TYPE_HOST, KN_DFLT(t), t->vsize);
    ucc_status_t st;
    CHK(prog_fanin(t, 0, radix, 1));
    CHK(prog_fanout(t, 0, radix, 2));
err:
    return st;
}
ucc_status_t ucc_tl_shm_fanin_knomial(ucc_tl_shm_task_t *t)
{ return prog_fanin(t, (ucc_rank_t)ARGS(t)->root, cfg_radix_u(SHM_CTX(t->team)->cfg.fanin_kn_radix, cfg_radix(&SHM_CTX(t->team)->cfg.barrier_kn_radix, 0, UCC_MEMORY_TYPE_HOST, KN_DFLT(t), t->vsize), t->vsize), 1); }
ucc_status_t ucc_tl_shm_fanout_knomial(ucc_tl_shm_task_t *t)
{ return prog_fanout(t, (ucc_rank_t)ARGS(t)->root, cfg_radix_u(SHM_CTX(t->team)->cfg.fanout_kn_radix, cfg_radix(&SHM_CTX(t->team)->cfg.barrier_kn_radix, 0, UCC_MEMORY_TYPE_HOST, KN_DFLT(t), t->vsize), t->vsize), 1); }

/* ================================================================== */
/* bcast                                                               */
/* ================================================================== */
ucc_status_t ucc_tl_shm_bcast_knomial_prog(ucc_tl_shm_task_t *t, void *buf, size_t len, ucc_memory_type_t mt, ucc_rank_t root, unsigned radix)
{
    ucc_kn_tree_t tr; ucc_status_t st = UCC_OK;
    ucc_kn_tree_init(&tr, t->vrank, t->vsize, root, radix);
    if (tr.parent != UCC_RANK_INVALID) { CHK(shm_prog_recv(t, tr.parent, buf, len, mt, 1)); CHK(shm_prog_wait(t)); }
    for (unsigned i = 0; i < tr.n_children; i++) CHK(shm_prog_send(t, tr.children[i], buf, len, mt, 1));
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_bcast_knomial(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t len = a->src.info.count * ucc_dt_size(a->src.info.datatype);
    unsigned radix = cfg_radix(&SHM_CTX(t->team)->cfg.bcast_kn_radix, len, a->src.info.mem_type, KN_DFLT(t), t->vsize);
    ucc_rank_t root = (ucc_rank_t)a->root;
    if (UCC_COLL_ARGS_ACTIVE_SET(a)) root = (ucc_rank_t)(((int64_t)a->root - (int64_t)a->active_set.start) / a->active_set.stride);
    return ucc_tl_shm_bcast_knomial_prog(t, a->src.info.buffer, len, a->src.info.mem_type, root, radix);
}
/* scatter (binomial over blocks) followed by ring allgather: bandwidth optimal for large messages */
ucc_status_t ucc_tl_shm_bcast_sag(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t dts = ucc_dt_size(a->src.info.datatype), count = a->src.info.count;
    ucc_memory_type_t mt = a->src.info.mem_type;
    ucc_rank_t N = t->vsize, root = (ucc_rank_t)a->root, vr = (t->vrank + N - root) % N;
    char *buf = (char *)a->src.info.buffer;
    ucc_status_t st = UCC_OK;
    ucc_kn_tree_t tr;
    unsigned radix = cfg_radix(&SHM_CTX(t->team)->cfg.bcast_sag_kn_radix, count * dts, mt, SHM_CTX(t->team)->cfg.kn_radix >= 2 ? SHM_CTX(t->team)->cfg.kn_radix : 2u, N);
    if (UCC_COLL_ARGS_ACTIVE_SET(a)) return UCC_ERR_NOT_SUPPORTED;
    /* scatter: a rank receives the blocks of its whole subtree [vr, vr+span), forwards sub-ranges to children */
    ucc_kn_tree_init(&tr, t->vrank, N, root, radix);
    if (tr.parent != UCC_RANK_INVALID) {
        ucc_rank_t span = ucc_kn_subtree_size(vr, N, radix);
        size_t off = ucc_buffer_block_offset(count, N, vr), end = ucc_buffer_block_offset(count, N, vr + span - 1) + ucc_buffer_block_count(count, N, vr + span - 1);
        CHK(shm_prog_recv(t, tr.parent, buf + off * dts, (end - off) * dts, mt, 1)); CHK(shm_prog_wait(t));
    }
    for (unsigned i = 0; i < tr.n_children; i++) {
        ucc_rank_t cvr = (tr.children[i] + N - root) % N, span = ucc_kn_subtree_size(cvr, N, radix);
        size_t off = ucc_buffer_block_offset(count, N, cvr), end = ucc_buffer_block_offset(count, N, cvr + span - 1) + ucc_buffer_block_count(count, N, cvr + span - 1);
        CHK(shm_prog_send(t, tr.children[i], buf + off * dts, (end - off) * dts, mt, 1));
    }
    CHK(shm_prog_wait(t));
    /* ring allgather of the blocks in virtual-rank space */
    for (ucc_rank_t s = 0; s + 1 < N; s++) {
        ucc_rank_t sb = (vr + N - s) % N, rb = (vr + N - s - 1) % N;
        ucc_rank_t next = (t->vrank + 1) % N, prev = (t->vrank + N - 1) % N;
        CHK(shm_prog_send(t, next, buf + ucc_buffer_block_offset(count, N, sb) * dts, ucc_buffer_block_count(count, N, sb) * dts, mt, 2 + s));
        CHK(shm_prog_recv(t, prev, buf + ucc_buffer_block_offset(count, N, rb) * dts, ucc_buffer_block_count(count, N, rb) * dts, mt, 2 + s));
        CHK(shm_prog_wait(t));
    }
err:
    return st;
}
/* double binary tree: the two halves of the message travel down two complementary trees */
ucc_status_t ucc_tl_shm_bcast_dbt(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t dts = ucc_dt_size(a->src.info.datatype), count = a->src.info.count, c1 = count / 2, c2 = count - c1;
    ucc_memory_type_t mt = a->src.info.mem_type;
    ucc_rank_t N = t->vsize, root = (ucc_rank_t)a->root, vr = (t->vrank + N - root) % N;
    char *buf = (char *)a->src.info.buffer;
    ucc_dbt_t dbt; ucc_status_t st = UCC_OK;
    if (UCC_COLL_ARGS_ACTIVE_SET(a) || N < 3) return UCC_ERR_NOT_SUPPORTED;
    /* trees span virtual ranks 1..N-1; virtual rank 0 (the root) feeds both tree roots */
    if (vr == 0) {
        ucc_rank_t r1, r2; ucc_dbt_roots(N - 1, &r1, &r2);
        CHK(shm_prog_send(t, (r1 + 1 + root) % N, buf, c1 * dts, mt, 1));
        CHK(shm_prog_send(t, (r2 + 1 + root) % N, buf + c1 * dts, c2 * dts, mt, 2));
        CHK(shm_prog_wait(t));
        return UCC_OK;
    }
    ucc_dbt_init(&dbt, vr - 1, N - 1);
    /* one tree at a time: waiting for both parents before forwarding can dead-lock
     * (x may be y's parent in one tree and y's child in the other) */
    for (int k = 0; k < 2; k++) {
        ucc_rank_t parent = dbt.parent[k] == UCC_RANK_INVALID ? root : (dbt.parent[k] + 1 + root) % N;
        CHK(shm_prog_recv(t, parent, k == 0 ? buf : buf + c1 * dts, (k == 0 ? c1 : c2) * dts, mt, 1 + (unsigned)k));
        CHK(shm_prog_wait(t));
        for (int c = 0; c < 2; c++) if (dbt.children[k][c] != UCC_RANK_INVALID)
            CHK(shm_prog_send(t, (dbt.children[k][c] + 1 + root) % N, k == 0 ? buf : buf + c1 * dts, (k == 0 ? c1 : c2) * dts, mt, 1 + (unsigned)k));
        CHK(shm_prog_wait(t));
    }
err:
    return st;
}

/* ================================================================== */
/* allreduce                                                           */
/* ================================================================== */
/* recursive k-ing on buffer `acc` (count elements); scratch must hold (radix-1)*count elements */
static ucc_status_t prog_allreduce_knomial(ucc_tl_shm_task_t *t, void *acc, void *scratch, size_t count, ucc_memory_type_t mt, unsigned radix, unsigned step0)
{
    ucc_kn_pattern_t p; ucc_rank_t peers[64]; size_t len = count * ucc_dt_size(t->dt);
    ucc_status_t st = UCC_OK; unsigned step = step0;
    ucc_kn_pattern_init(&p, t->vrank, t->vsize, radix);
    if (p.type == UCC_KN_NODE_EXTRA) {
        CHK(shm_prog_send(t, p.partner, acc, len, mt, step)); CHK(shm_prog_wait(t));
        CHK(shm_prog_recv(t, p.partner, acc, len, mt, step + 1)); CHK(shm_prog_wait(t));
        return UCC_OK;
    }
    if (p.type == UCC_KN_NODE_PROXY) { /* up to radix-1 extras fold their data into this rank first */
        for (unsigned j = 0; j < p.n_extras; j++) CHK(shm_prog_recv(t, ucc_kn_extra(&p, j), OFF(scratch, j * len), len, mt, step));
        CHK(shm_prog_wait(t));
        for (unsigned j = 0; j < p.n_extras; j++) CHK(shm_prog_reduce(t, acc, acc, OFF(scratch, j * len), count, mt, 0));
    }
    step += 2;
    for (uint64_t dist = 1; dist < p.n_full; dist *= p.radix, step++) {
        unsigned n = ucc_kn_round_peers(&p, dist, peers);
        for (unsigned i = 0; i < n; i++) CHK(shm_prog_recv(t, peers[i], OFF(scratch, i * len), len, mt, step));
        for (unsigned i = 0; i < n; i++) CHK(shm_prog_send(t, peers[i], acc, len, mt, step));
        CHK(shm_prog_wait(t));
        for (unsigned i = 0; i < n; i++) CHK(shm_prog_reduce(t, acc, acc, OFF(scratch, i * len), count, mt, 0));
    }
    if (t->op == UCC_OP_AVG) CHK(shm_prog_reduce(t, acc, acc, NULL, count, mt, 1));
    if (p.type == UCC_KN_NODE_PROXY) {
        for (unsigned j = 0; j < p.n_extras; j++) CHK(shm_prog_send(t, ucc_kn_extra(&p, j), acc, len, mt, step0 + 1));
        CHK(shm_prog_wait(t));
    }
err:
    return st;
}
ucc_status_t ucc_tl_shm_allreduce_knomial(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t count = a->dst.info.count, len = count * ucc_dt_size(a->dst.info.datatype);
    ucc_memory_type_t mt = a->dst.info.mem_type;
    unsigned radix = cfg_radix(&SHM_CTX(t->team)->cfg.allreduce_kn_radix, len, mt, KN_DFLT(t), t->vsize);
    void *scratch; ucc_status_t st;
    CHK(shm_task_scratch(t, (size_t)(radix - 1) * len, mt, &scratch));
    if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, a->dst.info.buffer, a->src.info.buffer, len, mt, a->src.info.mem_type));
    CHK(prog_allreduce_knomial(t, a->dst.info.buffer, scratch, count, mt, radix, 1));
err:
    return st;
}

/* reduce-scatter ring on `buf` (count elems split in N blocks), result block (vrank+shift)%N complete on each rank */
static ucc_status_t prog_rs_ring(ucc_tl_shm_task_t *t, char *buf, void *scratch, size_t count, ucc_memory_type_t mt, int final_block_is_own, unsigned step0)
{
    ucc_rank_t N = t->vsize, r = t->vrank, next = ucc_ring_next(r, N), prev = ucc_ring_prev(r, N);   /* coll_patterns/ring.h */
    size_t dts = ucc_dt_size(t->dt); ucc_status_t st = UCC_OK;
    int sh = final_block_is_own ? 1 : 0; /* own: last received block == r; else == r+1 */
    for (ucc_rank_t s = 0; s + 1 < N; s++) {
        ucc_rank_t sb = ucc_ring_rs_send_block(r, N, s, !sh), rb = ucc_ring_rs_recv_block(r, N, s, !sh);
        size_t rc = ucc_buffer_block_count(count, N, rb);
        CHK(shm_prog_send(t, next, buf + ucc_buffer_block_offset(count, N, sb) * dts, ucc_buffer_block_count(count, N, sb) * dts, mt, step0 + s));
        CHK(shm_prog_recv(t, prev, scratch, rc * dts, mt, step0 + s));
        CHK(shm_prog_wait(t));
        CHK(shm_prog_reduce(t, buf + ucc_buffer_block_offset(count, N, rb) * dts, buf + ucc_buffer_block_offset(count, N, rb) * dts, scratch, rc, mt, 0));
    }
err:
    return st;
}
/* allgather ring where rank r starts owning block (r+own_shift)%N */
static ucc_status_t prog_ag_ring(ucc_tl_shm_task_t *t, char *buf, size_t count, size_t dts, ucc_memory_type_t mt, int own_shift, unsigned step0)
{
    ucc_rank_t N = t->vsize, r = t->vrank, next = ucc_ring_next(r, N), prev = ucc_ring_prev(r, N);
    ucc_status_t st = UCC_OK;
    for (ucc_rank_t s = 0; s + 1 < N; s++) {
        ucc_rank_t sb = (r + own_shift + 2 * N - s) % N, rb = (r + own_shift + 2 * N - s - 1) % N;
        CHK(shm_prog_send(t, next, buf + ucc_buffer_block_offset(count, N, sb) * dts, ucc_buffer_block_count(count, N, sb) * dts, mt, step0 + s));
        CHK(shm_prog_recv(t, prev, buf + ucc_buffer_block_offset(count, N, rb) * dts, ucc_buffer_block_count(count, N, rb) * dts, mt, step0 + s));
        CHK(shm_prog_wait(t));
    }
err:
    return st;
}
ucc_status_t ucc_tl_shm_allreduce_ring(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t count = a->dst.info.count, dts = ucc_dt_size(a->dst.info.datatype);
    ucc_memory_type_t mt = a->dst.info.mem_type; ucc_rank_t N = t->vsize, own = (t->vrank + 1) % N;
    char *dst = (char *)a->dst.info.buffer; void *scratch; ucc_status_t st;
    if (count < N) return UCC_ERR_NOT_SUPPORTED; /* fallback handles tiny vectors */
    CHK(shm_task_scratch(t, ucc_div_round_up(count, N) * dts, mt, &scratch));
    if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, dst, a->src.info.buffer, count * dts, mt, a->src.info.mem_type));
    CHK(prog_rs_ring(t, dst, scratch, count, mt, 0, 1));
    if (t->op == UCC_OP_AVG) CHK(shm_prog_reduce(t, dst + ucc_buffer_block_offset(count, N, own) * dts, dst + ucc_buffer_block_offset(count, N, own) * dts, NULL, ucc_buffer_block_count(count, N, own), mt, 1));
    CHK(prog_ag_ring(t, dst, count, dts, mt, 1, 1 + N));
err:
    return st;
}
/* one-sided "sliding window" allreduce (role of reference tl/ucp allreduce_sliding_window.c, which gets/puts registered
 * memory): rank r owns slice r; it READS window after window of that slice from every peer's source buffer (pointer / CMA),
 * reduces into its dst, and after a token exchange reads the other finished slices out of the peers' dst buffers. */
ucc_status_t ucc_tl_shm_allreduce_sliding_window(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t count = a->dst.info.count, dts = ucc_dt_size(a->dst.info.datatype), win = SHM_CTX(t->team)->cfg.allreduce_sliding_win_buf_size / dts, mo, mc;
    ucc_memory_type_t mt = a->dst.info.mem_type; ucc_rank_t N = t->vsize, r = t->vrank;
    char *dst = (char *)a->dst.info.buffer, *src = UCC_IS_INPLACE(*a) ? dst : (char *)a->src.info.buffer, *tok;
    void *sv, *scratch; uint64_t *out, *in; ucc_status_t st;
    if (mt != UCC_MEMORY_TYPE_HOST || (!UCC_IS_INPLACE(*a) && a->src.info.mem_type != mt) || count < N) return UCC_ERR_NOT_SUPPORTED;
    for (ucc_rank_t p = 0; p < N; p++) if (p != r && !ucc_tl_shm_can_get(t->team, ucc_ep_map_eval(t->vmap, p))) return UCC_ERR_NOT_SUPPORTED;
    if (!win) win = 1;
    CHK(shm_task_scratch(t, (size_t)N * 48, UCC_MEMORY_TYPE_HOST, &sv)); CHK(shm_task_scratch(t, win * dts, mt, &scratch));
    out = (uint64_t *)sv; in = out + 2; tok = (char *)(in + 2 * N); /* out: my {src, dst}; in[2p], in[2p+1]: peer p's; tok: 4N bytes */
    out[0] = (uint64_t)(uintptr_t)src; out[1] = (uint64_t)(uintptr_t)dst;
    for (ucc_rank_t p = 0; p < N; p++) { if (p == r) continue; CHK(shm_prog_send(t, p, out, 16, UCC_MEMORY_TYPE_HOST, 1)); CHK(shm_prog_recv(t, p, &in[2 * p], 16, UCC_MEMORY_TYPE_HOST, 1)); }
    CHK(shm_prog_wait(t));
    mo = ucc_buffer_block_offset(count, N, r); mc = ucc_buffer_block_count(count, N, r);
    /* scatter-reduce of my slice, one window at a time (the peers' base addresses arrive at run time: GET steps carry
     * "slot holding the base" + byte offset) */
    for (size_t wo = 0; wo < mc; wo += win) {
        size_t wc = ucc_min(win, mc - wo);
        if (src != dst) CHK(shm_prog_copy(t, dst + (mo + wo) * dts, src + (mo + wo) * dts, wc * dts, mt, mt));
        for (ucc_rank_t s = 1; s < N; s++) {
            ucc_rank_t p = (r + s) % N;
            CHK(shm_prog_get_off(t, p, scratch, &in[2 * p], (mo + wo) * dts, wc * dts, mt));
            CHK(shm_prog_reduce(t, dst + (mo + wo) * dts, dst + (mo + wo) * dts, scratch, wc, mt, 0));
        }
        if (t->op == UCC_OP_AVG) CHK(shm_prog_reduce(t, dst + (mo + wo) * dts, dst + (mo + wo) * dts, NULL, wc, mt, 1));
    }
    /* everybody's slice is final once all tokens of step 2 are here */
    for (ucc_rank_t p = 0; p < N; p++) { if (p == r) continue; tok[p] = 1; CHK(shm_prog_send(t, p, &tok[p], 1, UCC_MEMORY_TYPE_HOST, 2)); CHK(shm_prog_recv(t, p, &tok[N + p], 1, UCC_MEMORY_TYPE_HOST, 2)); }
    CHK(shm_prog_wait(t));
    for (ucc_rank_t s = 1; s < N; s++) {
        ucc_rank_t p = (r + s) % N; size_t po = ucc_buffer_block_offset(count, N, p), pc = ucc_buffer_block_count(count, N, p);
        CHK(shm_prog_get_off(t, p, dst + po * dts, &in[2 * p + 1], po * dts, pc * dts, mt));
    }
    for (ucc_rank_t p = 0; p < N; p++) { if (p == r) continue; CHK(shm_prog_send(t, p, &tok[2 * N + p], 1, UCC_MEMORY_TYPE_HOST, 3)); CHK(shm_prog_recv(t, p, &tok[3 * N + p], 1, UCC_MEMORY_TYPE_HOST, 3)); }
    CHK(shm_prog_wait(t));
err:
    return st;
}
/* SRA k-nomial (reference tl/ucp allreduce_sra_knomial.c, reduce_srg_knomial.c; ALLREDUCE_SRA_KN_RADIX, REDUCE_SRG_KN_RADIX): scatter-reduce by
 * recursive vector splitting - in the round of weight d (most significant digit first) the owned segment is cut into k parts, a rank keeps the
 * part its digit selects, sends the other k-1 to the ranks that differ in this digit only and reduces their k-1 contributions into its part -
 * then the rounds are walked backwards: as an allgather (allreduce), or as a gather towards the rank whose digits are all zero (reduce: the
 * root, ranks are rotated so that it is virtual rank 0).  Ranks beyond the largest power of k add their vector at a proxy first (and get the
 * result from it in the allreduce).  `w` holds the rank's contribution and, where one is due, the result; scratch: count + k elements. */
static ucc_status_t prog_sra_kn(ucc_tl_shm_task_t *t, char *w, void *scratch, size_t count, size_t dts, ucc_memory_type_t mt, unsigned radix, ucc_rank_t root, int gather_only /* 0: allgather, 1: gather to the root, 2: neither */)
{
    const ucc_rank_t N = t->vsize, vr = (ucc_rank_t)((t->vrank + N - root) % N);
    const size_t len = count * dts;
    ucc_kn_pattern_t p; ucc_status_t st = UCC_OK; unsigned step = 3, nsteps = 0;
    ucc_sra_seg_t segs[33]; uint64_t dists[32]; ucc_rank_t peers[64];
#define SRA_RANK(_v) ((ucc_rank_t)(((_v) + root) % N))
    const int avg_pre = t->op == UCC_OP_AVG && SHM_CTX(t->team)->cfg.reduce_avg_pre_op; /* REDUCE_AVG_PRE_OP: contributions are scaled, not the sum */
    ucc_kn_pattern_init(&p, vr, N, radix);
    if (avg_pre) CHK(shm_prog_reduce(t, w, w, NULL, count, mt, 1));
    if (p.type == UCC_KN_NODE_EXTRA) {
        CHK(shm_prog_send(t, SRA_RANK(p.partner), w, len, mt, 1)); CHK(shm_prog_wait(t));
        if (!gather_only) { CHK(shm_prog_recv(t, SRA_RANK(p.partner), w, len, mt, 2)); CHK(shm_prog_wait(t)); }
        return UCC_OK;
    }
    if (p.type == UCC_KN_NODE_PROXY)
        for (unsigned j = 0; j < p.n_extras; j++) { CHK(shm_prog_recv(t, SRA_RANK(ucc_kn_extra(&p, j)), scratch, len, mt, 1)); CHK(shm_prog_wait(t)); CHK(shm_prog_reduce(t, w, w, scratch, count, mt, 0)); }
    segs[0].off = 0; segs[0].cnt = count;
    for (uint64_t d = p.n_full / p.radix; d >= 1; d /= p.radix, step++, nsteps++) {
        const unsigned np = ucc_kn_round_peers(&p, d, peers), digit = (unsigned)((vr / d) % p.radix);
        const ucc_sra_seg_t seg = segs[nsteps], keep = ucc_sra_part(seg, p.radix, digit);   /* coll_patterns/sra_knomial.h */
        const size_t slot = ucc_buffer_block_count(seg.cnt, p.radix, 0) * dts;              /* the largest part */
        dists[nsteps] = d; segs[nsteps + 1] = keep;
        for (unsigned i = 0; i < np; i++) {
            const ucc_sra_seg_t give = ucc_sra_part(seg, p.radix, (unsigned)((peers[i] / d) % p.radix));
            CHK(shm_prog_send(t, SRA_RANK(peers[i]), w + give.off * dts, give.cnt * dts, mt, step));
            CHK(shm_prog_recv(t, SRA_RANK(peers[i]), (char *)scratch + i * slot, keep.cnt * dts, mt, step));
        }
        CHK(shm_prog_wait(t));
        for (unsigned i = 0; i < np; i++) CHK(shm_prog_reduce(t, w + keep.off * dts, w + keep.off * dts, (char *)scratch + i * slot, keep.cnt, mt, 0));
    }
    if (t->op == UCC_OP_AVG && !avg_pre) CHK(shm_prog_reduce(t, w + segs[nsteps].off * dts, w + segs[nsteps].off * dts, NULL, segs[nsteps].cnt, mt, 1));
    if (gather_only == 2) return UCC_OK; /* reduce_scatter: every rank keeps its part */
    for (int i = (int)nsteps - 1; i >= 0; i--, step++) { /* round i's peers hold the other parts of segs[i] */
        const unsigned np = ucc_kn_round_peers(&p, dists[i], peers), digit = (unsigned)((vr / dists[i]) % p.radix);
        if (gather_only && digit) { /* hand the collected segment to the member of the group that is closer to the root; done */
            CHK(shm_prog_send(t, SRA_RANK((ucc_rank_t)(vr - digit * dists[i])), w + segs[i + 1].off * dts, segs[i + 1].cnt * dts, mt, step));
            CHK(shm_prog_wait(t));
            return UCC_OK;
        }
        for (unsigned q = 0; q < np; q++) {
            const ucc_sra_seg_t theirs = ucc_sra_part(segs[i], p.radix, (unsigned)((peers[q] / dists[i]) % p.radix));
            if (!gather_only) CHK(shm_prog_send(t, SRA_RANK(peers[q]), w + segs[i + 1].off * dts, segs[i + 1].cnt * dts, mt, step));
            CHK(shm_prog_recv(t, SRA_RANK(peers[q]), w + theirs.off * dts, theirs.cnt * dts, mt, step));
        }
        CHK(shm_prog_wait(t));
    }
    if (p.type == UCC_KN_NODE_PROXY && !gather_only) { for (unsigned j = 0; j < p.n_extras; j++) CHK(shm_prog_send(t, SRA_RANK(ucc_kn_extra(&p, j)), w, len, mt, 2)); CHK(shm_prog_wait(t)); }
#undef SRA_RANK
err:
    return st;
}
static unsigned sra_radix(ucc_tl_shm_task_t *t, const ucc_mrange_uint_t *r, size_t len, ucc_memory_type_t mt)
{ return cfg_radix(r, len, mt, SHM_CTX(t->team)->cfg.kn_radix >= 2 ? SHM_CTX(t->team)->cfg.kn_radix : 2u, t->vsize); }
static int sra_fits(ucc_rank_t N, unsigned radix, size_t count) { uint64_t f = 1; while (f * radix <= N) f *= radix; return count >= f; } /* every rank of the full power gets a part */
ucc_status_t ucc_tl_shm_allreduce_sra(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t count = a->dst.info.count, dts = ucc_dt_size(a->dst.info.datatype), len = count * dts;
    ucc_memory_type_t mt = a->dst.info.mem_type;
    unsigned radix = sra_radix(t, &SHM_CTX(t->team)->cfg.allreduce_sra_kn_radix, len, mt);
    void *scratch; ucc_status_t st;
    if (!sra_fits(t->vsize, radix, count)) return UCC_ERR_NOT_SUPPORTED;
    CHK(shm_task_scratch(t, len + (size_t)radix * dts, mt, &scratch)); /* k-1 slots of ceil(cnt / k) elements can exceed cnt by up to k-1 */
    if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, a->dst.info.buffer, a->src.info.buffer, len, mt, a->src.info.mem_type));
    CHK(prog_sra_kn(t, (char *)a->dst.info.buffer, scratch, count, dts, mt, radix, 0, 0));
err:
    return st;
}
/* reduce down one tree / bcast up the other on two halves (double binary tree) */
ucc_status_t ucc_tl_shm_allreduce_dbt(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    size_t count = a->dst.info.count, dts = ucc_dt_size(a->dst.info.datatype), c[2], o[2];
    ucc_memory_type_t mt = a->dst.info.mem_type;
    char *dst = (char *)a->dst.info.buffer; void *scratch;
    ucc_dbt_t dbt; ucc_status_t st; ucc_rank_t N = t->vsize;
    if (N < 2 || count < 2) return UCC_ERR_NOT_SUPPORTED;
    c[0] = count / 2; c[1] = count - c[0]; o[0] = 0; o[1] = c[0];
    CHK(shm_task_scratch(t, 2 * c[1] * dts, mt, &scratch));
    if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, dst, a->src.info.buffer, count * dts, mt, a->src.info.mem_type));
    ucc_dbt_init(&dbt, t->vrank, N);
    for (int k = 0; k < 2; k++) { /* reduce towards the root of tree k */
        unsigned nc = 0;
        for (int ch = 0; ch < 2; ch++) if (dbt.children[k][ch] != UCC_RANK_INVALID) { CHK(shm_prog_recv(t, dbt.children[k][ch], OFF(scratch, nc * c[1] * dts), c[k] * dts, mt, 1 + (unsigned)k)); nc++; }
        CHK(shm_prog_wait(t));
        for (unsigned i = 0; i < nc; i++) CHK(shm_prog_reduce(t, dst + o[k] * dts, dst + o[k] * dts, OFF(scratch, i * c[1] * dts), c[k], mt, 0));
        if (dbt.parent[k] != UCC_RANK_INVALID) { CHK(shm_prog_send(t, dbt.parent[k], dst + o[k] * dts, c[k] * dts, mt, 1 + (unsigned)k)); CHK(shm_prog_wait(t)); }
        else if (t->op == UCC_OP_AVG) CHK(shm_prog_reduce(t, dst + o[k] * dts, dst + o[k] * dts, NULL, c[k], mt, 1));
    }
    for (int k = 0; k < 2; k++) { /* broadcast back */
        if (dbt.parent[k] != UCC_RANK_INVALID) { CHK(shm_prog_recv(t, dbt.parent[k], dst + o[k] * dts, c[k] * dts, mt, 3 + (unsigned)k)); CHK(shm_prog_wait(t)); }
        for (int ch = 0; ch < 2; ch++) if (dbt.children[k][ch] != UCC_RANK_INVALID) CHK(shm_prog_send(t, dbt.children[k][ch], dst + o[k] * dts, c[k] * dts, mt, 3 + (unsigned)k));
        CHK(shm_prog_wait(t));
    }
err:
    return st;
}

/* ================================================================== */
/* reduce                                                              */
/* ================================================================== */
ucc_status_t ucc_tl_shm_reduce_knomial(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    ucc_rank_t root = (ucc_rank_t)a->root; int is_root = t->vrank == root;
    size_t count = is_root ? a->dst.info.count : a->src.info.count;
    ucc_datatype_t dt = is_root ? a->dst.info.datatype : a->src.info.datatype;
    ucc_memory_type_t mt = is_root ? a->dst.info.mem_type : a->src.info.mem_type;
    size_t len = count * ucc_dt_size(dt);
    unsigned radix = cfg_radix(&SHM_CTX(t->team)->cfg.reduce_kn_radix, len, mt, KN_DFLT(t), t->vsize);
    ucc_kn_tree_t tr; void *acc, *scratch = NULL; const void *first; ucc_status_t st;
    ucc_kn_tree_init(&tr, t->vrank, t->vsize, root, radix);
    first = (is_root && UCC_IS_INPLACE(*a)) ? a->dst.info.buffer : a->src.info.buffer;
    if (tr.n_children) CHK(shm_task_scratch(t, (size_t)tr.n_children * len, mt, &scratch));
    if (is_root) acc = a->dst.info.buffer;
    else if (tr.n_children) CHK(shm_task_scratch(t, len, mt, &acc));
    else acc = (void *)first; /* leaf: send the source as is */
    for (unsigned i = 0; i < tr.n_children; i++) CHK(shm_prog_recv(t, tr.children[i], OFF(scratch, i * len), len, mt, 1));
    CHK(shm_prog_wait(t));
    for (unsigned i = 0; i < tr.n_children; i++) CHK(shm_prog_reduce(t, acc, i == 0 ? first : acc, OFF(scratch, i * len), count, mt, 0));
    if (is_root) {
        if (!tr.n_children) CHK(shm_prog_copy(t, acc, first, len, mt, mt));
        if (t->op == UCC_OP_AVG) CHK(shm_prog_reduce(t, acc, acc, NULL, count, mt, 1));
    } else { CHK(shm_prog_send(t, tr.parent, acc, len, mt, 1)); CHK(shm_prog_wait(t)); }
err:
    return st;
}
/* scatter-reduce followed by a gather to the root, both k-nomial (reference reduce_srg_knomial.c): bandwidth optimal */
ucc_status_t ucc_tl_shm_reduce_srg(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    ucc_rank_t root = (ucc_rank_t)a->root; int is_root = t->vrank == root;
    size_t count = is_root ? a->dst.info.count : a->src.info.count;
    ucc_datatype_t dt = is_root ? a->dst.info.datatype : a->src.info.datatype;
    ucc_memory_type_t mt = is_root ? a->dst.info.mem_type : a->src.info.mem_type;
    size_t dts = ucc_dt_size(dt);
    unsigned radix = sra_radix(t, &SHM_CTX(t->team)->cfg.reduce_srg_kn_radix, count * dts, mt);
    void *work, *scratch; ucc_status_t st;
    if (!sra_fits(t->vsize, radix, count)) return UCC_ERR_NOT_SUPPORTED;
    CHK(shm_task_scratch(t, (count + radix) * dts, mt, &scratch));
    if (is_root) { work = a->dst.info.buffer; if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, work, a->src.info.buffer, count * dts, mt, a->src.info.mem_type)); }
    else { CHK(shm_task_scratch(t, count * dts, mt, &work)); CHK(shm_prog_copy(t, work, a->src.info.buffer, count * dts, mt, mt)); }
    CHK(prog_sra_kn(t, (char *)work, scratch, count, dts, mt, radix, root, 1));
err:
    return st;
}
ucc_status_t ucc_tl_shm_reduce_dbt(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    ucc_rank_t root = (ucc_rank_t)a->root, N = t->vsize, vr = (t->vrank + N - root) % N; int is_root = vr == 0;
    size_t count = is_root ? a->dst.info.count : a->src.info.count, c[2], o[2];
    ucc_datatype_t dt = is_root ? a->dst.info.datatype : a->src.info.datatype;
    ucc_memory_type_t mt = is_root ? a->dst.info.mem_type : a->src.info.mem_type;
    size_t dts = ucc_dt_size(dt);
    void *work, *scratch; char *w; ucc_dbt_t dbt; ucc_status_t st;
    if (N < 3 || count < 2) return UCC_ERR_NOT_SUPPORTED;
    c[0] = count / 2; c[1] = count - c[0]; o[0] = 0; o[1] = c[0];
    CHK(shm_task_scratch(t, 2 * c[1] * dts, mt, &scratch));
    if (is_root) { /* root (virtual 0) sits above both trees built over virtual ranks 1..N-1 */
        ucc_rank_t r1, r2; ucc_dbt_roots(N - 1, &r1, &r2);
        work = a->dst.info.buffer; w = (char *)work;
        if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, work, a->src.info.buffer, count * dts, mt, a->src.info.mem_type));
        CHK(shm_prog_recv(t, (r1 + 1 + root) % N, scratch, c[0] * dts, mt, 1));
        CHK(shm_prog_recv(t, (r2 + 1 + root) % N, OFF(scratch, c[1] * dts), c[1] * dts, mt, 2));
        CHK(shm_prog_wait(t));
        CHK(shm_prog_reduce(t, w, w, scratch, c[0], mt, t->op == UCC_OP_AVG));
        CHK(shm_prog_reduce(t, w + o[1] * dts, w + o[1] * dts, OFF(scratch, c[1] * dts), c[1], mt, t->op == UCC_OP_AVG));
        return UCC_OK;
    }
    CHK(shm_task_scratch(t, count * dts, mt, &work)); w = (char *)work;
    CHK(shm_prog_copy(t, work, a->src.info.buffer, count * dts, mt, mt));
    ucc_dbt_init(&dbt, vr - 1, N - 1);
    for (int k = 0; k < 2; k++) {
        unsigned nc = 0; ucc_rank_t parent = dbt.parent[k] == UCC_RANK_INVALID ? root : (dbt.parent[k] + 1 + root) % N;
        for (int ch = 0; ch < 2; ch++) if (dbt.children[k][ch] != UCC_RANK_INVALID) { CHK(shm_prog_recv(t, (dbt.children[k][ch] + 1 + root) % N, OFF(scratch, nc * c[1] * dts), c[k] * dts, mt, 1 + (unsigned)k)); nc++; }
        CHK(shm_prog_wait(t));
        for (unsigned i = 0; i < nc; i++) CHK(shm_prog_reduce(t, w + o[k] * dts, w + o[k] * dts, OFF(scratch, i * c[1] * dts), c[k], mt, 0));
        CHK(shm_prog_send(t, parent, w + o[k] * dts, c[k] * dts, mt, 1 + (unsigned)k)); CHK(shm_prog_wait(t));
    }
err:
    return st;
}

/* ================================================================== */
/* reduce_scatter(v)                                                   */
/* ================================================================== */
/* ring reduce_scatter(v): a block is reduced into exactly once on every rank (own contribution + the partial sum that arrived), so the
 * contributions are read where they are (`src`), partial sums live in `work` (touched only for the N-2 blocks that pass through) and the last
 * step reduces straight into `out` (the place of the rank's result): no staging copy of the whole source, no final copy. */
static ucc_status_t prog_rsv_ring(ucc_tl_shm_task_t *t, char *work, const char *src, char *out, void *scratch, const size_t *cnt, const size_t *off, ucc_memory_type_t mt, int bidir, size_t maxc)
{
    const ucc_rank_t *ord = ring_order(t);
    ucc_rank_t N = t->vsize, r = t->vrank, p = ord ? t->team->ring_pos[r] : r, next = RING_AT(ord, p + 1, N), prev = RING_AT(ord, p + N - 1, N);
    size_t dts = ucc_dt_size(t->dt); ucc_status_t st = UCC_OK;
    /* bidirectional (reference reduce_scatter_ring.c, REDUCE_SCATTER_RING_BIDIRECTIONAL): the first half of every block is reduced around the
     * ring r -> r+1, the second half around the inverted ring r -> r-1, both at once; a step moves half the bytes per direction */
#define LO(_b) (bidir ? cnt[_b] - cnt[_b] / 2 : cnt[_b])
#define HI(_b) (cnt[_b] / 2)
    char *scr2 = (char *)scratch + (maxc - maxc / 2) * dts;
    for (ucc_rank_t s = 0; s + 1 < N; s++) {
        ucc_rank_t sb = RING_AT(ord, p + 2 * N - s - 1, N), rb = RING_AT(ord, p + 2 * N - s - 2, N), sb2 = RING_AT(ord, p + s + 1, N), rb2 = RING_AT(ord, p + s + 2, N);
        const char *from = s == 0 ? src : work;                  /* step 0 forwards an own block, later steps the partial sum of the step before */
        const int last = s + 2 == N;                             /* rb == rb2 == r */
        CHK(shm_prog_send(t, next, from + off[sb] * dts, LO(sb) * dts, mt, 1 + s));
        CHK(shm_prog_recv(t, prev, scratch, LO(rb) * dts, mt, 1 + s));
        if (bidir) {
            CHK(shm_prog_send(t, prev, from + (off[sb2] + LO(sb2)) * dts, HI(sb2) * dts, mt, N + 1 + s));
            CHK(shm_prog_recv(t, next, scr2, HI(rb2) * dts, mt, N + 1 + s));
        }
        CHK(shm_prog_wait(t));
        CHK(shm_prog_reduce(t, last ? out : work + off[rb] * dts, src + off[rb] * dts, scratch, LO(rb), mt, 0));
        if (bidir && HI(rb2)) CHK(shm_prog_reduce(t, last ? out + LO(rb2) * dts : work + (off[rb2] + LO(rb2)) * dts, src + (off[rb2] + LO(rb2)) * dts, scr2, HI(rb2), mt, 0));
    }
#undef LO
#undef HI
err:
    return st;
}
static ucc_status_t reduce_scatter_common(ucc_tl_shm_task_t *t, int is_v)
{
    ucc_coll_args_t *a = ARGS(t);
    ucc_rank_t N = t->vsize, r = t->vrank; int inplace = UCC_IS_INPLACE(*a);
    ucc_datatype_t dt = is_v ? a->dst.info_v.datatype : a->dst.info.datatype;
    ucc_memory_type_t mt = is_v ? a->dst.info_v.mem_type : a->dst.info.mem_type;
    size_t dts = ucc_dt_size(dt), total = 0, maxc = 0, *cnt, *off;
    void *work, *scratch, *dstbuf = is_v ? a->dst.info_v.buffer : a->dst.info.buffer; const char *src; char *out; ucc_status_t st;
    cnt = (size_t *)malloc(2 * N * sizeof(size_t)); if (!cnt) return UCC_ERR_NO_MEMORY;
    off = cnt + N; t->host_copy = cnt;
    for (ucc_rank_t i = 0; i < N; i++) {
        if (is_v) cnt[i] = ucc_coll_args_get_count(a, a->dst.info_v.counts, i);
        else { size_t tot = inplace ? a->dst.info.count : a->dst.info.count * N; cnt[i] = ucc_buffer_block_count(tot, N, i); }
        off[i] = total; total += cnt[i]; if (cnt[i] > maxc) maxc = cnt[i];
    }
    CHK(shm_task_scratch(t, maxc * dts, mt, &scratch));
    if (inplace) { work = dstbuf; src = (const char *)dstbuf; out = (char *)dstbuf + off[r] * dts; }
    else {
        CHK(shm_task_scratch(t, total * dts, mt, &work));
        out = (char *)dstbuf;
        if (a->src.info.mem_type == mt) src = (const char *)a->src.info.buffer;
        else { CHK(shm_prog_copy(t, work, a->src.info.buffer, total * dts, mt, a->src.info.mem_type)); src = (const char *)work; } /* contributions in another memory type: staged */
    }
    CHK(prog_rsv_ring(t, (char *)work, src, out, scratch, cnt, off, mt, N > 2 && mt == UCC_MEMORY_TYPE_HOST /* device buffers: one reduce launch per step */ &&
                      (is_v ? SHM_CTX(t->team)->cfg.reduce_scatterv_ring_bidirectional : SHM_CTX(t->team)->cfg.reduce_scatter_ring_bidirectional), maxc));
    if (t->op == UCC_OP_AVG) CHK(shm_prog_reduce(t, out, out, NULL, cnt[r], mt, 1));
err:
    return st;
}
ucc_status_t ucc_tl_shm_reduce_scatter_ring(ucc_tl_shm_task_t *t) { return reduce_scatter_common(t, 0); }
ucc_status_t ucc_tl_shm_reduce_scatterv_ring(ucc_tl_shm_task_t *t) { return reduce_scatter_common(t, 1); }
/* k-nomial scatter-reduce on teams of k^m ranks; recursive halving (power-of-two teams, any block sizes): log2(N) steps, each exchanging half of
 * the remaining range */
ucc_status_t ucc_tl_shm_reduce_scatter_knomial(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t);
    ucc_rank_t N = t->vsize, r = t->vrank, lo = 0, hi = N; int inplace = UCC_IS_INPLACE(*a);
    ucc_memory_type_t mt = a->dst.info.mem_type;
    size_t dts = ucc_dt_size(a->dst.info.datatype), total = inplace ? a->dst.info.count : a->dst.info.count * N;
    void *work, *scratch; char *w; ucc_status_t st; unsigned step = 1;
    unsigned radix = sra_radix(t, &SHM_CTX(t->team)->cfg.reduce_scatter_kn_radix, total * dts, mt); uint64_t full = 1;
    while (full * radix <= N) full *= radix;
    if (full == N && total % N == 0) {
        /* REDUCE_SCATTER_KN_RADIX (reference reduce_scatter_knomial.c): on a team of k^m ranks with equal blocks the scatter-reduce half of the
         * SRA program leaves rank r with block r (the digits of r, most significant first, select the part of every round) */
        CHK(shm_task_scratch(t, (total + radix) * dts, mt, &scratch));
        if (inplace) work = a->dst.info.buffer;
        else { CHK(shm_task_scratch(t, total * dts, mt, &work)); CHK(shm_prog_copy(t, work, a->src.info.buffer, total * dts, mt, a->src.info.mem_type)); }
        CHK(prog_sra_kn(t, (char *)work, scratch, total, dts, mt, radix, 0, 2));
        if (!inplace) CHK(shm_prog_copy(t, a->dst.info.buffer, (char *)work + (size_t)r * (total / N) * dts, total / N * dts, mt, mt));
        return UCC_OK;
    }
    if (!ucc_is_pow2(N)) return UCC_ERR_NOT_SUPPORTED;
    CHK(shm_task_scratch(t, (total / 2 + N) * dts, mt, &scratch));
    if (inplace) work = a->dst.info.buffer;
    else { CHK(shm_task_scratch(t, total * dts, mt, &work)); CHK(shm_prog_copy(t, work, a->src.info.buffer, total * dts, mt, a->src.info.mem_type)); }
    w = (char *)work;
    while (hi - lo > 1) { /* block range [lo,hi) is what this rank is still responsible for */
        ucc_rank_t mid = (lo + hi) / 2, peer, klo, khi, slo, shi;
        if (r < mid) { peer = r + (mid - lo); klo = lo; khi = mid; slo = mid; shi = hi; } else { peer = r - (mid - lo); klo = mid; khi = hi; slo = lo; shi = mid; }
        size_t ko = ucc_buffer_block_offset(total, N, klo), kc = ucc_buffer_block_offset(total, N, khi - 1) + ucc_buffer_block_count(total, N, khi - 1) - ko;
        size_t so = ucc_buffer_block_offset(total, N, slo), sc = ucc_buffer_block_offset(total, N, shi - 1) + ucc_buffer_block_count(total, N, shi - 1) - so;
        CHK(shm_prog_send(t, peer, w + so * dts, sc * dts, mt, step)); CHK(shm_prog_recv(t, peer, scratch, kc * dts, mt, step)); CHK(shm_prog_wait(t));
        CHK(shm_prog_reduce(t, w + ko * dts, w + ko * dts, scratch, kc, mt, 0));
        lo = klo; hi = khi; step++;
    }
    { size_t o = ucc_buffer_block_offset(total, N, r), c = ucc_buffer_block_count(total, N, r);
      if (t->op == UCC_OP_AVG) CHK(shm_prog_reduce(t, w + o * dts, w + o * dts, NULL, c, mt, 1));
      if (!inplace) CHK(shm_prog_copy(t, a->dst.info.buffer, w + o * dts, c * dts, mt, mt)); }
err:
    return st;
}

/* ================================================================== */
/* allgather(v)                                                        */
/* ================================================================== */
static ucc_status_t ag_layout(ucc_tl_shm_task_t *t, int is_v, size_t **cnt_p, size_t **off_p, size_t *dts, ucc_memory_type_t *mt, char **dst)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize;
    size_t *cnt = (size_t *)malloc(2 * N * sizeof(size_t)), *off;
    if (!cnt) return UCC_ERR_NO_MEMORY;
    off = cnt + N; t->host_copy = cnt;
    if (is_v) {
        *dts = ucc_dt_size(a->dst.info_v.datatype); *mt = a->dst.info_v.mem_type; *dst = (char *)a->dst.info_v.buffer;
        for (ucc_rank_t i = 0; i < N; i++) { cnt[i] = ucc_coll_args_get_count(a, a->dst.info_v.counts, i) * *dts; off[i] = ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, i) * *dts; }
    } else {
        size_t blk = (a->dst.info.count / N) * ucc_dt_size(a->dst.info.datatype);
        *dts = ucc_dt_size(a->dst.info.datatype); *mt = a->dst.info.mem_type; *dst = (char *)a->dst.info.buffer;
        for (ucc_rank_t i = 0; i < N; i++) { cnt[i] = blk; off[i] = i * blk; }
    }
    *cnt_p = cnt; *off_p = off;
    return UCC_OK;
}
static ucc_status_t ag_own_block(ucc_tl_shm_task_t *t, char *dst, const size_t *cnt, const size_t *off, ucc_memory_type_t mt)
{
    ucc_coll_args_t *a = ARGS(t);
    if (UCC_IS_INPLACE(*a)) return UCC_OK;
    return shm_prog_copy(t, dst + off[t->vrank], a->src.info.buffer, cnt[t->vrank], mt, a->src.info.mem_type);
}
static ucc_status_t allgather_ring_common(ucc_tl_shm_task_t *t, int is_v)
{
    size_t *cnt, *off, dts; ucc_memory_type_t mt; char *dst; ucc_status_t st;
    const ucc_rank_t *ord = ring_order(t);
    ucc_rank_t N = t->vsize, r = t->vrank, p = ord ? t->team->ring_pos[r] : r, next = RING_AT(ord, p + 1, N), prev = RING_AT(ord, p + N - 1, N);
    CHK(ag_layout(t, is_v, &cnt, &off, &dts, &mt, &dst));
    CHK(ag_own_block(t, dst, cnt, off, mt));
    for (ucc_rank_t s = 0; s + 1 < N; s++) {
        ucc_rank_t sb = RING_AT(ord, p + N - s, N), rb = RING_AT(ord, p + N - s - 1, N);
        CHK(shm_prog_send(t, next, dst + off[sb], cnt[sb], mt, 1 + s)); CHK(shm_prog_recv(t, prev, dst + off[rb], cnt[rb], mt, 1 + s)); CHK(shm_prog_wait(t));
    }
err:
    return st;
}
ucc_status_t ucc_tl_shm_allgather_ring(ucc_tl_shm_task_t *t) { return allgather_ring_common(t, 0); }
ucc_status_t ucc_tl_shm_allgatherv_ring(ucc_tl_shm_task_t *t) { return allgather_ring_common(t, 1); }
static ucc_status_t allgather_linear_common(ucc_tl_shm_task_t *t, int is_v, unsigned batch)
{
    size_t *cnt, *off, dts; ucc_memory_type_t mt; char *dst; ucc_status_t st; ucc_rank_t N = t->vsize, r = t->vrank;
    CHK(ag_layout(t, is_v, &cnt, &off, &dts, &mt, &dst));
    CHK(ag_own_block(t, dst, cnt, off, mt));
    for (ucc_rank_t s = 1; s < N; s++) {
        ucc_rank_t to = (r + s) % N, from = (r + N - s) % N;
        CHK(shm_prog_recv(t, from, dst + off[from], cnt[from], mt, 1)); CHK(shm_prog_send(t, to, dst + off[r], cnt[r], mt, 1));
        if (batch && s % batch == 0) CHK(shm_prog_wait(t));
    }
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_allgather_linear(ucc_tl_shm_task_t *t) { return allgather_linear_common(t, 0, 0); }
ucc_status_t ucc_tl_shm_allgather_batched(ucc_tl_shm_task_t *t)
{ unsigned np = SHM_CTX(t->team)->cfg.allgather_batched_num_posts; return allgather_linear_common(t, 0, np == UCC_UUNITS_AUTO ? 4 : np); }
ucc_status_t ucc_tl_shm_allgatherv_linear(ucc_tl_shm_task_t *t) { return allgather_linear_common(t, 1, 0); }
/* neighbor exchange (even team sizes): N/2 steps, pairs alternate left/right */
ucc_status_t ucc_tl_shm_allgather_neighbor(ucc_tl_shm_task_t *t)
{
    size_t *cnt, *off, dts; ucc_memory_type_t mt; char *dst; ucc_status_t st; ucc_rank_t N = t->vsize, r = t->vrank;
    ucc_rank_t nb[2], rdf[2]; int od[2];
    if (N % 2) return UCC_ERR_NOT_SUPPORTED;
    CHK(ag_layout(t, 0, &cnt, &off, &dts, &mt, &dst));
    CHK(ag_own_block(t, dst, cnt, off, mt));
    if (r % 2 == 0) { nb[0] = (r + 1) % N; nb[1] = (r + N - 1) % N; rdf[0] = r; rdf[1] = r; od[0] = 2; od[1] = -2; }
    else { nb[0] = (r + N - 1) % N; nb[1] = (r + 1) % N; rdf[0] = nb[0]; rdf[1] = nb[0]; od[0] = -2; od[1] = 2; }
    /* step 0: exchange own block with neighbor 0 */
    CHK(shm_prog_send(t, nb[0], dst + off[r], cnt[r], mt, 1)); CHK(shm_prog_recv(t, nb[0], dst + off[nb[0]], cnt[nb[0]], mt, 1)); CHK(shm_prog_wait(t));
    { ucc_rank_t send_from = r % 2 == 0 ? r : nb[0]; /* lower index of the pair just completed */
      for (ucc_rank_t i = 1; i < N / 2; i++) {
        int p = (int)(i % 2);
        ucc_rank_t recv_from = (rdf[p] + (ucc_rank_t)((int)N + od[p])) % N;
        rdf[p] = recv_from;
        /* blocks are exchanged in pairs (send_from, send_from+1) */
        CHK(shm_prog_send(t, nb[p], dst + off[send_from], cnt[send_from] + cnt[(send_from + 1) % N], mt, 1 + i));
        CHK(shm_prog_recv(t, nb[p], dst + off[recv_from], cnt[recv_from] + cnt[(recv_from + 1) % N], mt, 1 + i));
        CHK(shm_prog_wait(t));
        send_from = recv_from;
      } }
err:
    return st;
}
/* Bruck: ceil(log2 N) steps through a rotated scratch copy */
ucc_status_t ucc_tl_shm_allgather_bruck(ucc_tl_shm_task_t *t)
{
    size_t *cnt, *off, dts, blk; ucc_memory_type_t mt; char *dst; ucc_status_t st; ucc_rank_t N = t->vsize, r = t->vrank;
    void *scr; char *s; unsigned step = 1; ucc_coll_args_t *a = ARGS(t);
    CHK(ag_layout(t, 0, &cnt, &off, &dts, &mt, &dst)); blk = cnt[0];
    CHK(shm_task_scratch(t, blk * N, mt, &scr)); s = (char *)scr;
    CHK(shm_prog_copy(t, s, UCC_IS_INPLACE(*a) ? (void *)(dst + off[r]) : a->src.info.buffer, blk, mt, UCC_IS_INPLACE(*a) ? mt : a->src.info.mem_type));
    for (ucc_rank_t d = 1; d < N; d *= 2, step++) {
        ucc_rank_t nblk = (2 * d <= N) ? d : N - d, to = (r + N - d) % N, from = (r + d) % N;
        CHK(shm_prog_send(t, to, s, nblk * blk, mt, step)); CHK(shm_prog_recv(t, from, s + d * blk, nblk * blk, mt, step)); CHK(shm_prog_wait(t));
    }
    /* scratch block i holds data of rank (r+i)%N */
    CHK(shm_prog_copy(t, dst + (size_t)r * blk, s, (N - r) * blk, mt, mt));
    if (r) CHK(shm_prog_copy(t, dst, s + (size_t)(N - r) * blk, (size_t)r * blk, mt, mt));
err:
    return st;
}
/* sparbit: log steps with distances N/2, N/4.. ; data locality friendly — same data motion as bruck without rotation for pow2 */
ucc_status_t ucc_tl_shm_allgather_sparbit(ucc_tl_shm_task_t *t)
{
    size_t *cnt, *off, dts, blk; ucc_memory_type_t mt; char *dst; ucc_status_t st; ucc_rank_t N = t->vsize, r = t->vrank; unsigned step = 1;
    if (!ucc_is_pow2(N)) return UCC_ERR_NOT_SUPPORTED;
    CHK(ag_layout(t, 0, &cnt, &off, &dts, &mt, &dst)); blk = cnt[0];
    CHK(ag_own_block(t, dst, cnt, off, mt));
    /* step k (distance d = N>>(k+1)): send every block b owned so far (b = r + j*2d) to r+d, receive from r-d */
    for (ucc_rank_t d = N / 2; d >= 1; d /= 2, step++) {
        ucc_rank_t to = (r + d) % N, from = (r + N - d) % N, stride = 2 * d;
        for (ucc_rank_t j = 0; j < N / stride; j++) {
            ucc_rank_t sb = (r + N - j * stride) % N, rb = (from + N - j * stride) % N;
            CHK(shm_prog_send(t, to, dst + (size_t)sb * blk, blk, mt, step * 64 + j)); CHK(shm_prog_recv(t, from, dst + (size_t)rb * blk, blk, mt, step * 64 + j));
        }
        CHK(shm_prog_wait(t));
        if (d == 1) break;
    }
err:
    return st;
}
/* allgather(v) by recursive k-ing (role of reference tl/ucp allgather_knomial.c / allgatherv knomial; ALLGATHER_KN_RADIX): log_k(N) rounds over
 * the digits of the rank, least significant first.  Before the round of weight `dist` a rank holds the blocks of the `dist` ranks that share
 * its higher digits; it swaps that set with the k-1 ranks that differ in this digit only.  Blocks keep their (arbitrary) displacements, so a
 * round is a set of block-sized messages.  Ranks beyond the largest power of the radix hand their block to a proxy first and get the complete
 * vector from it at the end (coll_patterns/knomial_tree.h: EXTRA / PROXY; a proxy serves up to k-1 extras). */
static ucc_status_t allgather_knomial_common(ucc_tl_shm_task_t *t, int is_v)
{
    size_t *cnt, *off, dts, tot = 0; ucc_memory_type_t mt; char *dst; ucc_status_t st;
    ucc_rank_t N = t->vsize, r = t->vrank, peers[64]; ucc_kn_pattern_t p; unsigned round = 1, radix;
    if ((uint64_t)N * 24 > 65000) return UCC_ERR_NOT_SUPPORTED; /* message ids: (round, block) must fit 16 bits */
    CHK(ag_layout(t, is_v, &cnt, &off, &dts, &mt, &dst));
    for (ucc_rank_t i = 0; i < N; i++) tot += cnt[i];
    radix = cfg_radix(&SHM_CTX(t->team)->cfg.allgather_kn_radix, tot, mt, SHM_CTX(t->team)->cfg.kn_radix >= 2 ? SHM_CTX(t->team)->cfg.kn_radix : 2u, N);
    ucc_kn_pattern_init(&p, r, N, radix);
    CHK(ag_own_block(t, dst, cnt, off, mt));
#define AGV_ID(_round, _b) ((unsigned)((_round) * N + (_b)))
    if (p.type == UCC_KN_NODE_EXTRA) {
        CHK(shm_prog_send(t, p.partner, dst + off[r], cnt[r], mt, AGV_ID(0, r))); CHK(shm_prog_wait(t));
        for (ucc_rank_t b = 0; b < N; b++) if (b != r) CHK(shm_prog_recv(t, p.partner, dst + off[b], cnt[b], mt, AGV_ID(23, b)));
        CHK(shm_prog_wait(t));
        return UCC_OK;
    }
    if (p.type == UCC_KN_NODE_PROXY) {
        for (unsigned j = 0; j < p.n_extras; j++) { ucc_rank_t e = ucc_kn_extra(&p, j); CHK(shm_prog_recv(t, e, dst + off[e], cnt[e], mt, AGV_ID(0, e))); }
        CHK(shm_prog_wait(t));
    }
    for (uint64_t dist = 1; dist < p.n_full; dist *= p.radix, round++) {
        unsigned np = ucc_kn_round_peers(&p, dist, peers);
        ucc_rank_t mine = (ucc_rank_t)((r / dist) * dist);
        for (unsigned i = 0; i < np; i++) {
            ucc_rank_t peer = peers[i], theirs = (ucc_rank_t)((peer / dist) * dist);
            for (ucc_rank_t q = 0; q < dist; q++) {
                /* every base rank of a set stands for itself and for the extras it proxies (rank + j * n_full) */
                for (ucc_rank_t b = mine + q; b < N; b += p.n_full) CHK(shm_prog_send(t, peer, dst + off[b], cnt[b], mt, AGV_ID(round, b)));
                for (ucc_rank_t b = theirs + q; b < N; b += p.n_full) CHK(shm_prog_recv(t, peer, dst + off[b], cnt[b], mt, AGV_ID(round, b)));
            }
        }
        CHK(shm_prog_wait(t));
    }
    if (p.type == UCC_KN_NODE_PROXY) {
        for (unsigned j = 0; j < p.n_extras; j++) {
            ucc_rank_t e = ucc_kn_extra(&p, j);
            for (ucc_rank_t b = 0; b < N; b++) if (b != e) CHK(shm_prog_send(t, e, dst + off[b], cnt[b], mt, AGV_ID(23, b)));
        }
        CHK(shm_prog_wait(t));
    }
#undef AGV_ID
err:
    return st;
}
ucc_status_t ucc_tl_shm_allgather_knomial(ucc_tl_shm_task_t *t) { return allgather_knomial_common(t, 0); }
ucc_status_t ucc_tl_shm_allgatherv_knomial(ucc_tl_shm_task_t *t) { return allgather_knomial_common(t, 1); }

/* ================================================================== */
/* alltoall(v)                                                         */
/* ================================================================== */
static ucc_status_t a2a_common(ucc_tl_shm_task_t *t, int is_v)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank; int inplace = UCC_IS_INPLACE(*a);
    unsigned nposts = is_v ? SHM_CTX(t->team)->cfg.alltoallv_pairwise_num_posts : SHM_CTX(t->team)->cfg.alltoall_pairwise_num_posts;
    size_t sdt, ddt; ucc_memory_type_t smt, dmt; char *src, *dst; ucc_status_t st = UCC_OK;
    if (is_v) { sdt = ucc_dt_size(a->src.info_v.datatype); ddt = ucc_dt_size(a->dst.info_v.datatype); smt = a->src.info_v.mem_type; dmt = a->dst.info_v.mem_type; src = (char *)a->src.info_v.buffer; dst = (char *)a->dst.info_v.buffer; }
    else { sdt = ucc_dt_size(a->src.info.datatype); ddt = ucc_dt_size(a->dst.info.datatype); smt = a->src.info.mem_type; dmt = a->dst.info.mem_type; src = (char *)a->src.info.buffer; dst = (char *)a->dst.info.buffer; }
    if (inplace) { /* stage the whole send side: every block is overwritten by an incoming one */
        size_t tot; void *tmp;
        if (is_v) { tot = 0; for (ucc_rank_t i = 0; i < N; i++) { size_t e = (ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, i) + ucc_coll_args_get_count(a, a->dst.info_v.counts, i)) * ddt; if (e > tot) tot = e; } }
        else tot = a->dst.info.count * ddt;
        CHK(shm_task_scratch(t, tot, dmt, &tmp)); CHK(shm_prog_copy(t, tmp, dst, tot, dmt, dmt));
        src = (char *)tmp; sdt = ddt; smt = dmt;
    }
#define SCNT(_i) (is_v ? ucc_coll_args_get_count(a, inplace ? a->dst.info_v.counts : a->src.info_v.counts, _i) * sdt : (inplace ? a->dst.info.count : a->src.info.count) / N * sdt)
#define SOFF(_i) (is_v ? ucc_coll_args_get_displacement(a, inplace ? a->dst.info_v.displacements : a->src.info_v.displacements, _i) * sdt : (size_t)(_i) * ((inplace ? a->dst.info.count : a->src.info.count) / N * sdt))
#define DCNT(_i) (is_v ? ucc_coll_args_get_count(a, a->dst.info_v.counts, _i) * ddt : a->dst.info.count / N * ddt)
#define DOFF(_i) (is_v ? ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, _i) * ddt : (size_t)(_i) * (a->dst.info.count / N * ddt))
    CHK(shm_prog_copy(t, dst + DOFF(r), src + SOFF(r), ucc_min(SCNT(r), DCNT(r)), dmt, smt));
    for (ucc_rank_t s = 1; s < N; s++) {
        ucc_rank_t to = (r + s) % N, from = (r + N - s) % N;
        CHK(shm_prog_recv(t, from, dst + DOFF(from), DCNT(from), dmt, 1)); CHK(shm_prog_send(t, to, src + SOFF(to), SCNT(to), smt, 1));
        if (nposts && nposts != UCC_UUNITS_AUTO && s % nposts == 0) CHK(shm_prog_wait(t));
    }
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_alltoall_pairwise(ucc_tl_shm_task_t *t) { return a2a_common(t, 0); }
ucc_status_t ucc_tl_shm_alltoallv_pairwise(ucc_tl_shm_task_t *t) { return a2a_common(t, 1); }

/* one-sided alltoall(v) (role of reference tl/ucp alltoall_onesided.c / alltoallv_onesided.c, which put into registered
 * memory): every rank tells each peer where that peer's block starts, then READS its blocks straight out of the peers'
 * source buffers (pointer inside one process, process_vm_readv across processes) and finally confirms to each peer that
 * its buffer is no longer needed.  One copy per block, no intermediate cells; host memory only. */
static ucc_status_t a2a_onesided(ucc_tl_shm_task_t *t, int is_v)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank; const int inplace = 0;
    size_t sdt, ddt; ucc_memory_type_t smt, dmt; char *src, *dst; ucc_status_t st = UCC_OK; void *sv; uint64_t *out, *in; char *tok;
    if (UCC_IS_INPLACE(*a)) return UCC_ERR_NOT_SUPPORTED;
    if (is_v) { sdt = ucc_dt_size(a->src.info_v.datatype); ddt = ucc_dt_size(a->dst.info_v.datatype); smt = a->src.info_v.mem_type; dmt = a->dst.info_v.mem_type; src = (char *)a->src.info_v.buffer; dst = (char *)a->dst.info_v.buffer; }
    else { sdt = ucc_dt_size(a->src.info.datatype); ddt = ucc_dt_size(a->dst.info.datatype); smt = a->src.info.mem_type; dmt = a->dst.info.mem_type; src = (char *)a->src.info.buffer; dst = (char *)a->dst.info.buffer; }
    if (smt != UCC_MEMORY_TYPE_HOST || dmt != UCC_MEMORY_TYPE_HOST) return UCC_ERR_NOT_SUPPORTED;
    for (ucc_rank_t p = 0; p < N; p++) if (p != r && !ucc_tl_shm_can_get(t->team, ucc_ep_map_eval(t->vmap, p))) return UCC_ERR_NOT_SUPPORTED;
    CHK(shm_task_scratch(t, (size_t)N * 18, UCC_MEMORY_TYPE_HOST, &sv));
    out = (uint64_t *)sv; in = out + N; tok = (char *)(in + N);
    /* registered destinations (ucc_mem_map handles of every member, UCC_COLL_ARGS_FLAG_DST_MEMH_GLOBAL): PUT variant, the shape of
     * the reference's alltoall_onesided.c - every rank writes its blocks straight into the members' destination buffers, whose
     * addresses follow from the handles (same offset inside the registered segment on every member), then tells each peer "my
     * block is in".  One message round instead of two.  One-sided contract: a destination may be written as soon as ANY member
     * has entered the collective - the application orders that (e.g. with a barrier), as with the reference. */
    if (!is_v && (a->mask & UCC_COLL_ARGS_FIELD_MEM_MAP_DST_MEMH) && (a->flags & UCC_COLL_ARGS_FLAG_DST_MEMH_GLOBAL) && a->dst_memh.global_memh &&
        N == UCC_TL_TEAM_SIZE(t->team)) {
        const ucc_base_context_t *bctx = t->team->super.super.context;
        ucc_mem_map_mem_h *gl = a->dst_memh.global_memh;
        shm_memh_t *mine = (shm_memh_t *)ucc_mem_map_tl_handle(gl[r], bctx);
        size_t blk = DCNT(0), off = 0; int ok = mine && mine->magic == UCC_TL_SHM_MEMH_MAGIC && (uint64_t)(uintptr_t)dst >= mine->addr && (uint64_t)(uintptr_t)dst + blk * N <= mine->addr + mine->len;
        if (ok) off = (size_t)((uint64_t)(uintptr_t)dst - mine->addr);
        for (ucc_rank_t p = 0; p < N && ok; p++) {
            shm_memh_t *h = (shm_memh_t *)ucc_mem_map_tl_handle(gl[p], bctx);
            if (!h || h->magic != UCC_TL_SHM_MEMH_MAGIC || (p != r && !h->reachable) || off + blk * N > h->len || (p != r && !ucc_tl_shm_can_put(t->team, ucc_ep_map_eval(t->vmap, p)))) ok = 0;
        }
        if (ok) {
            tl_debug(UCC_TL_TEAM_LIB(t->team), "alltoall onesided: put into the members' registered destinations (offset %zu in the segment)", off);
            for (ucc_rank_t s = 1; s < N; s++) {
                ucc_rank_t p = (r + s) % N;
                shm_memh_t *h = (shm_memh_t *)ucc_mem_map_tl_handle(gl[p], bctx);
                CHK(shm_prog_put(t, p, src + SOFF(p), h->addr + off + (size_t)r * blk, blk, smt));
            }
            CHK(shm_prog_copy(t, dst + DOFF(r), src + SOFF(r), blk, dmt, smt));
            for (ucc_rank_t p = 0; p < N; p++) {
                if (p == r) continue;
                tok[p] = 1;
                CHK(shm_prog_send(t, p, &tok[p], 1, UCC_MEMORY_TYPE_HOST, 2)); CHK(shm_prog_recv(t, p, &tok[N + p], 1, UCC_MEMORY_TYPE_HOST, 2));
            }
            CHK(shm_prog_wait(t));
            return UCC_OK;
        }
    }
    for (ucc_rank_t p = 0; p < N; p++) {
        if (p == r) continue;
        out[p] = (uint64_t)(uintptr_t)(src + SOFF(p)); tok[p] = 1;
        CHK(shm_prog_send(t, p, &out[p], 8, UCC_MEMORY_TYPE_HOST, 1)); CHK(shm_prog_recv(t, p, &in[p], 8, UCC_MEMORY_TYPE_HOST, 1));
    }
    CHK(shm_prog_copy(t, dst + DOFF(r), src + SOFF(r), ucc_min(SCNT(r), DCNT(r)), dmt, smt));
    CHK(shm_prog_wait(t));
    for (ucc_rank_t s = 1; s < N; s++) { ucc_rank_t p = (r + s) % N; CHK(shm_prog_get(t, p, dst + DOFF(p), &in[p], DCNT(p), dmt)); }
    for (ucc_rank_t p = 0; p < N; p++) { /* "I have read your buffer" in both directions */
        if (p == r) continue;
        CHK(shm_prog_send(t, p, &tok[p], 1, UCC_MEMORY_TYPE_HOST, 2)); CHK(shm_prog_recv(t, p, &tok[N + p], 1, UCC_MEMORY_TYPE_HOST, 2));
    }
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_alltoall_onesided(ucc_tl_shm_task_t *t) { return a2a_onesided(t, 0); }
ucc_status_t ucc_tl_shm_alltoallv_onesided(ucc_tl_shm_task_t *t) { return a2a_onesided(t, 1); }
/* Bruck alltoall: log2(N) rounds, each moving the blocks whose index has bit k set (latency optimal for small blocks) */
#define UCC_TL_SHM_MAX_BRUCK 1024
/* the log2(N) exchange rounds of Bruck's alltoall on N slots of `blk` bytes: on entry w[i] = data for rank (r + i) % N, on exit
 * w[i] = data from rank (r - i + N) % N (index math: coll_patterns/bruck_alltoall.h) */
static ucc_status_t bruck_rounds(ucc_tl_shm_task_t *t, char *w, char *pk, size_t blk, ucc_memory_type_t mt, unsigned step)
{
    ucc_rank_t N = t->vsize, r = t->vrank; ucc_status_t st = UCC_OK;
    for (unsigned k = 0; k < ucc_bruck_n_steps(N); k++, step++) {
        ucc_rank_t idx[UCC_TL_SHM_MAX_BRUCK], n = ucc_bruck_step_blocks(N, k, idx);
        for (ucc_rank_t j = 0; j < n; j++) CHK(shm_prog_copy(t, pk + (size_t)j * blk, w + (size_t)idx[j] * blk, blk, mt, mt));
        CHK(shm_prog_send(t, ucc_bruck_send_peer(r, N, k), pk, (size_t)n * blk, mt, step)); CHK(shm_prog_recv(t, ucc_bruck_recv_peer(r, N, k), pk + (size_t)n * blk, (size_t)n * blk, mt, step)); CHK(shm_prog_wait(t));
        for (ucc_rank_t j = 0; j < n; j++) CHK(shm_prog_copy(t, w + (size_t)idx[j] * blk, pk + (size_t)(n + j) * blk, blk, mt, mt));
    }
err:
    return st;
}
/* Bruck alltoall: log2(N) rounds, each moving the blocks whose index has bit k set (latency optimal for small blocks) */
ucc_status_t ucc_tl_shm_alltoall_bruck(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank; int inplace = UCC_IS_INPLACE(*a);
    size_t blk = a->dst.info.count / N * ucc_dt_size(a->dst.info.datatype);
    ucc_memory_type_t mt = a->dst.info.mem_type; char *dst = (char *)a->dst.info.buffer, *src = inplace ? dst : (char *)a->src.info.buffer;
    void *wv, *pv; char *w, *pk; ucc_status_t st;
    if ((!inplace && a->src.info.mem_type != mt) || N > UCC_TL_SHM_MAX_BRUCK) return UCC_ERR_NOT_SUPPORTED;
    CHK(shm_task_scratch(t, blk * N, mt, &wv)); CHK(shm_task_scratch(t, blk * N, mt, &pv)); w = (char *)wv; pk = (char *)pv;
    /* phase 1: local rotation, w[i] = src[(r+i)%N] */
    CHK(shm_prog_copy(t, w, src + (size_t)r * blk, (N - r) * blk, mt, mt)); if (r) CHK(shm_prog_copy(t, w + (size_t)(N - r) * blk, src, (size_t)r * blk, mt, mt));
    CHK(bruck_rounds(t, w, pk, blk, mt, 1));
    /* phase 3: inverse rotation, dst[(r - i + N) % N] = w[i] */
    for (ucc_rank_t i = 0; i < N; i++) CHK(shm_prog_copy(t, dst + (size_t)ucc_bruck_final_src(r, N, i) * blk, w + (size_t)i * blk, blk, mt, mt));
err:
    return st;
}
/* hybrid alltoallv (role of reference tl/ucp alltoallv_hybrid.c): a message of at most ALLTOALLV_HYBRID_THRESH bytes travels in a
 * fixed slot through Bruck's log2(N) rounds together with all the other small ones (N-1 latencies become log2 N), everything
 * bigger goes pairwise, straight from the send to the receive buffer.  Sender and receiver classify a message by its size, which
 * both know (what s sends to d is what d receives from s), so the two paths never disagree. */
ucc_status_t ucc_tl_shm_alltoallv_hybrid(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank;
    size_t T = SHM_CTX(t->team)->cfg.alltoallv_hybrid_thresh, sdt = ucc_dt_size(a->src.info_v.datatype), ddt = ucc_dt_size(a->dst.info_v.datatype);
    ucc_memory_type_t smt = a->src.info_v.mem_type, dmt = a->dst.info_v.mem_type;
    char *src = (char *)a->src.info_v.buffer, *dst = (char *)a->dst.info_v.buffer, *w, *pk; void *wv, *pv; ucc_status_t st; int any_small = 0;
    if (UCC_IS_INPLACE(*a) || smt != dmt || N > UCC_TL_SHM_MAX_BRUCK || N < 3 || !T) return UCC_ERR_NOT_SUPPORTED;
#define HS(_i) (ucc_coll_args_get_count(a, a->src.info_v.counts, _i) * sdt)
#define HSO(_i) (ucc_coll_args_get_displacement(a, a->src.info_v.displacements, _i) * sdt)
#define HD(_i) (ucc_coll_args_get_count(a, a->dst.info_v.counts, _i) * ddt)
#define HDO(_i) (ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, _i) * ddt)
    T = ucc_align_up(T, 8);
    CHK(shm_task_scratch(t, T * N, dmt, &wv)); CHK(shm_task_scratch(t, T * N, dmt, &pv)); w = (char *)wv; pk = (char *)pv;
    /* big messages first: they are in flight while the small ones hop */
    for (ucc_rank_t s = 1; s < N; s++) {
        ucc_rank_t to = (r + s) % N, from = (r + N - s) % N;
        if (HD(from) > T) CHK(shm_prog_recv(t, from, dst + HDO(from), HD(from), dmt, 100));
        if (HS(to) > T) CHK(shm_prog_send(t, to, src + HSO(to), HS(to), smt, 100));
    }
    /* my own block: small ones ride in slot 0 of the Bruck rounds (which never leaves this rank), a big one is a plain copy
     * (found by the random-program test: it was simply missing) */
    if (HS(r) > T) CHK(shm_prog_copy(t, dst + HDO(r), src + HSO(r), ucc_min(HS(r), HD(r)), dmt, smt));
    for (ucc_rank_t i = 0; i < N; i++) { ucc_rank_t d = (r + i) % N; if (HS(d) <= T) { any_small = 1; if (HS(d) != 0) CHK(shm_prog_copy(t, w + (size_t)i * T, src + HSO(d), HS(d), dmt, smt)); } }
    (void)any_small; /* the rounds run even if I have nothing small to send: other ranks' small blocks are routed through me */
    CHK(bruck_rounds(t, w, pk, T, dmt, 1));
    for (ucc_rank_t i = 0; i < N; i++) { ucc_rank_t s_ = ucc_bruck_final_src(r, N, i); if (HD(s_) <= T && HD(s_) != 0) CHK(shm_prog_copy(t, dst + HDO(s_), w + (size_t)i * T, HD(s_), dmt, dmt)); }
    CHK(shm_prog_wait(t));
#undef HS
#undef HSO
#undef HD
#undef HDO
err:
    return st;
}

/* ================================================================== */
/* gather(v) / scatter(v)                                              */
/* ================================================================== */
ucc_status_t ucc_tl_shm_gather_linear(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank, root = (ucc_rank_t)a->root; ucc_status_t st = UCC_OK;
    if (r == root) {
        size_t blk = a->dst.info.count / N * ucc_dt_size(a->dst.info.datatype); char *dst = (char *)a->dst.info.buffer;
        if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, dst + (size_t)r * blk, a->src.info.buffer, blk, a->dst.info.mem_type, a->src.info.mem_type));
        for (ucc_rank_t i = 0; i < N; i++) if (i != root) CHK(shm_prog_recv(t, i, dst + (size_t)i * blk, blk, a->dst.info.mem_type, 1));
    } else CHK(shm_prog_send(t, root, a->src.info.buffer, a->src.info.count * ucc_dt_size(a->src.info.datatype), a->src.info.mem_type, 1));
    CHK(shm_prog_wait(t));
err:
    return st;
}
/* knomial gather: subtrees are forwarded as contiguous ranges of virtual ranks through a scratch */
ucc_status_t ucc_tl_shm_gather_knomial(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank, root = (ucc_rank_t)a->root, vr = (r + N - root) % N;
    int is_root = r == root; ucc_kn_tree_t tr; ucc_status_t st = UCC_OK;
    size_t blk = is_root ? a->dst.info.count / N * ucc_dt_size(a->dst.info.datatype) : a->src.info.count * ucc_dt_size(a->src.info.datatype);
    ucc_memory_type_t mt = is_root ? a->dst.info.mem_type : a->src.info.mem_type;
    unsigned radix = cfg_radix_u(SHM_CTX(t->team)->cfg.scatter_kn_radix, cfg_radix(&SHM_CTX(t->team)->cfg.gather_kn_radix, blk * N, mt, KN_DFLT(t), N), N);
    ucc_rank_t span = ucc_kn_subtree_size(vr, N, radix);
    void *sv; char *s;
    ucc_kn_tree_init(&tr, r, N, root, radix);
    /* scratch holds virtual ranks [vr, vr+span) */
    CHK(shm_task_scratch(t, (size_t)span * blk, mt, &sv)); s = (char *)sv;
    CHK(shm_prog_copy(t, s, (is_root && UCC_IS_INPLACE(*a)) ? (void *)((char *)a->dst.info.buffer + (size_t)r * blk) : a->src.info.buffer, blk, mt,
                      (is_root && UCC_IS_INPLACE(*a)) ? mt : a->src.info.mem_type));
    for (unsigned i = 0; i < tr.n_children; i++) {
        ucc_rank_t cvr = (tr.children[i] + N - root) % N, cspan = ucc_kn_subtree_size(cvr, N, radix);
        CHK(shm_prog_recv(t, tr.children[i], s + (size_t)(cvr - vr) * blk, (size_t)cspan * blk, mt, 1));
    }
    CHK(shm_prog_wait(t));
    if (!is_root) { CHK(shm_prog_send(t, tr.parent, s, (size_t)span * blk, mt, 1)); CHK(shm_prog_wait(t)); }
    else { /* un-rotate virtual ranks into team ranks */
        char *dst = (char *)a->dst.info.buffer;
        CHK(shm_prog_copy(t, dst + (size_t)root * blk, s, (size_t)(N - root) * blk, mt, mt));
        if (root) CHK(shm_prog_copy(t, dst, s + (size_t)(N - root) * blk, (size_t)root * blk, mt, mt));
    }
err:
    return st;
}
ucc_status_t ucc_tl_shm_gatherv_linear(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank, root = (ucc_rank_t)a->root; ucc_status_t st = UCC_OK;
    unsigned np = NPOSTS(SHM_CTX(t->team)->cfg.gatherv_linear_num_posts), posted = 0;
    if (r == root) {
        size_t dts = ucc_dt_size(a->dst.info_v.datatype); char *dst = (char *)a->dst.info_v.buffer;
        for (ucc_rank_t i = 0; i < N; i++) {
            size_t c = ucc_coll_args_get_count(a, a->dst.info_v.counts, i) * dts, o = ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, i) * dts;
            if (i == root) { if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, dst + o, a->src.info.buffer, c, a->dst.info_v.mem_type, a->src.info.mem_type)); }
            else { CHK(shm_prog_recv(t, i, dst + o, c, a->dst.info_v.mem_type, 1)); if (np && ++posted % np == 0) CHK(shm_prog_wait(t)); }
        }
    } else CHK(shm_prog_send(t, root, a->src.info.buffer, a->src.info.count * ucc_dt_size(a->src.info.datatype), a->src.info.mem_type, 1));
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_scatter_linear(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank, root = (ucc_rank_t)a->root; ucc_status_t st = UCC_OK;
    if (r == root) {
        size_t blk = a->src.info.count / N * ucc_dt_size(a->src.info.datatype); char *src = (char *)a->src.info.buffer;
        for (ucc_rank_t i = 0; i < N; i++) if (i != root) CHK(shm_prog_send(t, i, src + (size_t)i * blk, blk, a->src.info.mem_type, 1));
        if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, a->dst.info.buffer, src + (size_t)r * blk, blk, a->dst.info.mem_type, a->src.info.mem_type));
    } else CHK(shm_prog_recv(t, root, a->dst.info.buffer, a->dst.info.count * ucc_dt_size(a->dst.info.datatype), a->dst.info.mem_type, 1));
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_scatter_knomial(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank, root = (ucc_rank_t)a->root, vr = (r + N - root) % N;
    int is_root = r == root; ucc_kn_tree_t tr; ucc_status_t st = UCC_OK;
    size_t blk = is_root ? a->src.info.count / N * ucc_dt_size(a->src.info.datatype) : a->dst.info.count * ucc_dt_size(a->dst.info.datatype);
    ucc_memory_type_t mt = is_root ? a->src.info.mem_type : a->dst.info.mem_type;
    unsigned radix = cfg_radix(&SHM_CTX(t->team)->cfg.gather_kn_radix, blk * N, mt, KN_DFLT(t), N);
    ucc_rank_t span = ucc_kn_subtree_size(vr, N, radix);
    void *sv; char *s;
    ucc_kn_tree_init(&tr, r, N, root, radix);
    CHK(shm_task_scratch(t, (size_t)span * blk, mt, &sv)); s = (char *)sv;
    if (is_root) { /* rotate into virtual rank order */
        char *src = (char *)a->src.info.buffer;
        CHK(shm_prog_copy(t, s, src + (size_t)root * blk, (size_t)(N - root) * blk, mt, mt));
        if (root) CHK(shm_prog_copy(t, s + (size_t)(N - root) * blk, src, (size_t)root * blk, mt, mt));
    } else { CHK(shm_prog_recv(t, tr.parent, s, (size_t)span * blk, mt, 1)); CHK(shm_prog_wait(t)); }
    for (unsigned i = 0; i < tr.n_children; i++) {
        ucc_rank_t cvr = (tr.children[i] + N - root) % N, cspan = ucc_kn_subtree_size(cvr, N, radix);
        CHK(shm_prog_send(t, tr.children[i], s + (size_t)(cvr - vr) * blk, (size_t)cspan * blk, mt, 1));
    }
    if (!(is_root && UCC_IS_INPLACE(*a))) CHK(shm_prog_copy(t, a->dst.info.buffer, s, blk, is_root ? a->dst.info.mem_type : mt, mt));
    CHK(shm_prog_wait(t));
err:
    return st;
}
ucc_status_t ucc_tl_shm_scatterv_linear(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = ARGS(t); ucc_rank_t N = t->vsize, r = t->vrank, root = (ucc_rank_t)a->root; ucc_status_t st = UCC_OK;
    unsigned np = NPOSTS(SHM_CTX(t->team)->cfg.scatterv_linear_num_posts), posted = 0;
    if (r == root) {
        size_t dts = ucc_dt_size(a->src.info_v.datatype); char *src = (char *)a->src.info_v.buffer;
        for (ucc_rank_t i = 0; i < N; i++) {
            size_t c = ucc_coll_args_get_count(a, a->src.info_v.counts, i) * dts, o = ucc_coll_args_get_displacement(a, a->src.info_v.displacements, i) * dts;
            if (i == root) { if (!UCC_IS_INPLACE(*a)) CHK(shm_prog_copy(t, a->dst.info.buffer, src + o, c, a->dst.info.mem_type, a->src.info_v.mem_type)); }
            else { CHK(shm_prog_send(t, i, src + o, c, a->src.info_v.mem_type, 1)); if (np && ++posted % np == 0) CHK(shm_prog_wait(t)); }
        }
    } else CHK(shm_prog_recv(t, root, a->dst.info.buffer, a->dst.info.count * ucc_dt_size(a->dst.info.datatype), a->dst.info.mem_type, 1));
    CHK(shm_prog_wait(t));
err:
    return st;
}

/* ================================================================== */
/* service collectives (any subset of the team)                        */
/* ================================================================== */
ucc_status_t ucc_tl_shm_service_allreduce_prog(ucc_tl_shm_task_t *t, void *sbuf, void *rbuf, size_t count)
{
    size_t len = count * ucc_dt_size(t->dt); void *scratch; ucc_status_t st;
    unsigned radix = t->vsize < 4 ? 2 : 4;
    CHK(shm_task_scratch(t, (size_t)(radix - 1) * len, UCC_MEMORY_TYPE_HOST, &scratch));
    if (sbuf != rbuf) CHK(shm_prog_copy(t, rbuf, sbuf, len, UCC_MEMORY_TYPE_HOST, UCC_MEMORY_TYPE_HOST));
    CHK(prog_allreduce_knomial(t, rbuf, scratch, count, UCC_MEMORY_TYPE_HOST, radix, 1));
err:
    return st;
}
ucc_status_t ucc_tl_shm_service_allgather_prog(ucc_tl_shm_task_t *t, void *sbuf, void *rbuf, size_t msgsize)
{
    ucc_rank_t N = t->vsize, r = t->vrank; char *dst = (char *)rbuf; ucc_status_t st;
    CHK(shm_prog_copy(t, dst + (size_t)r * msgsize, sbuf, msgsize, UCC_MEMORY_TYPE_HOST, UCC_MEMORY_TYPE_HOST));
    CHK(shm_prog_wait(t));
    if (N <= 8 || msgsize <= 4096) { /* linear: one round */
        for (ucc_rank_t s = 1; s < N; s++) {
            ucc_rank_t to = (r + s) % N, from = (r + N - s) % N;
            CHK(shm_prog_recv(t, from, dst + (size_t)from * msgsize, msgsize, UCC_MEMORY_TYPE_HOST, 1));
            CHK(shm_prog_send(t, to, sbuf, msgsize, UCC_MEMORY_TYPE_HOST, 1));
        }
        CHK(shm_prog_wait(t));
    } else {
        ucc_rank_t next = (r + 1) % N, prev = (r + N - 1) % N;
        for (ucc_rank_t s = 0; s + 1 < N; s++) {
            ucc_rank_t sb = (r + N - s) % N, rb = (r + N - s - 1) % N;
            CHK(shm_prog_send(t, next, dst + (size_t)sb * msgsize, msgsize, UCC_MEMORY_TYPE_HOST, 1 + s));
            CHK(shm_prog_recv(t, prev, dst + (size_t)rb * msgsize, msgsize, UCC_MEMORY_TYPE_HOST, 1 + s));
            CHK(shm_prog_wait(t));
        }
    }
err:
    return st;
}
