/* tl/nvl zero-copy support: per-collective exchange of the members' USER buffers.
 *
 * The staged kernels copy src into the symmetric heap and the result out of it: two extra passes over HBM for
 * every byte.  For large messages each rank instead publishes (CUDA IPC handle of the allocation, offset) of its
 * src / dst in a small POSIX shm board it owns; every peer maps the allocation once (cached by allocation base,
 * validated by handle bytes) and the kernel then reads / writes the user buffers in place over NVLink.
 * Inside one process (the gtest-style harness) the board simply carries raw pointers.
 *
 * Compared with the reference (tl/cuda alltoallv_ce.c:70-126, tl_cuda_cache.c:62-353): same idea (IPC handle
 * exchange through a shared segment + an open-handle cache) but the segment is single-writer per rank, lock free
 * (sequence numbers), there is no shm barrier, and a rank whose buffer cannot be exported simply says so, which
 * makes all ranks fall back to the staged kernel for that collective - the decision is symmetric by construction. */
#include "tl_nvl.h"
#include "utils/ucc_sys.h"
#include "utils/ucc_atomic.h"
#include <stdio.h>
#include <unistd.h>

#define NVL_LIB(_t) ((_t)->super.super.context->lib)

ucc_status_t ucc_tl_nvl_xb_create(ucc_tl_nvl_team_t *team)
{
    static uint32_t counter = 0;
    void *addr = NULL;
    ucc_status_t st;
    uint32_t my = ucc_atomic_fadd32(&counter, 1);
    if (my == 0) ucc_shm_reap_stale("ucc_b200_nvlxb.");
    snprintf(team->xb_name, sizeof(team->xb_name), "/ucc_b200_nvlxb.%d.%u", (int)getpid(), my);
    st = ucc_shm_create(team->xb_name, sizeof(nvl_xb_seg_t), &addr);
    if (st != UCC_OK) { team->xb_name[0] = 0; return st; }
    memset(addr, 0, sizeof(nvl_xb_seg_t));
    team->xb_mine = (nvl_xb_seg_t *)addr;
    team->xb_named = 1;
    return UCC_OK;
}

ucc_status_t ucc_tl_nvl_xb_attach(ucc_tl_nvl_team_t *team)
{
    ucc_rank_t N = UCC_TL_TEAM_SIZE(team), me = UCC_TL_TEAM_RANK(team);
    for (ucc_rank_t p = 0; p < N; p++) {
        void *addr = NULL;
        if (p == me) { team->xb[p] = team->xb_mine; continue; }
        if (!team->infos[p].xb_name[0] || ucc_shm_attach(team->infos[p].xb_name, sizeof(nvl_xb_seg_t), &addr) != UCC_OK) return UCC_ERR_NO_RESOURCE;
        team->xb[p] = (nvl_xb_seg_t *)addr;
    }
    return UCC_OK;
}

void ucc_tl_nvl_xb_unlink(ucc_tl_nvl_team_t *team)
{ if (team->xb_named) { ucc_shm_unlink(team->xb_name); team->xb_named = 0; } }

void ucc_tl_nvl_xb_release(ucc_tl_nvl_team_t *team)
{
    ucc_rank_t N = UCC_TL_TEAM_SIZE(team), me = UCC_TL_TEAM_RANK(team);
    for (ucc_rank_t p = 0; p < N && p < NVL_MAX_PEERS; p++) {
        nvl_ipc_cache_t *c = &team->ipc_cache[p];
        for (unsigned i = 0; i < c->n; i++) if (c->e[i].mapped) { if (cudaIpcCloseMemHandle(c->e[i].mapped) != cudaSuccess) (void)cudaGetLastError(); }
        c->n = 0;
        if (p != me && team->xb[p]) ucc_shm_detach(team->xb[p], sizeof(nvl_xb_seg_t));
        team->xb[p] = NULL;
    }
    ucc_tl_nvl_xb_unlink(team);
    if (team->xb_mine) { ucc_shm_detach(team->xb_mine, sizeof(nvl_xb_seg_t)); team->xb_mine = NULL; }
}

static void export_buf(ucc_tl_nvl_team_t *team, const void *ptr, size_t len, nvl_xb_buf_t *b)
{
    memset(b, 0, sizeof(*b));
    if (!ptr || !len) { b->kind = NVL_XB_EMPTY; return; }
    b->len = len;
    if (team->heap_kind == NVL_HEAP_LOCAL) { b->kind = NVL_XB_RAW; b->base = (uint64_t)(uintptr_t)ptr; return; }
    {
        void *base = NULL; size_t alen = 0;
        if (ucc_cuda_get_address_range(ptr, &base, &alen) != UCC_OK) { b->kind = NVL_XB_NONE; return; }
        if (cudaIpcGetMemHandle(&b->handle, base) != cudaSuccess) { (void)cudaGetLastError(); b->kind = NVL_XB_NONE; return; } /* VMM / pool / managed memory */
        b->kind = NVL_XB_IPC; b->base = (uint64_t)(uintptr_t)base; b->alloc_len = alen; b->off = (uint64_t)((const char *)ptr - (const char *)base);
    }
}

void ucc_tl_nvl_xb_export(ucc_tl_nvl_team_t *team, const void *ptr, size_t len, nvl_xb_buf_t *b) { export_buf(team, ptr, len, b); }

/* returns 0 when the slot is still in use by a slow peer (caller retries later) */
int ucc_tl_nvl_xb_publish(ucc_tl_nvl_team_t *team, uint64_t cseq, const void *src, size_t src_len, void *dst, size_t dst_len, int usable, const size_t *aux)
{
    ucc_rank_t N = UCC_TL_TEAM_SIZE(team), me = UCC_TL_TEAM_RANK(team);
    nvl_xb_entry_t *e = &team->xb_mine->e[cseq % NVL_XB_SLOTS];
    if (cseq >= NVL_XB_SLOTS)
        for (ucc_rank_t p = 0; p < N; p++) if (p != me && ucc_load_acquire(&team->xb[p]->consumed) + NVL_XB_SLOTS <= cseq) return 0;
    if (usable) { export_buf(team, src, src_len, &e->src); export_buf(team, dst, dst_len, &e->dst); }
    else { memset(&e->src, 0, sizeof(e->src)); memset(&e->dst, 0, sizeof(e->dst)); e->src.kind = e->dst.kind = NVL_XB_NONE; }
    for (ucc_rank_t p = 0; p < N && p < NVL_MAX_PEERS; p++) e->aux[p] = aux ? (uint64_t)aux[p] : 0;
    ucc_store_release(&e->seq, cseq + 1);
    return 1;
}

int ucc_tl_nvl_xb_ready(ucc_tl_nvl_team_t *team, uint64_t cseq)
{
    ucc_rank_t N = UCC_TL_TEAM_SIZE(team), me = UCC_TL_TEAM_RANK(team);
    for (ucc_rank_t p = 0; p < N; p++) if (p != me && ucc_load_acquire(&team->xb[p]->e[cseq % NVL_XB_SLOTS].seq) != cseq + 1) return 0;
    return 1;
}

static char *import_buf(ucc_tl_nvl_team_t *team, ucc_rank_t p, const nvl_xb_buf_t *b)
{
    nvl_ipc_cache_t *c = &team->ipc_cache[p];
    void *mapped = NULL;
    unsigned i;
    if (b->kind == NVL_XB_RAW) return (char *)(uintptr_t)b->base;
    if (b->kind != NVL_XB_IPC) return NULL;
    for (i = 0; i < c->n; i++) {
        if (c->e[i].base != b->base) continue;
        if (!memcmp(&c->e[i].handle, &b->handle, sizeof(b->handle))) return (char *)c->e[i].mapped + b->off;
        /* same address, different allocation: the peer freed and re-allocated */
        if (cudaIpcCloseMemHandle(c->e[i].mapped) != cudaSuccess) (void)cudaGetLastError();
        c->e[i] = c->e[--c->n];
        break;
    }
    if (c->n == NVL_IPC_CACHE_MAX) { /* rare: drop everything that is certainly idle */
        cudaDeviceSynchronize();
        for (i = 0; i < c->n; i++) if (cudaIpcCloseMemHandle(c->e[i].mapped) != cudaSuccess) (void)cudaGetLastError();
        c->n = 0;
    }
    if (cudaIpcOpenMemHandle(&mapped, b->handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        tl_debug(NVL_LIB(team), "cudaIpcOpenMemHandle of a buffer of rank %u failed: %s", p, cudaGetErrorString(cudaGetLastError()));
        return NULL;
    }
    c->e[c->n].base = b->base; c->e[c->n].handle = b->handle; c->e[c->n].mapped = mapped; c->n++;
    return (char *)mapped + b->off;
}

char *ucc_tl_nvl_xb_import(ucc_tl_nvl_team_t *team, ucc_rank_t p, const nvl_xb_buf_t *b) { return import_buf(team, p, b); }

/* all members published: decide (identically on every rank) whether the buffers can be used in place and map
 * them.  need_src / need_dst say which sides the kernel touches remotely.  Marks the entries consumed.
 * Returns 1 = in place, 0 = every member stages (symmetric decision), -1 = local mapping failure (the collective must fail). */
int ucc_tl_nvl_xb_resolve(ucc_tl_nvl_team_t *team, uint64_t cseq, int need_src, int need_dst, int need_align, const void *my_src, void *my_dst, nvl_direct_t *d, size_t *aux_for_me)
{
    ucc_rank_t N = UCC_TL_TEAM_SIZE(team), me = UCC_TL_TEAM_RANK(team);
    int ok = 1;
    memset(d, 0, sizeof(*d));
    if (aux_for_me) for (ucc_rank_t p = 0; p < N; p++) aux_for_me[p] = (size_t)team->xb[p]->e[cseq % NVL_XB_SLOTS].aux[me];
    for (ucc_rank_t p = 0; p < N && ok; p++) {
        const nvl_xb_entry_t *e = &team->xb[p]->e[cseq % NVL_XB_SLOTS];
        if (need_src && e->src.kind != NVL_XB_EMPTY && e->src.kind != NVL_XB_RAW && e->src.kind != NVL_XB_IPC) ok = 0;
        if (need_dst && e->dst.kind != NVL_XB_EMPTY && e->dst.kind != NVL_XB_RAW && e->dst.kind != NVL_XB_IPC) ok = 0;
        /* 16-byte alignment keeps every transaction a full vector; anything else takes the staged kernel */
        if (need_align && need_src && e->src.kind != NVL_XB_EMPTY && ((e->src.base + e->src.off) & 15)) ok = 0;
        if (need_align && need_dst && e->dst.kind != NVL_XB_EMPTY && ((e->dst.base + e->dst.off) & 15)) ok = 0;
    }
    /* up to here every member took the same decision (it only depends on what all of them published).  Mapping a peer's
     * allocation can still fail on THIS rank only; the peers will launch the in-place kernel, so falling back to the staged
     * one here would mix two protocols and corrupt data silently: report it (-1), the caller fails the collective. */
    for (ucc_rank_t p = 0; p < N && ok > 0; p++) {
        const nvl_xb_entry_t *e = &team->xb[p]->e[cseq % NVL_XB_SLOTS];
        if (p == me) { d->src[p] = (const char *)my_src; d->dst[p] = (char *)my_dst; continue; }
        if (need_src && e->src.kind != NVL_XB_EMPTY) { d->src[p] = import_buf(team, p, &e->src); if (!d->src[p]) ok = -1; }
        if (need_dst && e->dst.kind != NVL_XB_EMPTY) { d->dst[p] = import_buf(team, p, &e->dst); if (!d->dst[p]) ok = -1; }
    }
    if (team->xb_mine->consumed < cseq + 1) ucc_store_release(&team->xb_mine->consumed, cseq + 1);
    return ok;
}
