/* tl/shm point-to-point layer: MPSC cell ring per context, tag matching,
 * unexpected stash, same-process rendezvous. */
#include "tl_shm.h"
#include "utils/ucc_sys.h"
#include <unistd.h>
#include <sys/uio.h>
#include <sys/prctl.h>

#define CTX_LOCK(_c)   do { if ((_c)->tm == UCC_THREAD_MULTIPLE) ucc_recursive_spin_lock(&(_c)->lock); } while (0)
#define CTX_UNLOCK(_c) do { if ((_c)->tm == UCC_THREAD_MULTIPLE) ucc_recursive_spin_unlock(&(_c)->lock); } while (0)

static inline void shm_copy(void *dst, const void *src, size_t len, ucc_memory_type_t dmt, ucc_memory_type_t smt)
{
    if (!len) return;
    if (dmt == UCC_MEMORY_TYPE_HOST && smt == UCC_MEMORY_TYPE_HOST) memcpy(dst, src, len);
    else ucc_mc_memcpy(dst, src, len, dmt, smt);
}

/* cross-memory attach: read `len` bytes at `remote` of process `pid` into local host memory */
static uint64_t cma_probe_word = 0x75636362323030ull; /* the word peers read to verify that CMA works */
static int cma_read(int pid, void *local, const void *remote, size_t len)
{
    size_t done = 0;
    while (done < len) {
        struct iovec l = {(char *)local + done, len - done}, r = {(char *)(uintptr_t)remote + done, len - done};
        ssize_t n = process_vm_readv((pid_t)pid, &l, 1, &r, 1, 0);
        if (n <= 0) return -1;
        done += (size_t)n;
    }
    return 0;
}
static int cma_write(int pid, const void *local, void *remote, size_t len)
{
    size_t done = 0;
    while (done < len) {
        struct iovec l = {(char *)(uintptr_t)local + done, len - done}, r = {(char *)remote + done, len - done};
        ssize_t n = process_vm_writev((pid_t)pid, &l, 1, &r, 1, 0);
        if (n <= 0) return -1;
        done += (size_t)n;
    }
    return 0;
}
/* payload of a rendezvous: the source lives in this process (plain copy) or in a peer process (CMA, host memory only) */
static void rndv_fetch(ucc_tl_shm_context_t *ctx, ucc_tl_shm_ep_t *ep, void *dst, ucc_memory_type_t dmt, const void *src, ucc_memory_type_t smt, size_t len)
{
    if (!len) return;
    if (ep->same_process) { shm_copy(dst, src, len, dmt, smt); return; }
    if (dmt == UCC_MEMORY_TYPE_HOST) { if (cma_read(ep->addr.pid, dst, src, len)) tl_error(ctx->super.super.lib, "process_vm_readv from pid %d failed: %m", ep->addr.pid); return; }
    { void *tmp = malloc(len); /* device destination: bounce through host memory */
      if (!tmp || cma_read(ep->addr.pid, tmp, src, len)) tl_error(ctx->super.super.lib, "process_vm_readv from pid %d failed: %m", ep->addr.pid);
      else shm_copy(dst, tmp, len, dmt, UCC_MEMORY_TYPE_HOST);
      free(tmp); }
}

ucc_status_t ucc_tl_shm_ctx_p2p_init(ucc_tl_shm_context_t *ctx)
{
    static uint32_t seq = 0;
    uint32_t my_seq = ucc_atomic_fadd32(&seq, 1);
    unsigned n_cells = ctx->cfg.n_cells;
    size_t cell_size;
    ucc_status_t st;
    if (!ucc_is_pow2(n_cells)) { unsigned p = 1; while (p < n_cells) p <<= 1; n_cells = p; }
    cell_size = ucc_align_up(sizeof(shm_cell_hdr_t) + ctx->cfg.cell_payload, 64);
    ctx->ring_len = sizeof(shm_ring_hdr_t) + (size_t)n_cells * cell_size;
    memset(&ctx->addr, 0, sizeof(ctx->addr));
    ctx->addr.pid = (int32_t)getpid(); ctx->addr.host_hash = ucc_sys_host_hash(); /* physical host: an injected (synthetic) placement only shapes the topology */
    ctx->addr.n_cells = n_cells; ctx->addr.cell_size = (uint32_t)cell_size;
    ctx->addr.ep_id = ucc_hash_mix(((uint64_t)(uint32_t)ctx->addr.pid << 32 | my_seq) ^ ucc_sys_host_hash());
    ctx->addr.probe_addr = (uint64_t)(uintptr_t)&cma_probe_word; ctx->addr.probe_val = cma_probe_word;
#ifdef PR_SET_PTRACER
    if (ctx->cfg.cma != UCC_NO) prctl(PR_SET_PTRACER, PR_SET_PTRACER_ANY, 0, 0, 0); /* yama ptrace_scope=1: let the peers read my buffers */
#endif
    if (my_seq == 0) ucc_shm_reap_stale("ucc_b200.");   /* first context of the process: collect the rings of processes that died without cleanup */
    snprintf(ctx->addr.name, sizeof(ctx->addr.name), "/ucc_b200.%d.%u.%llx", (int)getpid(), my_seq, (unsigned long long)(ctx->addr.ep_id & 0xffffff));
    st = ucc_shm_create(ctx->addr.name, ctx->ring_len, (void **)&ctx->ring);
    if (st != UCC_OK) { ucc_shm_unlink(ctx->addr.name); st = ucc_shm_create(ctx->addr.name, ctx->ring_len, (void **)&ctx->ring); }
    if (st != UCC_OK) { tl_error(ctx->super.super.lib, "failed to create shm segment %s (%zu bytes)", ctx->addr.name, ctx->ring_len); return st; }
    ctx->ring->n_cells = n_cells; ctx->ring->cell_size = (uint32_t)cell_size; ctx->ring->head = ctx->ring->tail = 0;
    for (unsigned i = 0; i < n_cells; i++) SHM_CELL(ctx->ring, i)->seq = i;
    ucc_memory_cpu_store_fence();
    ctx->ring->magic = SHM_RING_MAGIC;
    ucc_hash_init(&ctx->eps);
    ucc_list_head_init(&ctx->posted_recvs); ucc_list_head_init(&ctx->pending_sends);
    ucc_list_head_init(&ctx->rndv_sends); ucc_list_head_init(&ctx->unexpected); ucc_list_head_init(&ctx->pending_acks);
    ucc_recursive_spinlock_init(&ctx->lock);
    return ucc_mpool_init(&ctx->req_mp, 0, sizeof(shm_req_t), 0, 64, 64, (unsigned)-1, NULL, ctx->tm, "tl_shm_reqs");
}

void ucc_tl_shm_ctx_p2p_cleanup(ucc_tl_shm_context_t *ctx)
{
    uint64_t k; void *v; shm_unexp_t *u, *t;
    ucc_hash_for_each(&ctx->eps, k, v) {
        ucc_tl_shm_ep_t *ep = (ucc_tl_shm_ep_t *)v; (void)k;
        if (ep->ring != ctx->ring) ucc_shm_detach(ep->ring, ep->map_len);
        free(ep);
    }
    ucc_hash_destroy(&ctx->eps);
    ucc_list_for_each_safe(u, t, &ctx->unexpected, list) { ucc_list_del(&u->list); free(u->data); free(u); }
    while (!ucc_list_is_empty(&ctx->pending_acks)) { ucc_list_link_t *l = ctx->pending_acks.next; ucc_list_del(l); free(l); } /* list link is the first member */
    if (ctx->ring) { ucc_shm_detach(ctx->ring, ctx->ring_len); ucc_shm_unlink(ctx->addr.name); ctx->ring = NULL; }
    ucc_mpool_cleanup(&ctx->req_mp, 1);
}

static ucc_status_t ep_from_addr(ucc_tl_shm_context_t *ctx, const ucc_tl_shm_addr_t *addr, ucc_tl_shm_ep_t **ep_p)
{
    ucc_tl_shm_ep_t *ep = (ucc_tl_shm_ep_t *)ucc_hash_get(&ctx->eps, addr->ep_id);
    if (!ep) {
        ep = (ucc_tl_shm_ep_t *)calloc(1, sizeof(*ep));
        if (!ep) return UCC_ERR_NO_MEMORY;
        ep->addr = *addr;
        ep->same_process = (addr->pid == ctx->addr.pid && addr->host_hash == ctx->addr.host_hash);
        ep->map_len = sizeof(shm_ring_hdr_t) + (size_t)addr->n_cells * addr->cell_size;
        if (addr->ep_id == ctx->addr.ep_id) ep->ring = ctx->ring;
        else {
            ucc_status_t st = ucc_shm_attach(addr->name, ep->map_len, (void **)&ep->ring);
            if (st != UCC_OK) { tl_error(ctx->super.super.lib, "failed to attach to peer segment %s", addr->name); free(ep); return UCC_ERR_NO_RESOURCE; }
            if (ep->ring->magic != SHM_RING_MAGIC) { ucc_shm_detach(ep->ring, ep->map_len); free(ep); return UCC_ERR_NO_RESOURCE; }
            if (!ep->same_process && ctx->cfg.cma != UCC_NO && addr->probe_addr) { /* can I read the peer's memory? (assumed symmetric) */
                uint64_t v = 0;
                ep->cma_ok = cma_read(addr->pid, &v, (const void *)(uintptr_t)addr->probe_addr, sizeof(v)) == 0 && v == addr->probe_val;
                if (!ep->cma_ok && ctx->cfg.cma == UCC_YES) tl_warn(ctx->super.super.lib, "cross-memory attach to pid %d is not permitted: large messages use the copy-in/copy-out ring", addr->pid);
            }
        }
        ucc_hash_put(&ctx->eps, addr->ep_id, ep);
    }
    *ep_p = ep;
    return UCC_OK;
}

ucc_status_t ucc_tl_shm_get_ep(ucc_tl_shm_team_t *team, ucc_rank_t rank, ucc_tl_shm_ep_t **ep_p)
{
    ucc_tl_shm_context_t *ctx = SHM_CTX(team);
    ucc_status_t st = UCC_OK;
    if (!team->eps[rank]) {
        ucc_team_t *core_team = UCC_TL_CORE_TEAM(team);
        ucc_rank_t core_rank = ucc_ep_map_eval(UCC_TL_TEAM_MAP(team), rank);
        uint32_t len = 0;
        ucc_tl_shm_addr_t *addr = (ucc_tl_shm_addr_t *)ucc_get_team_ep_addr(UCC_TL_CORE_CTX(team), core_team, core_rank, ucc_tl_shm.super.id, &len);
        if (!addr || len < sizeof(*addr)) { tl_debug(UCC_TL_TEAM_LIB(team), "rank %u does not run tl/shm", rank); return UCC_ERR_NOT_FOUND; }
        CTX_LOCK(ctx);
        st = ep_from_addr(ctx, addr, &team->eps[rank]);
        CTX_UNLOCK(ctx);
    }
    *ep_p = team->eps[rank];
    return st;
}

/* one-sided access (the "onesided" algorithms): can this context read the memory of team rank `rank` directly? */
int ucc_tl_shm_can_get(ucc_tl_shm_team_t *team, ucc_rank_t rank)
{
    ucc_tl_shm_ep_t *ep;
    return ucc_tl_shm_get_ep(team, rank, &ep) == UCC_OK && (ep->same_process || ep->cma_ok);
}
ucc_status_t ucc_tl_shm_get(ucc_tl_shm_team_t *team, ucc_rank_t rank, void *dst, ucc_memory_type_t dmt, uint64_t remote_addr, size_t len)
{
    ucc_tl_shm_ep_t *ep;
    UCC_CHECK_RET(ucc_tl_shm_get_ep(team, rank, &ep));
    if (!ep->same_process && !ep->cma_ok) return UCC_ERR_NOT_SUPPORTED;
    rndv_fetch(SHM_CTX(team), ep, dst, dmt, (const void *)(uintptr_t)remote_addr, UCC_MEMORY_TYPE_HOST, len);
    return UCC_OK;
}

/* one-sided write into a peer's HOST memory (registered destination of a put-based collective): pointer inside one process,
 * process_vm_writev otherwise; the permission is the one probed for reads (ptrace access is symmetric in read / write) */
int ucc_tl_shm_can_put(ucc_tl_shm_team_t *team, ucc_rank_t rank) { return ucc_tl_shm_can_get(team, rank); }
ucc_status_t ucc_tl_shm_put(ucc_tl_shm_team_t *team, ucc_rank_t rank, const void *src, ucc_memory_type_t smt, uint64_t remote_addr, size_t len)
{
    ucc_tl_shm_ep_t *ep;
    UCC_CHECK_RET(ucc_tl_shm_get_ep(team, rank, &ep));
    if (!len) return UCC_OK;
    if (ep->same_process) { shm_copy((void *)(uintptr_t)remote_addr, src, len, UCC_MEMORY_TYPE_HOST, smt); return UCC_OK; }
    if (!ep->cma_ok || smt != UCC_MEMORY_TYPE_HOST) return UCC_ERR_NOT_SUPPORTED;
    if (cma_write(ep->addr.pid, src, (void *)(uintptr_t)remote_addr, len)) { tl_error(SHM_CTX(team)->super.super.lib, "process_vm_writev to pid %d failed: %m", ep->addr.pid); return UCC_ERR_NO_MESSAGE; }
    return UCC_OK;
}

/* reserve one cell in `ring`; NULL when full */
static inline shm_cell_hdr_t *ring_reserve(shm_ring_hdr_t *ring, uint64_t *pos_p)
{
    uint64_t pos = ucc_load_acquire(&ring->tail);
    for (;;) {
        shm_cell_hdr_t *c = SHM_CELL(ring, pos);
        int64_t dif = (int64_t)ucc_load_acquire(&c->seq) - (int64_t)pos;
        if (dif == 0) { if (ucc_atomic_bool_cswap64(&ring->tail, pos, pos + 1)) { *pos_p = pos; return c; } pos = ucc_load_acquire(&ring->tail); }
        else if (dif < 0) return NULL;
        else pos = ucc_load_acquire(&ring->tail);
    }
}
static inline void ring_publish(shm_cell_hdr_t *c, uint64_t pos) { ucc_store_release(&c->seq, pos + 1); }

/* push as much of a send request as the destination ring accepts */
static void send_push(ucc_tl_shm_context_t *ctx, shm_req_t *r)
{
    shm_ring_hdr_t *ring = r->ep->ring;
    size_t payload = ring->cell_size - sizeof(shm_cell_hdr_t);
    uint64_t pos;
    if (r->rndv) {
        shm_cell_hdr_t *c;
        if (r->progressed) return; /* RTS already out, waiting for ACK */
        c = ring_reserve(ring, &pos);
        if (!c) return;
        c->tag = r->tag; c->src_ep = ctx->addr.ep_id; c->total_len = r->len; c->offset = (uint64_t)(uintptr_t)r->buf;
        c->cookie = (uint64_t)(uintptr_t)r; c->len = 0; c->type = SHM_CELL_RTS; c->src_mt = (uint16_t)r->mt;
        ring_publish(c, pos);
        r->progressed = 1;
        return;
    }
    do {
        shm_cell_hdr_t *c = ring_reserve(ring, &pos);
        size_t chunk;
        if (!c) return;
        chunk = r->len - r->progressed; if (chunk > payload) chunk = payload;
        c->tag = r->tag; c->src_ep = ctx->addr.ep_id; c->total_len = r->len; c->offset = r->progressed; c->cookie = 0;
        c->len = (uint32_t)chunk; c->type = SHM_CELL_DATA; c->src_mt = 0;
        shm_copy(SHM_CELL_PAYLOAD(c), (char *)r->buf + r->progressed, chunk, UCC_MEMORY_TYPE_HOST, r->mt);
        ring_publish(c, pos);
        r->progressed += chunk;
    } while (r->progressed < r->len);
    shm_req_complete(r);
}

/* ACK of a rendezvous: a tiny control cell in the sender's ring.  Never waits: when that ring is full the ACK is parked on
 * ctx->pending_acks and flushed from progress (two contexts that spin for a cell in each other's full ring while holding
 * their locks - both rings full of eager cells behind an RTS - would never drain them). */
typedef struct shm_pending_ack { ucc_list_link_t list; ucc_tl_shm_ep_t *ep; uint64_t cookie; } shm_pending_ack_t;
static int ack_try(ucc_tl_shm_context_t *ctx, ucc_tl_shm_ep_t *ep, uint64_t cookie)
{
    uint64_t pos; shm_cell_hdr_t *c = ring_reserve(ep->ring, &pos);
    if (!c) return 0;
    c->tag = 0; c->src_ep = ctx->addr.ep_id; c->total_len = 0; c->offset = cookie; c->cookie = 0; c->len = 0; c->type = SHM_CELL_ACK; c->src_mt = 0;
    ring_publish(c, pos);
    return 1;
}
static void send_ack(ucc_tl_shm_context_t *ctx, ucc_tl_shm_ep_t *ep, uint64_t cookie)
{
    shm_pending_ack_t *pa;
    if (ucc_list_is_empty(&ctx->pending_acks) && ack_try(ctx, ep, cookie)) return;
    pa = (shm_pending_ack_t *)malloc(sizeof(*pa));
    pa->ep = ep; pa->cookie = cookie;
    ucc_list_add_tail(&ctx->pending_acks, &pa->list);
}
static void flush_acks(ucc_tl_shm_context_t *ctx)
{
    shm_pending_ack_t *pa, *t;
    ucc_list_for_each_safe(pa, t, &ctx->pending_acks, list) {
        if (!ack_try(ctx, pa->ep, pa->cookie)) break;
        ucc_list_del(&pa->list); free(pa);
    }
}

ucc_status_t ucc_tl_shm_send_nb(ucc_tl_shm_team_t *team, ucc_rank_t dst, uint64_t tag, void *buf, size_t len, ucc_memory_type_t mt, shm_req_t **req_p)
{
    ucc_tl_shm_context_t *ctx = SHM_CTX(team);
    ucc_tl_shm_ep_t *ep;
    shm_req_t *r;
    UCC_CHECK_RET(ucc_tl_shm_get_ep(team, dst, &ep));
    r = (shm_req_t *)ucc_mpool_get(&ctx->req_mp);
    if (!r) return UCC_ERR_NO_MEMORY;
    r->is_send = 1; r->done = 0; r->tag = tag; r->peer_ep = ep->addr.ep_id; r->buf = buf; r->len = len; r->progressed = 0; r->mt = mt; r->ep = ep;
    r->rndv = len >= ctx->cfg.rndv_thresh && (ep->same_process || (ep->cma_ok && mt == UCC_MEMORY_TYPE_HOST));
    CTX_LOCK(ctx);
    if (ucc_list_is_empty(&ctx->pending_sends)) send_push(ctx, r); /* keep per-destination FIFO order simple: never overtake */
    if (r->rndv) ucc_list_add_tail(r->progressed ? &ctx->rndv_sends : &ctx->pending_sends, &r->list);
    else if (!r->done) ucc_list_add_tail(&ctx->pending_sends, &r->list);
    CTX_UNLOCK(ctx);
    *req_p = r;
    return UCC_OK;
}

static shm_unexp_t *unexp_find(ucc_tl_shm_context_t *ctx, uint64_t src_ep, uint64_t tag)
{
    shm_unexp_t *u;
    ucc_list_for_each(u, &ctx->unexpected, list) if (u->src_ep == src_ep && u->tag == tag) return u;
    return NULL;
}

ucc_status_t ucc_tl_shm_recv_nb(ucc_tl_shm_team_t *team, ucc_rank_t src, uint64_t tag, void *buf, size_t len, ucc_memory_type_t mt, shm_req_t **req_p)
{
    ucc_tl_shm_context_t *ctx = SHM_CTX(team);
    ucc_tl_shm_ep_t *ep;
    shm_req_t *r;
    shm_unexp_t *u;
    UCC_CHECK_RET(ucc_tl_shm_get_ep(team, src, &ep));
    r = (shm_req_t *)ucc_mpool_get(&ctx->req_mp);
    if (!r) return UCC_ERR_NO_MEMORY;
    r->is_send = 0; r->done = 0; r->tag = tag; r->peer_ep = ep->addr.ep_id; r->buf = buf; r->len = len; r->progressed = 0; r->mt = mt; r->ep = ep; r->rndv = 0;
    CTX_LOCK(ctx);
    u = unexp_find(ctx, r->peer_ep, tag);
    if (u) {
        if (u->total_len > len) tl_error(UCC_TL_TEAM_LIB(team), "message truncated: incoming %llu bytes, receive buffer %zu", (unsigned long long)u->total_len, len);
        if (u->is_rts) {
            rndv_fetch(ctx, ep, buf, mt, u->rts_ptr, (ucc_memory_type_t)u->rts_mt, ucc_min((size_t)u->total_len, len));
            send_ack(ctx, ep, u->rts_cookie);
            r->progressed = u->total_len; shm_req_complete(r);
        } else {
            shm_copy(buf, u->data, ucc_min((size_t)u->received, len), mt, UCC_MEMORY_TYPE_HOST);
            r->progressed = u->received;
            r->len = u->total_len < len ? (size_t)u->total_len : len;
            if (u->received >= u->total_len) shm_req_complete(r);
        }
        ucc_list_del(&u->list); free(u->data); free(u);
    }
    if (!r->done) ucc_list_add_tail(&ctx->posted_recvs, &r->list);
    CTX_UNLOCK(ctx);
    *req_p = r;
    return UCC_OK;
}

void ucc_tl_shm_req_free(ucc_tl_shm_context_t *ctx, shm_req_t *req) { (void)ctx; ucc_mpool_put(req); }

static void handle_cell(ucc_tl_shm_context_t *ctx, shm_cell_hdr_t *c)
{
    shm_req_t *r;
    shm_unexp_t *u;
    if (c->type == SHM_CELL_ACK) {
        shm_req_t *s = (shm_req_t *)(uintptr_t)c->offset;
        ucc_list_for_each(r, &ctx->rndv_sends, list) if (r == s) { ucc_list_del(&r->list); shm_req_complete(r); return; }
        return;
    }
    ucc_list_for_each(r, &ctx->posted_recvs, list) {
        if (r->peer_ep != c->src_ep || r->tag != c->tag) continue;
        if (c->type == SHM_CELL_RTS) {
            rndv_fetch(ctx, r->ep, r->buf, r->mt, (void *)(uintptr_t)c->offset, (ucc_memory_type_t)c->src_mt, ucc_min((size_t)c->total_len, r->len));
            send_ack(ctx, r->ep, c->cookie);
            r->progressed = c->total_len;
        } else {
            if (c->offset + c->len <= r->len) shm_copy((char *)r->buf + c->offset, SHM_CELL_PAYLOAD(c), c->len, r->mt, UCC_MEMORY_TYPE_HOST);
            r->progressed += c->len;
        }
        if (r->progressed >= c->total_len) { ucc_list_del(&r->list); shm_req_complete(r); }
        return;
    }
    /* nobody waits for it yet: stash */
    /* a later cell of a message that is already being stashed joins it; a NEW message with the same (source, tag) - a
     * persistent collective posted again before the receiver caught up - must get its own entry (per-source FIFO order
     * guarantees the messages do not interleave) */
    u = NULL;
    if (c->type == SHM_CELL_DATA) {
        shm_unexp_t *it;
        ucc_list_for_each(it, &ctx->unexpected, list) if (it->src_ep == c->src_ep && it->tag == c->tag && !it->is_rts && it->received < it->total_len) { u = it; break; }
    }
    if (!u) {
        u = (shm_unexp_t *)calloc(1, sizeof(*u));
        u->tag = c->tag; u->src_ep = c->src_ep; u->total_len = c->total_len;
        if (c->type == SHM_CELL_RTS) { u->is_rts = 1; u->rts_ptr = (void *)(uintptr_t)c->offset; u->rts_cookie = c->cookie; u->rts_mt = c->src_mt; }
        else u->data = malloc(c->total_len ? c->total_len : 1);
        ucc_list_add_tail(&ctx->unexpected, &u->list);
    }
    if (c->type == SHM_CELL_DATA) { memcpy((char *)u->data + c->offset, SHM_CELL_PAYLOAD(c), c->len); u->received += c->len; }
}

ucc_status_t ucc_tl_shm_progress(void *arg)
{
    ucc_tl_shm_context_t *ctx = (ucc_tl_shm_context_t *)arg;
    shm_ring_hdr_t *ring = ctx->ring;
    shm_req_t *r, *t;
    CTX_LOCK(ctx);
    ctx->n_progress++;
    for (;;) { /* drain my ring */
        uint64_t pos = ring->head;
        shm_cell_hdr_t *c = SHM_CELL(ring, pos);
        if ((int64_t)ucc_load_acquire(&c->seq) - (int64_t)(pos + 1) != 0) break;
        ring->head = pos + 1; /* advance first: handle_cell may re-enter through send_ack on my own ring */
        handle_cell(ctx, c);
        ucc_store_release(&c->seq, pos + ring->n_cells);
    }
    flush_acks(ctx);
    ucc_list_for_each_safe(r, t, &ctx->pending_sends, list) {
        send_push(ctx, r);
        if (r->rndv) { if (r->progressed) { ucc_list_del(&r->list); ucc_list_add_tail(&ctx->rndv_sends, &r->list); } else break; }
        else if (r->done) ucc_list_del(&r->list);
        else break; /* ring full: keep order */
    }
    CTX_UNLOCK(ctx);
    return UCC_OK;
}
