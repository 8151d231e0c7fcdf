#include "tl_shm_prog.h"
#include "utils/profile/ucc_profile.h"

static ucc_status_t op_push(ucc_tl_shm_task_t *t, const shm_op_t *op)
{
    if (t->n_ops == t->cap_ops) {
        unsigned ncap = t->cap_ops ? t->cap_ops * 2 : 32;
        shm_op_t *n = (shm_op_t *)realloc(t->ops, ncap * sizeof(*n));
        if (!n) return UCC_ERR_NO_MEMORY;
        t->ops = n; t->cap_ops = ncap;
    }
    t->ops[t->n_ops++] = *op;
    return UCC_OK;
}
ucc_status_t shm_prog_send(ucc_tl_shm_task_t *t, ucc_rank_t peer, const void *buf, size_t len, ucc_memory_type_t mt, unsigned step)
{ shm_op_t op; memset(&op, 0, sizeof(op)); op.type = SHM_OP_SEND; op.peer = peer; op.dst = (void *)buf; op.len = len; op.mt = mt; op.step = (uint16_t)step; return op_push(t, &op); }
ucc_status_t shm_prog_recv(ucc_tl_shm_task_t *t, ucc_rank_t peer, void *buf, size_t len, ucc_memory_type_t mt, unsigned step)
{ shm_op_t op; memset(&op, 0, sizeof(op)); op.type = SHM_OP_RECV; op.peer = peer; op.dst = buf; op.len = len; op.mt = mt; op.step = (uint16_t)step; return op_push(t, &op); }
ucc_status_t shm_prog_wait(ucc_tl_shm_task_t *t)
{ shm_op_t op; memset(&op, 0, sizeof(op)); op.type = SHM_OP_WAIT; if (t->n_ops && t->ops[t->n_ops - 1].type == SHM_OP_WAIT) return UCC_OK; return op_push(t, &op); }
ucc_status_t shm_prog_reduce(ucc_tl_shm_task_t *t, void *dst, const void *src1, const void *src2, size_t count, ucc_memory_type_t mt, int with_alpha)
{
    shm_op_t op; memset(&op, 0, sizeof(op));
    if (!count) return UCC_OK;
    op.type = SHM_OP_REDUCE; op.dst = dst; op.src1 = (void *)src1; op.src2 = (void *)src2; op.len = count; op.mt = mt; op.with_alpha = (uint8_t)with_alpha;
    return op_push(t, &op);
}
ucc_status_t shm_prog_copy(ucc_tl_shm_task_t *t, void *dst, const void *src, size_t len, ucc_memory_type_t dmt, ucc_memory_type_t smt)
{
    shm_op_t op; memset(&op, 0, sizeof(op));
    if (!len || dst == src) return UCC_OK;
    op.type = SHM_OP_COPY; op.dst = dst; op.src1 = (void *)src; op.len = len; op.mt = dmt; op.mt_src = smt;
    return op_push(t, &op);
}
ucc_status_t shm_prog_get_off(ucc_tl_shm_task_t *t, ucc_rank_t peer, void *dst, const uint64_t *remote_addr_slot, size_t off, size_t len, ucc_memory_type_t mt)
{
    shm_op_t op; memset(&op, 0, sizeof(op));
    if (!len) return UCC_OK;
    op.type = SHM_OP_GET; op.peer = peer; op.dst = dst; op.src1 = (void *)(uintptr_t)remote_addr_slot; op.src2 = (void *)(uintptr_t)off; op.len = len; op.mt = mt;
    return op_push(t, &op);
}
ucc_status_t shm_prog_get(ucc_tl_shm_task_t *t, ucc_rank_t peer, void *dst, const uint64_t *remote_addr_slot, size_t len, ucc_memory_type_t mt)
{ return shm_prog_get_off(t, peer, dst, remote_addr_slot, 0, len, mt); }
ucc_status_t shm_prog_put(ucc_tl_shm_task_t *t, ucc_rank_t peer, const void *src, uint64_t remote_addr, size_t len, ucc_memory_type_t smt)
{
    shm_op_t op; memset(&op, 0, sizeof(op));
    if (!len) return UCC_OK;
    op.type = SHM_OP_PUT; op.peer = peer; op.dst = (void *)(uintptr_t)src; op.src2 = (void *)(uintptr_t)remote_addr; op.len = len; op.mt = smt;
    return op_push(t, &op);
}
ucc_status_t shm_task_scratch(ucc_tl_shm_task_t *t, size_t len, ucc_memory_type_t mt, void **ptr)
{
    ucc_status_t st;
    if (t->n_scratch >= 4) return UCC_ERR_NO_RESOURCE;
    st = ucc_mc_alloc(&t->scratch[t->n_scratch], len ? len : 1, mt);
    if (st != UCC_OK) return st;
    *ptr = t->scratch[t->n_scratch++]->addr;
    return UCC_OK;
}

ucc_status_t ucc_tl_shm_task_alloc(ucc_base_coll_args_t *bargs, ucc_base_team_t *b_team, ucc_tl_shm_task_t **task_p)
{
    ucc_tl_shm_team_t *team = ucc_derived_of(b_team, ucc_tl_shm_team_t);
    ucc_tl_shm_context_t *ctx = SHM_CTX(team);
    ucc_tl_shm_task_t *t = (ucc_tl_shm_task_t *)ucc_mpool_get(&ctx->task_mp);
    if (!t) return UCC_ERR_NO_MEMORY;
    ucc_coll_task_init(&t->super, bargs, b_team);
    t->team = team; t->vrank = UCC_TL_TEAM_RANK(team); t->vsize = UCC_TL_TEAM_SIZE(team); t->vmap = ucc_ep_map_create_full(t->vsize);
    t->coll_seq = (team->seq_num++) & 0x7fff;
    t->ops = NULL; t->n_ops = t->cap_ops = t->pc = 0;
    t->reqs = NULL; t->n_reqs = t->cap_reqs = 0; t->etask = NULL; t->own_exec = NULL;
    t->dt = UCC_DT_INT8; t->op = UCC_OP_SUM; t->alpha = 1.0; t->n_scratch = 0; t->host_copy = NULL; t->build = NULL; t->setup = NULL;
    t->super.post = ucc_tl_shm_task_post; t->super.progress = ucc_tl_shm_task_progress; t->super.finalize = ucc_tl_shm_task_finalize;
    *task_p = t;
    return UCC_OK;
}

ucc_status_t ucc_tl_shm_task_finalize(ucc_coll_task_t *ct)
{
    ucc_tl_shm_task_t *t = ucc_derived_of(ct, ucc_tl_shm_task_t);
    for (unsigned i = 0; i < t->n_scratch; i++) ucc_mc_free(t->scratch[i]);
    if (t->own_exec) ucc_ee_executor_finalize(t->own_exec);
    free(t->ops); free(t->reqs); free(t->host_copy);
    ucc_coll_task_destruct(ct);
    ucc_mpool_put(t);
    return UCC_OK;
}

static ucc_status_t req_track(ucc_tl_shm_task_t *t, shm_req_t *r)
{
    if (t->n_reqs == t->cap_reqs) {
        unsigned ncap = t->cap_reqs ? t->cap_reqs * 2 : 16;
        shm_req_t **n = (shm_req_t **)realloc(t->reqs, ncap * sizeof(*n));
        if (!n) return UCC_ERR_NO_MEMORY;
        t->reqs = n; t->cap_reqs = ncap;
    }
    t->reqs[t->n_reqs++] = r;
    return UCC_OK;
}

static ucc_ee_executor_t *task_executor(ucc_tl_shm_task_t *t, ucc_memory_type_t mt)
{
    ucc_coll_task_t *c = &t->super;
    while (c) { if (c->executor) return c->executor; c = c->schedule ? &c->schedule->super : NULL; }
    if (!t->own_exec) {
        ucc_ee_executor_params_t p; p.mask = UCC_EE_EXECUTOR_PARAM_FIELD_TYPE; p.ee_type = ucc_ee_type_from_mem_type(mt);
        if (ucc_ee_executor_init(&p, &t->own_exec) != UCC_OK) return NULL;
        if (ucc_ee_executor_start(t->own_exec, NULL) != UCC_OK) { ucc_ee_executor_finalize(t->own_exec); t->own_exec = NULL; return NULL; }
    }
    return t->own_exec;
}

void ucc_tl_shm_task_progress(ucc_coll_task_t *ct)
{
    ucc_tl_shm_task_t *t = ucc_derived_of(ct, ucc_tl_shm_task_t);
    ucc_tl_shm_team_t *team = t->team;
    ucc_tl_shm_context_t *ctx = SHM_CTX(team);
    ucc_status_t st;
    for (;;) {
        shm_op_t *op;
        int at_end = t->pc >= t->n_ops;
        if (at_end || t->ops[t->pc].type == SHM_OP_WAIT) {
            int all_done = 1;
            if (t->n_reqs) {
                ucc_tl_shm_progress(ctx);
                for (unsigned i = 0; i < t->n_reqs; i++) if (!shm_req_is_done(t->reqs[i])) { all_done = 0; break; }
                if (!all_done) return;
                for (unsigned i = 0; i < t->n_reqs; i++) ucc_tl_shm_req_free(ctx, t->reqs[i]);
                t->n_reqs = 0;
            }
            if (at_end) {
                UCC_PROFILE_REQUEST_EVENT_NAMED(t, "shm", ucc_coll_type_str(ct->bargs.args.coll_type), "done");
                ct->status = UCC_OK; return;
            }
            t->pc++;
            continue;
        }
        op = &t->ops[t->pc];
        switch (op->type) {
        case SHM_OP_SEND: case SHM_OP_RECV: {
            shm_req_t *r;
            ucc_rank_t peer = ucc_ep_map_eval(t->vmap, op->peer);
            uint64_t tag = shm_make_tag(team, t->coll_seq, op->step);
            tl_trace(UCC_TL_TEAM_LIB(team), "[r%u/%u scope %d.%d] %s %s peer %u tag %lx buf %p len %zu", UCC_TL_TEAM_RANK(team), UCC_TL_TEAM_SIZE(team), team->super.super.params.scope, team->super.super.params.scope_id, ucc_coll_type_str(ct->bargs.args.coll_type), op->type == SHM_OP_SEND ? "send" : "recv", peer, (unsigned long)tag, op->dst, op->len);
            st = op->type == SHM_OP_SEND ? ucc_tl_shm_send_nb(team, peer, tag, op->dst, op->len, op->mt, &r)
                                         : ucc_tl_shm_recv_nb(team, peer, tag, op->dst, op->len, op->mt, &r);
            if (st == UCC_OK) st = req_track(t, r);
            if (st != UCC_OK) { ct->status = st; return; }
            t->pc++;
            break;
        }
        case SHM_OP_GET:
            st = ucc_tl_shm_get(team, ucc_ep_map_eval(t->vmap, op->peer), op->dst, op->mt, *(const uint64_t *)op->src1 + (uint64_t)(uintptr_t)op->src2, op->len);
            if (st != UCC_OK) { ct->status = st; return; }
            t->pc++;
            break;
        case SHM_OP_PUT:
            st = ucc_tl_shm_put(team, ucc_ep_map_eval(t->vmap, op->peer), op->dst, op->mt, (uint64_t)(uintptr_t)op->src2, op->len);
            if (st != UCC_OK) { ct->status = st; return; }
            t->pc++;
            break;
        case SHM_OP_REDUCE: case SHM_OP_COPY:
            if (!t->etask) {
                ucc_ee_executor_task_args_t a;
                ucc_ee_executor_t *exec = task_executor(t, op->mt);
                if (!exec) { ct->status = UCC_ERR_NO_RESOURCE; return; }
                memset(&a, 0, sizeof(a));
                if (op->type == SHM_OP_COPY) {
                    if (op->mt != op->mt_src) { /* cross memory-type copy: the MC knows how */
                        st = ucc_mc_memcpy(op->dst, op->src1, op->len, op->mt, op->mt_src);
                        if (st != UCC_OK) { ct->status = st; return; }
                        t->pc++;
                        break;
                    }
                    a.task_type = UCC_EE_EXECUTOR_TASK_COPY; a.copy.dst = op->dst; a.copy.src = op->src1; a.copy.len = op->len;
                } else {
                    a.task_type = UCC_EE_EXECUTOR_TASK_REDUCE; a.reduce.dst = op->dst; a.reduce.srcs[0] = op->src1; a.reduce.n_srcs = 1;
                    if (op->src2) { a.reduce.srcs[1] = op->src2; a.reduce.n_srcs = 2; }
                    a.reduce.count = op->len; a.reduce.dt = t->dt; a.reduce.op = t->op;
                    if (op->with_alpha) { a.flags |= UCC_EEE_TASK_FLAG_REDUCE_WITH_ALPHA; a.reduce.alpha = t->alpha; }
                    else if (t->op == UCC_OP_AVG) a.reduce.op = UCC_OP_SUM; /* intermediate steps of an average are plain sums */
                }
                st = ucc_ee_executor_task_post(exec, &a, &t->etask);
                if (st != UCC_OK) { ct->status = st; return; }
            }
            st = ucc_ee_executor_task_test(t->etask);
            if (st == UCC_INPROGRESS) return;
            ucc_ee_executor_task_finalize(t->etask); t->etask = NULL;
            if (st != UCC_OK) { ct->status = st; return; }
            t->pc++;
            break;
        default: ct->status = UCC_ERR_INVALID_PARAM; return;
        }
    }
}

ucc_status_t ucc_tl_shm_task_post(ucc_coll_task_t *ct)
{
    ucc_tl_shm_task_t *t = ucc_derived_of(ct, ucc_tl_shm_task_t);
    t->pc = 0; t->n_reqs = 0; t->etask = NULL;
    UCC_PROFILE_REQUEST_EVENT_NAMED(t, "shm", ucc_coll_type_str(ct->bargs.args.coll_type), "start");
    if (ucc_unlikely(ct->flags & UCC_COLL_TASK_FLAG_ARGS_UPDATED) && t->build) {
        /* a pipelined parent re-targeted the buffers / counts: rebuild the step program in place */
        ucc_status_t st;
        ct->flags &= ~UCC_COLL_TASK_FLAG_ARGS_UPDATED;
        for (unsigned i = 0; i < t->n_scratch; i++) ucc_mc_free(t->scratch[i]);
        t->n_scratch = 0; t->n_ops = 0; free(t->host_copy); t->host_copy = NULL;
        st = t->setup ? t->setup(t) : UCC_OK;
        if (st == UCC_OK) st = t->build(t);
        if (st != UCC_OK) { ct->status = st; return ucc_task_complete(ct); }
    }
    return ucc_progress_queue_enqueue(UCC_TL_CORE_CTX(t->team)->pq, ct);
}
