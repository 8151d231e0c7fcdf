/* Collective operations front-end: init / post / triggered post / finalize.
 * Behaviour: reference core/ucc_coll.c:25-661 (zero-size shortcut, memory
 * type detection, asymmetric root staging, fallback-aware init through the
 * score map, executor life cycle, persistent re-post, stream-triggered post). */
#include "ucc_team.h"
#include "ucc_context.h"
#include "ucc_ee.h"
#include "ucc_global_opts.h"
#include "ucc_service_coll.h"
#include "components/mc/ucc_mc.h"
#include "components/ec/ucc_ec.h"
#include "utils/profile/ucc_profile.h"

#define COLL_TRACE(_lvl, _fmt, ...) ucc_log_component(_lvl, &ucc_global_config.coll_trace, _fmt, ##__VA_ARGS__)

/* ------------------------------------------------------------------ */
/* argument helpers                                                    */
/* ------------------------------------------------------------------ */
static int coll_is_zero_size(const ucc_coll_args_t *a, ucc_rank_t rank, ucc_rank_t size)
{
    int root = (ucc_rank_t)a->root == rank;
    (void)size;
    switch (a->coll_type) {
    case UCC_COLL_TYPE_ALLREDUCE: case UCC_COLL_TYPE_ALLGATHER: case UCC_COLL_TYPE_ALLTOALL: return a->dst.info.count == 0;
    case UCC_COLL_TYPE_BCAST: return a->src.info.count == 0;
    case UCC_COLL_TYPE_GATHER: return root ? a->dst.info.count == 0 : a->src.info.count == 0;
    case UCC_COLL_TYPE_REDUCE: return UCC_IS_INPLACE(*a) && root ? a->dst.info.count == 0 : (root ? a->dst.info.count == 0 : a->src.info.count == 0);
    case UCC_COLL_TYPE_SCATTER: return root ? a->src.info.count == 0 : a->dst.info.count == 0;
    default: return 0;
    }
}

static ucc_status_t detect_mem_type(void *ptr, ucc_memory_type_t *mt)
{
    ucc_mem_attr_t attr;
    ucc_status_t st;
    if (*mt != UCC_MEMORY_TYPE_UNKNOWN) return UCC_OK;
    if (!ptr) { *mt = UCC_MEMORY_TYPE_HOST; return UCC_OK; }
    attr.field_mask = UCC_MEM_ATTR_FIELD_MEM_TYPE;
    st = ucc_mc_get_mem_attr(ptr, &attr);
    if (st != UCC_OK) return st;
    *mt = attr.mem_type;
    return UCC_OK;
}

/* resolve UCC_MEMORY_TYPE_UNKNOWN on every buffer that is significant on this rank */
static ucc_status_t coll_args_fill_mem_types(ucc_coll_args_t *a, ucc_rank_t rank)
{
    int root = (ucc_rank_t)a->root == rank, inplace = UCC_IS_INPLACE(*a);
    switch (a->coll_type) {
    case UCC_COLL_TYPE_BARRIER: case UCC_COLL_TYPE_FANIN: case UCC_COLL_TYPE_FANOUT: return UCC_OK;
    case UCC_COLL_TYPE_BCAST: return detect_mem_type(a->src.info.buffer, &a->src.info.mem_type);
    case UCC_COLL_TYPE_ALLREDUCE: case UCC_COLL_TYPE_ALLGATHER: case UCC_COLL_TYPE_ALLTOALL: case UCC_COLL_TYPE_REDUCE_SCATTER:
        UCC_CHECK_RET(detect_mem_type(a->dst.info.buffer, &a->dst.info.mem_type));
        if (!inplace) UCC_CHECK_RET(detect_mem_type(a->src.info.buffer, &a->src.info.mem_type)); else a->src.info.mem_type = a->dst.info.mem_type;
        return UCC_OK;
    case UCC_COLL_TYPE_ALLGATHERV: case UCC_COLL_TYPE_REDUCE_SCATTERV:
        UCC_CHECK_RET(detect_mem_type(a->dst.info_v.buffer, &a->dst.info_v.mem_type));
        if (!inplace) UCC_CHECK_RET(detect_mem_type(a->src.info.buffer, &a->src.info.mem_type)); else a->src.info.mem_type = a->dst.info_v.mem_type;
        return UCC_OK;
    case UCC_COLL_TYPE_ALLTOALLV:
        UCC_CHECK_RET(detect_mem_type(a->dst.info_v.buffer, &a->dst.info_v.mem_type));
        if (!inplace) UCC_CHECK_RET(detect_mem_type(a->src.info_v.buffer, &a->src.info_v.mem_type)); else a->src.info_v.mem_type = a->dst.info_v.mem_type;
        return UCC_OK;
    case UCC_COLL_TYPE_REDUCE: case UCC_COLL_TYPE_GATHER:
        if (root) { UCC_CHECK_RET(detect_mem_type(a->dst.info.buffer, &a->dst.info.mem_type)); if (inplace) { a->src.info.mem_type = a->dst.info.mem_type; return UCC_OK; } }
        return detect_mem_type(a->src.info.buffer, &a->src.info.mem_type);
    case UCC_COLL_TYPE_GATHERV:
        if (root) { UCC_CHECK_RET(detect_mem_type(a->dst.info_v.buffer, &a->dst.info_v.mem_type)); if (inplace) { a->src.info.mem_type = a->dst.info_v.mem_type; return UCC_OK; } }
        return detect_mem_type(a->src.info.buffer, &a->src.info.mem_type);
    case UCC_COLL_TYPE_SCATTER:
        if (root) { UCC_CHECK_RET(detect_mem_type(a->src.info.buffer, &a->src.info.mem_type)); if (inplace) { a->dst.info.mem_type = a->src.info.mem_type; return UCC_OK; } }
        return detect_mem_type(a->dst.info.buffer, &a->dst.info.mem_type);
    case UCC_COLL_TYPE_SCATTERV:
        if (root) { UCC_CHECK_RET(detect_mem_type(a->src.info_v.buffer, &a->src.info_v.mem_type)); if (inplace) { a->dst.info.mem_type = a->src.info_v.mem_type; return UCC_OK; } }
        return detect_mem_type(a->dst.info.buffer, &a->dst.info.mem_type);
    default: return UCC_ERR_INVALID_PARAM;
    }
}

static ucc_status_t coll_args_check(const ucc_coll_args_t *a, ucc_rank_t rank, ucc_rank_t size)
{
    if (!a->coll_type || !ucc_is_pow2((uint64_t)a->coll_type) || ucc_coll_type_index(a->coll_type) >= UCC_COLL_TYPE_NUM) { ucc_error("invalid coll_type %d", (int)a->coll_type); return UCC_ERR_INVALID_PARAM; }
    if (ucc_coll_is_rooted(a->coll_type) && a->root >= size) { ucc_error("root %llu is out of team range [0,%u)", (unsigned long long)a->root, size); return UCC_ERR_INVALID_PARAM; }
    if (ucc_coll_has_reduction(a->coll_type) && a->op >= UCC_OP_LAST) { ucc_error("invalid reduction op %d", (int)a->op); return UCC_ERR_INVALID_PARAM; }
    if (UCC_COLL_ARGS_ACTIVE_SET(a)) {
        if (a->coll_type != UCC_COLL_TYPE_BCAST) { ucc_warn("active set is only supported with bcast"); return UCC_ERR_NOT_SUPPORTED; }
        if (a->active_set.size == 0 || a->active_set.size > size) return UCC_ERR_INVALID_PARAM;
    }
    if (a->coll_type == UCC_COLL_TYPE_ALLTOALL && !UCC_IS_INPLACE(*a) && a->src.info.count != a->dst.info.count &&
        ucc_dt_size(a->src.info.datatype) * a->src.info.count != ucc_dt_size(a->dst.info.datatype) * a->dst.info.count) {
        ucc_error("alltoall: src and dst sizes differ"); return UCC_ERR_INVALID_PARAM;
    }
    (void)rank;
    return UCC_OK;
}

/* ------------------------------------------------------------------ */
/* asymmetric memory types on rooted collectives                       */
/*   the root's "other side" buffer is staged through a scratch of the  */
/*   memory type every rank agrees on                                   */
/* ------------------------------------------------------------------ */
static ucc_status_t coll_asymm_setup(ucc_base_coll_args_t *b, ucc_rank_t rank, ucc_rank_t size)
{
    ucc_coll_args_t *a = &b->args;
    int root = (ucc_rank_t)a->root == rank;
    size_t len;
    ucc_status_t st;
    if (!root || UCC_IS_INPLACE(*a)) return UCC_OK;
    switch (a->coll_type) {
    case UCC_COLL_TYPE_REDUCE: case UCC_COLL_TYPE_GATHER: /* src is the "network" side, dst is root-only */
        if (a->src.info.mem_type == a->dst.info.mem_type) return UCC_OK;
        len = a->dst.info.count * ucc_dt_size(a->dst.info.datatype);
        st = ucc_mc_alloc(&b->asymm.scratch, len ? len : 1, a->src.info.mem_type);
        if (st != UCC_OK) return st;
        b->asymm.saved.info = a->dst.info; b->asymm.is_src = 0;
        a->dst.info.buffer = b->asymm.scratch->addr; a->dst.info.mem_type = a->src.info.mem_type;
        return UCC_OK;
    case UCC_COLL_TYPE_GATHERV:
        if (a->src.info.mem_type == a->dst.info_v.mem_type) return UCC_OK;
        { size_t max_end = 0; for (ucc_rank_t i = 0; i < size; i++) { size_t e = ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, i) + ucc_coll_args_get_count(a, a->dst.info_v.counts, i); if (e > max_end) max_end = e; }
          len = max_end * ucc_dt_size(a->dst.info_v.datatype); }
        st = ucc_mc_alloc(&b->asymm.scratch, len ? len : 1, a->src.info.mem_type);
        if (st != UCC_OK) return st;
        b->asymm.saved.info_v = a->dst.info_v; b->asymm.is_src = 0;
        a->dst.info_v.buffer = b->asymm.scratch->addr; a->dst.info_v.mem_type = a->src.info.mem_type;
        return UCC_OK;
    case UCC_COLL_TYPE_SCATTER:
        if (a->src.info.mem_type == a->dst.info.mem_type) return UCC_OK;
        len = a->src.info.count * ucc_dt_size(a->src.info.datatype);
        st = ucc_mc_alloc(&b->asymm.scratch, len ? len : 1, a->dst.info.mem_type);
        if (st != UCC_OK) return st;
        st = ucc_mc_memcpy(b->asymm.scratch->addr, a->src.info.buffer, len, a->dst.info.mem_type, a->src.info.mem_type);
        if (st != UCC_OK) { ucc_mc_free(b->asymm.scratch); b->asymm.scratch = NULL; return st; }
        b->asymm.saved.info = a->src.info; b->asymm.is_src = 1;
        a->src.info.buffer = b->asymm.scratch->addr; a->src.info.mem_type = a->dst.info.mem_type;
        return UCC_OK;
    case UCC_COLL_TYPE_SCATTERV:
        if (a->src.info_v.mem_type == a->dst.info.mem_type) return UCC_OK;
        { size_t max_end = 0; for (ucc_rank_t i = 0; i < size; i++) { size_t e = ucc_coll_args_get_displacement(a, a->src.info_v.displacements, i) + ucc_coll_args_get_count(a, a->src.info_v.counts, i); if (e > max_end) max_end = e; }
          len = max_end * ucc_dt_size(a->src.info_v.datatype); }
        st = ucc_mc_alloc(&b->asymm.scratch, len ? len : 1, a->dst.info.mem_type);
        if (st != UCC_OK) return st;
        st = ucc_mc_memcpy(b->asymm.scratch->addr, a->src.info_v.buffer, len, a->dst.info.mem_type, a->src.info_v.mem_type);
        if (st != UCC_OK) { ucc_mc_free(b->asymm.scratch); b->asymm.scratch = NULL; return st; }
        b->asymm.saved.info_v = a->src.info_v; b->asymm.is_src = 1;
        a->src.info_v.buffer = b->asymm.scratch->addr; a->src.info_v.mem_type = a->dst.info.mem_type;
        return UCC_OK;
    default: return UCC_OK;
    }
}

/* called from ucc_task_complete on the top-level task */
ucc_status_t ucc_coll_asymm_copy_out(ucc_coll_task_t *task)
{
    ucc_base_coll_args_t *b = &task->bargs;
    ucc_coll_args_t *a = &b->args;
    if (!b->asymm.scratch || b->asymm.is_src) return UCC_OK;
    if (a->coll_type == UCC_COLL_TYPE_GATHERV) {
        ucc_rank_t size = ucc_team_size_(b->team); size_t max_end = 0;
        for (ucc_rank_t i = 0; i < size; i++) { size_t e = ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, i) + ucc_coll_args_get_count(a, a->dst.info_v.counts, i); if (e > max_end) max_end = e; }
        return ucc_mc_memcpy(b->asymm.saved.info_v.buffer, b->asymm.scratch->addr, max_end * ucc_dt_size(a->dst.info_v.datatype),
                             b->asymm.saved.info_v.mem_type, a->dst.info_v.mem_type);
    }
    return ucc_mc_memcpy(b->asymm.saved.info.buffer, b->asymm.scratch->addr, a->dst.info.count * ucc_dt_size(a->dst.info.datatype),
                         b->asymm.saved.info.mem_type, a->dst.info.mem_type);
}

static void coll_asymm_release(ucc_base_coll_args_t *b)
{
    if (!b->asymm.scratch) return;
    ucc_mc_free(b->asymm.scratch); b->asymm.scratch = NULL;
}

/* ------------------------------------------------------------------ */
/* stub task for zero-size collectives                                 */
/* ------------------------------------------------------------------ */
static ucc_status_t stub_post(ucc_coll_task_t *task) { task->status = UCC_OK; return ucc_task_complete(task) < 0 ? task->status : UCC_OK; }
static ucc_status_t stub_triggered_post(ucc_ee_h ee, ucc_ev_t *ev, ucc_coll_task_t *task) { (void)ee; (void)ev; return stub_post(task); }
static ucc_status_t stub_finalize(ucc_coll_task_t *task) { ucc_coll_task_destruct(task); ucc_mpool_put(task); return UCC_OK; }

/* ------------------------------------------------------------------ */
/* init                                                                */
/* ------------------------------------------------------------------ */
UCC_EXPORT ucc_status_t ucc_collective_init(ucc_coll_args_t *coll_args, ucc_coll_req_h *request, ucc_team_h team)
{
    ucc_base_coll_args_t op_args;
    ucc_coll_task_t *task = NULL;
    ucc_status_t st;
    UCC_PROFILE_FUNC_BEGIN("ucc_collective_init");
    if (!coll_args || !request || !team) return UCC_ERR_INVALID_PARAM;
    if (ucc_unlikely(team->state != UCC_TEAM_ACTIVE)) { ucc_error("team %p is used before team_create is completed", (void *)team); return UCC_ERR_INVALID_PARAM; }
    st = coll_args_check(coll_args, team->rank, team->size);
    if (st != UCC_OK) return st;
    memset(&op_args, 0, sizeof(op_args));
    memcpy(&op_args.args, coll_args, sizeof(ucc_coll_args_t));
    op_args.team = team;
    if (!(op_args.args.mask & UCC_COLL_ARGS_FIELD_FLAGS)) op_args.args.flags = 0;
    op_args.args.mask |= UCC_COLL_ARGS_FIELD_FLAGS; /* flags is always valid internally */

    if (coll_is_zero_size(&op_args.args, team->rank, team->size)) {
        task = (ucc_coll_task_t *)ucc_mpool_get(&team->contexts[0]->lib->stub_tasks_mp);
        if (!task) return UCC_ERR_NO_MEMORY;
        ucc_coll_task_init(task, &op_args, NULL);
        task->post = stub_post; task->triggered_post = stub_triggered_post; task->finalize = stub_finalize;
        goto have_task;
    }
    st = coll_args_fill_mem_types(&op_args.args, team->rank);
    if (st != UCC_OK) { ucc_error("memory type detection failed"); return st; }
    st = coll_asymm_setup(&op_args, team->rank, team->size);
    if (st != UCC_OK) return st;

    st = ucc_coll_init(team->score_map, &op_args, &task);
    if (st == UCC_ERR_NOT_SUPPORTED) {
        char buf[256]; ucc_coll_args_str(&op_args.args, team->rank, team->size, buf, sizeof(buf));
        ucc_debug("failed to init collective: not supported: %s", buf);
        coll_asymm_release(&op_args);
        return st;
    }
    if (st < 0) { char buf[256]; ucc_coll_args_str(&op_args.args, team->rank, team->size, buf, sizeof(buf));
        ucc_error("failed to init collective: %s: %s", ucc_status_string(st), buf); coll_asymm_release(&op_args); return st; }
    task->bargs.asymm = op_args.asymm;

    if (ucc_global_config.check_asymmetric_dt && (op_args.args.coll_type & (UCC_COLL_TYPE_GATHERV | UCC_COLL_TYPE_SCATTERV | UCC_COLL_TYPE_GATHER | UCC_COLL_TYPE_SCATTER))) {
        /* optional cross-rank datatype / memory type agreement check, chained in front of the collective */
        ucc_coll_task_t *chk; ucc_schedule_t *sch; ucc_subset_t sub; int64_t v[4];
        ucc_memory_type_t mt = ucc_coll_args_mem_type(&op_args.args, team->rank);
        ucc_datatype_t dt = (op_args.args.coll_type & (UCC_COLL_TYPE_SCATTER | UCC_COLL_TYPE_SCATTERV)) ? op_args.args.dst.info.datatype : op_args.args.src.info.datatype;
        v[0] = (int64_t)dt; v[1] = -(int64_t)dt; v[2] = (int64_t)mt; v[3] = -(int64_t)mt;
        sub.map = ucc_ep_map_create_full(team->size); sub.myrank = team->rank;
        if (ucc_service_dt_check(team, v, sub, &chk) == UCC_OK) {
            sch = (ucc_schedule_t *)calloc(1, sizeof(*sch));
            ucc_schedule_init(sch, &op_args, NULL);
            ucc_schedule_add_task(sch, chk); ucc_schedule_add_task(sch, task);
            ucc_task_subscribe_dep(&sch->super, chk, UCC_EVENT_SCHEDULE_STARTED);
            ucc_task_subscribe_dep(chk, task, UCC_EVENT_COMPLETED);
            sch->super.bargs.asymm = op_args.asymm; task->bargs.asymm.scratch = NULL;
            sch->super.flags |= (task->flags & UCC_COLL_TASK_FLAG_EXECUTOR);
            sch->super.finalize = ucc_schedule_finalize; /* sub tasks finalized, schedule struct leaks nothing but itself */
            task = &sch->super;
        }
    }
have_task:
    task->flags |= UCC_COLL_TASK_FLAG_TOP_LEVEL;
    if (task->flags & UCC_COLL_TASK_FLAG_EXECUTOR) {
        ucc_ee_executor_params_t p;
        p.mask = UCC_EE_EXECUTOR_PARAM_FIELD_TYPE; p.ee_type = ucc_ee_type_from_mem_type(ucc_coll_args_mem_type(&op_args.args, team->rank));
        st = ucc_ee_executor_init(&p, &task->executor);
        if (st != UCC_OK) { ucc_error("failed to init executor: %s", ucc_status_string(st)); task->finalize(task); coll_asymm_release(&op_args); return st; }
    }
    if (op_args.args.mask & UCC_COLL_ARGS_FIELD_CB) { task->cb = op_args.args.cb; task->flags |= UCC_COLL_TASK_FLAG_CB; }
    task->seq_num = team->seq_num++;
    if (UCC_COLL_ARGS_HAS_TIMEOUT(&op_args.args)) task->timeout = op_args.args.timeout;
    *request = &task->super;
    if (ucc_global_config.coll_trace.log_level >= UCC_LOG_LEVEL_INFO && !(task->flags & UCC_COLL_TASK_FLAG_INTERNAL)) {
        if (team->rank == 0 || ucc_global_config.coll_trace.log_level >= UCC_LOG_LEVEL_DEBUG) {
            char buf[512]; ucc_coll_args_str(&op_args.args, team->rank, team->size, buf, sizeof(buf));
            /* a TL task was selected straight through cl/basic's score map; a CL task (cl/hier schedule) names itself */
            const char *tn = task->team ? ucc_base_team_name(task->team) : "-";
            int is_cl = !strncmp(tn, "CL_", 3) || !strncmp(tn, "cl_", 3);
            COLL_TRACE(UCC_LOG_LEVEL_INFO, "coll_init: %s; %s {%s}, team_id %u, req %p, seq_num %u", buf,
                       !task->team ? "STUB" : (is_cl ? tn : "CL_BASIC"), is_cl ? "schedule" : tn, team->id, (void *)task, task->seq_num);
        }
    }
    UCC_PROFILE_FUNC_END();
    return UCC_OK;
}

/* ------------------------------------------------------------------ */
/* post                                                                */
/* ------------------------------------------------------------------ */
static inline ucc_status_t check_postable(ucc_coll_task_t *task)
{
    ucc_status_t s = task->super.status;
    if (s == UCC_OPERATION_INITIALIZED) return UCC_OK;
    if (s == UCC_OK && UCC_IS_PERSISTENT(task->bargs.args)) return UCC_OK;
    ucc_error("attempt to post a request with invalid status: %s", ucc_status_string(s));
    return UCC_ERR_INVALID_PARAM;
}

UCC_EXPORT ucc_status_t ucc_collective_post(ucc_coll_req_h request)
{
    ucc_coll_task_t *task = ucc_derived_of(request, ucc_coll_task_t);
    ucc_status_t st;
    UCC_PROFILE_FUNC_BEGIN("ucc_collective_post");
    if (!request) return UCC_ERR_INVALID_PARAM;
    st = check_postable(task);
    if (st != UCC_OK) return st;
    if (ucc_global_config.coll_trace.log_level >= UCC_LOG_LEVEL_DEBUG && !(task->flags & UCC_COLL_TASK_FLAG_INTERNAL))
        COLL_TRACE(UCC_LOG_LEVEL_DEBUG, "coll_post: req %p, seq_num %u", (void *)task, task->seq_num);
    if (task->timeout > 0) task->start_time = ucc_get_time();
    if (task->executor) {
        st = ucc_ee_executor_start(task->executor, NULL);
        if (st != UCC_OK) { ucc_error("failed to start executor: %s", ucc_status_string(st)); return st; }
        task->flags |= UCC_COLL_TASK_FLAG_EXECUTOR_STOP;
    }
    st = task->post(task);
    UCC_PROFILE_FUNC_END();
    return st < 0 ? st : UCC_OK;
}

UCC_EXPORT ucc_status_t ucc_collective_init_and_post(ucc_coll_args_t *coll_args, ucc_coll_req_h *request, ucc_team_h team)
{
    /* the reference leaves this entry point unimplemented (core/ucc_coll.c:439-445); here it is the obvious composition */
    ucc_status_t st = ucc_collective_init(coll_args, request, team);
    if (st != UCC_OK) return st;
    st = ucc_collective_post(*request);
    if (st != UCC_OK) { ucc_collective_finalize(*request); *request = NULL; }
    return st;
}

UCC_EXPORT ucc_status_t ucc_collective_finalize(ucc_coll_req_h request)
{
    ucc_coll_task_t *task = ucc_derived_of(request, ucc_coll_task_t);
    ucc_status_t st;
    UCC_PROFILE_FUNC_BEGIN("ucc_collective_finalize");
    if (!request) return UCC_ERR_INVALID_PARAM;
    /* acquire: pairs with the release in ucc_task_complete, which may have run on another thread (THREAD_MULTIPLE) */
    if (ucc_load_acquire((volatile int *)&task->super.status) == UCC_INPROGRESS) { ucc_error("attempt to finalize an in-progress request"); return UCC_ERR_INVALID_PARAM; }
    if (ucc_global_config.coll_trace.log_level >= UCC_LOG_LEVEL_DEBUG && !(task->flags & UCC_COLL_TASK_FLAG_INTERNAL))
        COLL_TRACE(UCC_LOG_LEVEL_DEBUG, "coll_finalize: req %p, seq_num %u", (void *)task, task->seq_num);
    if (task->executor) { ucc_ee_executor_finalize(task->executor); task->executor = NULL; }
    coll_asymm_release(&task->bargs);
    if (task->flags & UCC_COLL_TASK_FLAG_IS_SCHEDULE && !task->team && !(task->flags & UCC_COLL_TASK_FLAG_IS_PIPELINED_SCHEDULE)) {
        /* dt-check wrapper schedule allocated in ucc_collective_init */
        st = task->finalize(task);
        free(task);
    } else st = task->finalize(task);
    UCC_PROFILE_FUNC_END();
    return st;
}

/* ------------------------------------------------------------------ */
/* triggered (stream ordered) post                                     */
/* ------------------------------------------------------------------ */
typedef struct ucc_ev_task { ucc_coll_task_t super; ucc_coll_task_t *coll_task; } ucc_ev_task_t;

static ucc_status_t ev_task_finalize(ucc_coll_task_t *t) { ucc_coll_task_destruct(t); free(t); return UCC_OK; }

/* the wrapped collective is posted once the trigger condition holds */
static ucc_status_t trigger_complete(ucc_coll_task_t *parent, ucc_coll_task_t *task)
{
    ucc_status_t st;
    (void)parent;
    if (task->timeout > 0) task->start_time = ucc_get_time();
    st = task->post(task);
    if (st < 0) { task->status = st; ucc_task_complete(task); }
    return UCC_OK;
}

static void trigger_test(ucc_coll_task_t *t)
{
    ucc_ev_task_t *evt = ucc_derived_of(t, ucc_ev_task_t);
    ucc_coll_task_t *coll = evt->coll_task;
    ucc_ev_t post_event;
    ucc_status_t st;
    /* For CUDA-stream EEs the compute-complete condition is implicit in stream order:
     * everything the collective enqueues on ee->ee_context runs after the user's prior work. */
    if (!t->ee_task) { /* first call: set the collective up for this stream */
        st = coll->triggered_post_setup(coll);
        if (st != UCC_OK) { t->status = st; return; }
        if ((coll->flags & UCC_COLL_TASK_FLAG_EXECUTOR) && coll->executor) {
            st = ucc_ee_executor_start(coll->executor, t->ee->ee_context);
            if (st != UCC_OK) { t->status = st; return; }
            coll->flags |= UCC_COLL_TASK_FLAG_EXECUTOR_STOP;
        }
        post_event.ev_type = UCC_EVENT_COLLECTIVE_POST; post_event.ev_context_size = 0; post_event.ev_context = NULL; post_event.req = &coll->super;
        ucc_ee_set_event_internal(t->ee, &post_event, &t->ee->event_out_queue);
        t->ee_task = (void *)1;
    }
    if (coll->executor && ucc_ee_executor_status(coll->executor) == UCC_INPROGRESS) return; /* persistent kernel not up yet */
    t->status = UCC_OK;
}

ucc_status_t ucc_triggered_post(ucc_ee_h ee, ucc_ev_t *ev, ucc_coll_task_t *task)
{
    ucc_ev_task_t *evt = (ucc_ev_task_t *)calloc(1, sizeof(*evt));
    ucc_context_t *ctx = task->bargs.team->contexts[0];
    if (!evt) return UCC_ERR_NO_MEMORY;
    ucc_coll_task_init(&evt->super, NULL, task->team);
    evt->coll_task = task;
    evt->super.ee = ee; evt->super.ev = ev; evt->super.ee_task = NULL;
    evt->super.flags = UCC_COLL_TASK_FLAG_INTERNAL | UCC_COLL_TASK_FLAG_AUTO_FINALIZE;
    evt->super.finalize = ev_task_finalize; evt->super.progress = trigger_test;
    evt->super.bargs.team = task->bargs.team;
    if (task->timeout > 0) { evt->super.timeout = task->timeout; evt->super.start_time = ucc_get_time(); }
    task->ee = ee;
    /* the wrapper posts the collective on completion and then releases itself (AUTO_FINALIZE) */
    ucc_event_manager_subscribe(&evt->super, UCC_EVENT_COMPLETED, task, trigger_complete);
    task->super.status = UCC_INPROGRESS;
    return ucc_progress_queue_enqueue(ctx->pq, &evt->super);
}

UCC_EXPORT ucc_status_t ucc_collective_triggered_post(ucc_ee_h ee, ucc_ev_t *ev)
{
    ucc_coll_task_t *task;
    ucc_status_t st;
    if (!ee || !ev || !ev->req) return UCC_ERR_INVALID_PARAM;
    task = ucc_derived_of(ev->req, ucc_coll_task_t);
    st = check_postable(task);
    if (st != UCC_OK) return st;
    if (ucc_global_config.coll_trace.log_level >= UCC_LOG_LEVEL_DEBUG)
        COLL_TRACE(UCC_LOG_LEVEL_DEBUG, "coll_triggered_post: req %p, seq_num %u", (void *)task, task->seq_num);
    return task->triggered_post(ee, ev, task);
}
