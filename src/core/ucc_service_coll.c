#include "ucc_service_coll.h"
#include "ucc_global_opts.h"

/* pick the TL team that serves `team`, translating the subset into its rank space */
static ucc_status_t service_prepare(ucc_team_t *team, ucc_subset_t subset, ucc_tl_team_t **steam, ucc_service_coll_req_t **req_p)
{
    ucc_context_t *ctx = team->contexts[0];
    ucc_service_coll_req_t *req;
    if (!team->service_team && !ctx->service_team) { ucc_error("service team is not available for team id %u", team->id); return UCC_ERR_NOT_SUPPORTED; }
    req = (ucc_service_coll_req_t *)calloc(1, sizeof(*req));
    if (!req) return UCC_ERR_NO_MEMORY;
    req->team = team;
    if (team->service_team) { *steam = team->service_team; req->subset = subset; }
    else {
        /* context-wide service team is indexed by ctx ranks */
        ucc_status_t st = ucc_ep_map_create_nested(&team->ctx_map, &subset.map, &req->subset.map);
        if (st != UCC_OK) { free(req); return st; }
        req->subset.myrank = subset.myrank; req->subset_owned = 1;
        *steam = ctx->service_team;
    }
    *req_p = req;
    return UCC_OK;
}

static ucc_status_t service_post(ucc_service_coll_req_t *req, ucc_status_t init_st, ucc_coll_task_t *task)
{
    ucc_status_t st = init_st;
    if (st == UCC_OK) {
        task->flags |= UCC_COLL_TASK_FLAG_INTERNAL;
        st = task->post(task);
        if (st < 0) task->finalize(task);
    }
    if (st < 0) { if (req->subset_owned) ucc_ep_map_destroy(&req->subset.map); free(req); return st; }
    req->task = task;
    return UCC_OK;
}

ucc_status_t ucc_service_allreduce(ucc_team_t *team, void *sbuf, void *rbuf, ucc_datatype_t dt, size_t count,
                                   ucc_reduction_op_t op, ucc_subset_t subset, ucc_service_coll_req_t **req_p)
{
    ucc_tl_team_t *steam; ucc_service_coll_req_t *req; ucc_coll_task_t *task = NULL; ucc_status_t st;
    UCC_CHECK_RET(service_prepare(team, subset, &steam, &req));
    st = UCC_TL_TEAM_IFACE(steam)->scoll.allreduce(&steam->super, sbuf, rbuf, dt, count, op, req->subset, &task);
    st = service_post(req, st, task);
    if (st == UCC_OK) *req_p = req;
    return st;
}
ucc_status_t ucc_service_allgather(ucc_team_t *team, void *sbuf, void *rbuf, size_t msgsize, ucc_subset_t subset, ucc_service_coll_req_t **req_p)
{
    ucc_tl_team_t *steam; ucc_service_coll_req_t *req; ucc_coll_task_t *task = NULL; ucc_status_t st;
    UCC_CHECK_RET(service_prepare(team, subset, &steam, &req));
    st = UCC_TL_TEAM_IFACE(steam)->scoll.allgather(&steam->super, sbuf, rbuf, msgsize, req->subset, &task);
    st = service_post(req, st, task);
    if (st == UCC_OK) *req_p = req;
    return st;
}
ucc_status_t ucc_service_bcast(ucc_team_t *team, void *buf, size_t msgsize, ucc_rank_t root, ucc_subset_t subset, ucc_service_coll_req_t **req_p)
{
    ucc_tl_team_t *steam; ucc_service_coll_req_t *req; ucc_coll_task_t *task = NULL; ucc_status_t st;
    UCC_CHECK_RET(service_prepare(team, subset, &steam, &req));
    st = UCC_TL_TEAM_IFACE(steam)->scoll.bcast(&steam->super, buf, msgsize, root, req->subset, &task);
    st = service_post(req, st, task);
    if (st == UCC_OK) *req_p = req;
    return st;
}
ucc_status_t ucc_service_coll_test(ucc_service_coll_req_t *req)
{
    ucc_status_t st = req->task->super.status;
    if (st == UCC_INPROGRESS) { ucc_context_progress(req->team->contexts[0]); st = req->task->super.status; }
    return st;
}
ucc_status_t ucc_service_coll_finalize(ucc_service_coll_req_t *req)
{
    ucc_status_t st = req->task->finalize(req->task);
    if (req->subset_owned) ucc_ep_map_destroy(&req->subset.map);
    free(req->data); free(req);
    return st;
}

/* ---- internal OOB ---- */
typedef struct ucc_internal_oob_coll_info { ucc_team_t *team; ucc_subset_t subset; } ucc_internal_oob_coll_info_t;
static ucc_status_t internal_oob_allgather(void *sbuf, void *rbuf, size_t size, void *coll_info, void **request)
{
    ucc_internal_oob_coll_info_t *ci = (ucc_internal_oob_coll_info_t *)coll_info;
    ucc_service_coll_req_t *req = NULL;
    ucc_status_t st = ucc_service_allgather(ci->team, sbuf, rbuf, size, ci->subset, &req);
    *request = req;
    return st;
}
static ucc_status_t internal_oob_test(void *request) { return ucc_service_coll_test((ucc_service_coll_req_t *)request); }
static ucc_status_t internal_oob_free(void *request) { return ucc_service_coll_finalize((ucc_service_coll_req_t *)request); }

ucc_status_t ucc_internal_oob_init(ucc_team_t *team, ucc_subset_t subset, ucc_team_oob_coll_t *oob)
{
    ucc_internal_oob_coll_info_t *ci = (ucc_internal_oob_coll_info_t *)malloc(sizeof(*ci));
    if (!ci) return UCC_ERR_NO_MEMORY;
    ci->team = team; ci->subset = subset;
    oob->coll_info = ci; oob->allgather = internal_oob_allgather; oob->req_test = internal_oob_test; oob->req_free = internal_oob_free;
    oob->n_oob_eps = (uint32_t)subset.map.ep_num; oob->oob_ep = subset.myrank;
    return UCC_OK;
}
void ucc_internal_oob_finalize(ucc_team_oob_coll_t *oob) { free(oob->coll_info); oob->coll_info = NULL; }

/* ---- dt / mem-type consistency check -------------------------------
 * values = {dt, -dt, mem, -mem}; after allreduce(MIN) a mismatch shows up
 * as values[0] != -values[1] (or [2] != -[3]).  The check is a task so it
 * composes with the user collective in a schedule. */
typedef struct dt_check_task { ucc_coll_task_t super; ucc_service_coll_req_t *req; int64_t in[4], out[4]; ucc_team_t *core_team; ucc_subset_t subset; } dt_check_task_t;
static void dt_check_progress(ucc_coll_task_t *t)
{
    dt_check_task_t *c = ucc_derived_of(t, dt_check_task_t);
    ucc_status_t st = c->req->task->super.status;
    if (st == UCC_INPROGRESS) return;
    if (st == UCC_OK && (c->out[0] != -c->out[1] || c->out[2] != -c->out[3])) {
        ucc_error("datatype or memory type mismatch across ranks (CHECK_ASYMMETRIC_DT)");
        st = UCC_ERR_INVALID_PARAM;
    }
    t->status = st;
}
static ucc_status_t dt_check_post(ucc_coll_task_t *t)
{
    dt_check_task_t *c = ucc_derived_of(t, dt_check_task_t);
    ucc_status_t st = ucc_service_allreduce(c->core_team, c->in, c->out, UCC_DT_INT64, 4, UCC_OP_MIN, c->subset, &c->req);
    if (st != UCC_OK) return st;
    return ucc_progress_queue_enqueue(c->core_team->contexts[0]->pq, t);
}
static ucc_status_t dt_check_finalize(ucc_coll_task_t *t)
{
    dt_check_task_t *c = ucc_derived_of(t, dt_check_task_t);
    if (c->req) ucc_service_coll_finalize(c->req);
    ucc_coll_task_destruct(t); free(c);
    return UCC_OK;
}
ucc_status_t ucc_service_dt_check(ucc_team_t *team, int64_t *values, ucc_subset_t subset, ucc_coll_task_t **check_task)
{
    dt_check_task_t *c = (dt_check_task_t *)calloc(1, sizeof(*c));
    ucc_base_coll_args_t ba;
    if (!c) return UCC_ERR_NO_MEMORY;
    memset(&ba, 0, sizeof(ba)); ba.team = team; ba.args.coll_type = UCC_COLL_TYPE_ALLREDUCE;
    ucc_coll_task_init(&c->super, &ba, NULL);
    memcpy(c->in, values, sizeof(c->in));
    c->core_team = team; c->subset = subset;
    c->super.post = dt_check_post; c->super.progress = dt_check_progress; c->super.finalize = dt_check_finalize;
    c->super.flags |= UCC_COLL_TASK_FLAG_INTERNAL;
    *check_task = &c->super;
    return UCC_OK;
}
