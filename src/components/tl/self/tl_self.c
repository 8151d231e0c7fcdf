/* tl/self: transport for teams of size 1 — every collective degenerates to
 * a local copy (or nothing).  Role of reference tl/self (708 LoC). */
#include "components/tl/ucc_tl.h"
#include "components/mc/ucc_mc.h"
#include "core/ucc_context.h"
#include "core/ucc_team.h"
#include "utils/ucc_mpool.h"

#define UCC_TL_SELF_DEFAULT_SCORE 50
#define UCC_TL_SELF_SUPPORTED_COLLS UCC_COLL_TYPE_ALL

typedef struct ucc_tl_self_lib { ucc_tl_lib_t super; } ucc_tl_self_lib_t;
typedef struct ucc_tl_self_context { ucc_tl_context_t super; ucc_tl_context_config_t cfg; char *tune; ucc_mpool_t req_mp; } ucc_tl_self_context_t;
typedef struct ucc_tl_self_team { ucc_tl_team_t super; } ucc_tl_self_team_t;
typedef struct ucc_tl_self_task { ucc_coll_task_t super; void *src, *dst; size_t size; ucc_memory_type_t src_mt, dst_mt; } ucc_tl_self_task_t;

extern ucc_tl_iface_t ucc_tl_self;

static ucc_config_field_t tl_self_lib_config_table[] = {
    {"", "", NULL, 0, UCC_CONFIG_TYPE_TABLE(ucc_tl_lib_config_table)}, {NULL}};
static ucc_config_field_t tl_self_context_config_table[] = {
    {"", "", NULL, 0, UCC_CONFIG_TYPE_TABLE(ucc_tl_context_config_table)}, {NULL}};

static ucc_status_t self_lib_init(const ucc_base_lib_params_t *p, const ucc_base_lib_config_t *config, ucc_base_lib_t **lib_p)
{
    ucc_tl_self_lib_t *lib = (ucc_tl_self_lib_t *)calloc(1, sizeof(*lib));
    (void)p;
    if (!lib) return UCC_ERR_NO_MEMORY;
    ucc_tl_lib_init_base(&lib->super, &ucc_tl_self, ucc_derived_of(config, ucc_tl_lib_config_t));
    *lib_p = &lib->super.super;
    return UCC_OK;
}
static void self_lib_finalize(ucc_base_lib_t *lib) { free(lib); }
static ucc_status_t self_lib_get_attr(const ucc_base_lib_t *lib, ucc_base_lib_attr_t *attr)
{ (void)lib; attr->attr.thread_mode = UCC_THREAD_MULTIPLE; attr->attr.coll_types = UCC_TL_SELF_SUPPORTED_COLLS; attr->flags = 0; attr->min_team_size = 1; attr->max_team_size = 1; return UCC_OK; }
static ucc_status_t self_lib_get_properties(ucc_base_lib_properties_t *p) { p->default_team_size = 1; p->min_team_size = 1; p->max_team_size = 1; return UCC_OK; }

static ucc_status_t self_ctx_create(const ucc_base_context_params_t *p, const ucc_base_ctx_config_t *config, ucc_base_context_t **ctx_p)
{
    ucc_tl_self_context_t *ctx = (ucc_tl_self_context_t *)calloc(1, sizeof(*ctx));
    if (!ctx) return UCC_ERR_NO_MEMORY;
    ctx->super.super.ucc_context = p->context; ctx->super.super.lib = config->lib;
    ctx->tune = config->score_str ? strdup(config->score_str) : NULL;
    ucc_mpool_init(&ctx->req_mp, 0, sizeof(ucc_tl_self_task_t), 0, 64, 8, (unsigned)-1, NULL, p->thread_mode, "tl_self_req_mp");
    *ctx_p = &ctx->super.super;
    return UCC_OK;
}
static void self_ctx_destroy(ucc_base_context_t *b) { ucc_tl_self_context_t *ctx = ucc_derived_of(b, ucc_tl_self_context_t); ucc_mpool_cleanup(&ctx->req_mp, 1); free(ctx->tune); free(ctx); }
static ucc_status_t self_ctx_get_attr(const ucc_base_context_t *ctx, ucc_base_ctx_attr_t *attr)
{ (void)ctx; if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN) attr->attr.ctx_addr_len = 0; attr->topo_required = 0; attr->attr.global_work_buffer_size = 0; return UCC_OK; }

static ucc_status_t self_team_create_post(ucc_base_context_t *ctx, const ucc_base_team_params_t *params, ucc_base_team_t **team_p)
{
    ucc_tl_self_team_t *team;
    if (params->size != 1) return UCC_ERR_NOT_SUPPORTED;
    team = (ucc_tl_self_team_t *)calloc(1, sizeof(*team));
    if (!team) return UCC_ERR_NO_MEMORY;
    team->super.super.context = ctx; team->super.super.params = *params;
    *team_p = &team->super.super;
    return UCC_OK;
}
static ucc_status_t self_team_create_test(ucc_base_team_t *t) { (void)t; return UCC_OK; }
static ucc_status_t self_team_destroy(ucc_base_team_t *t) { free(t); return UCC_OK; }

/* ---- collectives ---- */
static ucc_status_t self_finalize(ucc_coll_task_t *t) { ucc_coll_task_destruct(t); ucc_mpool_put(t); return UCC_OK; }
static ucc_status_t self_post(ucc_coll_task_t *t)
{
    ucc_tl_self_task_t *task = ucc_derived_of(t, ucc_tl_self_task_t);
    ucc_status_t st = UCC_OK;
    if (task->size && task->src != task->dst && task->src && task->dst) st = ucc_mc_memcpy(task->dst, task->src, task->size, task->dst_mt, task->src_mt);
    t->status = st;
    t->super.status = UCC_INPROGRESS;
    return ucc_task_complete(t) < 0 ? t->status : UCC_OK;
}
static ucc_status_t self_triggered_post(ucc_ee_h ee, ucc_ev_t *ev, ucc_coll_task_t *t) { (void)ee; (void)ev; return self_post(t); }

static ucc_status_t self_coll_init(ucc_base_coll_args_t *b, ucc_base_team_t *team, ucc_coll_task_t **task_p)
{
    ucc_tl_self_context_t *ctx = ucc_derived_of(team->context, ucc_tl_self_context_t);
    ucc_tl_self_task_t *task = (ucc_tl_self_task_t *)ucc_mpool_get(&ctx->req_mp);
    ucc_coll_args_t *a = &b->args;
    if (!task) return UCC_ERR_NO_MEMORY;
    ucc_coll_task_init(&task->super, b, team);
    task->super.post = self_post; task->super.finalize = self_finalize; task->super.triggered_post = self_triggered_post;
    task->src = task->dst = NULL; task->size = 0; task->src_mt = task->dst_mt = UCC_MEMORY_TYPE_HOST;
    if (!UCC_IS_INPLACE(*a)) {
        switch (a->coll_type) {
        case UCC_COLL_TYPE_ALLGATHER: case UCC_COLL_TYPE_ALLREDUCE: case UCC_COLL_TYPE_ALLTOALL: case UCC_COLL_TYPE_GATHER:
        case UCC_COLL_TYPE_REDUCE: case UCC_COLL_TYPE_REDUCE_SCATTER: case UCC_COLL_TYPE_SCATTER:
            task->src = a->src.info.buffer; task->dst = a->dst.info.buffer; task->size = a->src.info.count * ucc_dt_size(a->src.info.datatype);
            if (a->coll_type == UCC_COLL_TYPE_SCATTER || a->coll_type == UCC_COLL_TYPE_REDUCE_SCATTER) task->size = a->dst.info.count * ucc_dt_size(a->dst.info.datatype);
            task->src_mt = a->src.info.mem_type; task->dst_mt = a->dst.info.mem_type;
            break;
        case UCC_COLL_TYPE_ALLGATHERV: case UCC_COLL_TYPE_GATHERV: case UCC_COLL_TYPE_REDUCE_SCATTERV:
            task->src = a->src.info.buffer;
            task->dst = PTR_OFFSET(a->dst.info_v.buffer, (a->dst.info_v.displacements ? ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, 0) : 0) * ucc_dt_size(a->dst.info_v.datatype));
            task->size = ucc_coll_args_get_count(a, a->dst.info_v.counts, 0) * ucc_dt_size(a->dst.info_v.datatype);
            task->src_mt = a->src.info.mem_type; task->dst_mt = a->dst.info_v.mem_type;
            break;
        case UCC_COLL_TYPE_ALLTOALLV:
            task->src = PTR_OFFSET(a->src.info_v.buffer, ucc_coll_args_get_displacement(a, a->src.info_v.displacements, 0) * ucc_dt_size(a->src.info_v.datatype));
            task->dst = PTR_OFFSET(a->dst.info_v.buffer, ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, 0) * ucc_dt_size(a->dst.info_v.datatype));
            task->size = ucc_coll_args_get_count(a, a->src.info_v.counts, 0) * ucc_dt_size(a->src.info_v.datatype);
            task->src_mt = a->src.info_v.mem_type; task->dst_mt = a->dst.info_v.mem_type;
            break;
        case UCC_COLL_TYPE_SCATTERV:
            task->src = PTR_OFFSET(a->src.info_v.buffer, (a->src.info_v.displacements ? ucc_coll_args_get_displacement(a, a->src.info_v.displacements, 0) : 0) * ucc_dt_size(a->src.info_v.datatype));
            task->dst = a->dst.info.buffer; task->size = a->dst.info.count * ucc_dt_size(a->dst.info.datatype);
            task->src_mt = a->src.info_v.mem_type; task->dst_mt = a->dst.info.mem_type;
            break;
        default: break; /* barrier, bcast, fanin, fanout: nothing to move */
        }
    }
    *task_p = &task->super;
    return UCC_OK;
}

static ucc_status_t self_team_get_scores(ucc_base_team_t *b_team, ucc_coll_score_t **score_p)
{
    ucc_tl_self_team_t *team = ucc_derived_of(b_team, ucc_tl_self_team_t);
    ucc_tl_self_context_t *ctx = ucc_derived_of(b_team->context, ucc_tl_self_context_t);
    ucc_memory_type_t mt[] = {UCC_MEMORY_TYPE_HOST, UCC_MEMORY_TYPE_CUDA, UCC_MEMORY_TYPE_CUDA_MANAGED, UCC_MEMORY_TYPE_ROCM, UCC_MEMORY_TYPE_ROCM_MANAGED};
    ucc_coll_score_team_info_t info = {UCC_TL_SELF_DEFAULT_SCORE, 1, UCC_TL_SELF_SUPPORTED_COLLS, mt, 5, self_coll_init, NULL};
    ucc_coll_score_t *score;
    ucc_status_t st = ucc_coll_score_build_default(b_team, UCC_TL_SELF_DEFAULT_SCORE, self_coll_init, UCC_TL_SELF_SUPPORTED_COLLS, mt, 5, &score);
    if (st != UCC_OK) return st;
    st = ucc_tl_apply_tune(&team->super, score, &info, NULL, ctx->tune);
    if (st != UCC_OK) { ucc_coll_score_free(score); return st; }
    *score_p = score;
    return UCC_OK;
}

ucc_tl_iface_t ucc_tl_self = {
    .super = {.name = "self", .score = UCC_TL_SELF_DEFAULT_SCORE},
    .tl_lib_config = {"TL_SELF lib", "TL_SELF_", tl_self_lib_config_table, sizeof(ucc_tl_lib_config_t), {NULL, NULL}},
    .tl_context_config = {"TL_SELF context", "TL_SELF_", tl_self_context_config_table, sizeof(ucc_tl_context_config_t), {NULL, NULL}},
    .lib = {self_lib_init, self_lib_finalize, self_lib_get_attr, self_lib_get_properties},
    .context = {self_ctx_create, NULL, self_ctx_destroy, self_ctx_get_attr, NULL, NULL, NULL},
    .team = {self_team_create_post, self_team_create_test, self_team_destroy, self_team_get_scores},
    .coll = {self_coll_init},
};
UCC_COMPONENT_REGISTER_STATIC(tl, ucc_tl_self)
static void UCC_CTOR tl_self_cfg_register(void) { ucc_config_table_register(&ucc_tl_self.tl_lib_config); ucc_config_table_register(&ucc_tl_self.tl_context_config); }
