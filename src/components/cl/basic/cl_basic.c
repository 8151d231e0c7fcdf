/* cl/basic: the flat collective layer — one TL team per usable TL over the
 * whole team, TL scores merged into one selection map (reference cl/basic, 585 LoC). */
#include "components/cl/ucc_cl.h"
#include "components/tl/ucc_tl.h"
#include "core/ucc_context.h"
#include "core/ucc_team.h"
#include "core/ucc_global_opts.h"

#define UCC_CL_BASIC_DEFAULT_SCORE 10
typedef struct ucc_cl_basic_lib { ucc_cl_lib_t super; } ucc_cl_basic_lib_t;
typedef struct ucc_cl_basic_context { ucc_cl_context_t super; ucc_tl_context_t **tl_ctxs; unsigned n_tl_ctxs; char *tune; } ucc_cl_basic_context_t;
typedef struct ucc_cl_basic_team {
    ucc_cl_team_t            super;
    ucc_team_multiple_req_t *team_create_req;
    ucc_tl_team_t          **tl_teams;
    unsigned                 n_tl_teams;
    ucc_coll_score_t        *score;
    ucc_score_map_t         *score_map;
} ucc_cl_basic_team_t;
extern ucc_cl_iface_t ucc_cl_basic;

static ucc_config_field_t cl_basic_lib_config_table[] = {{"", "", NULL, 0, UCC_CONFIG_TYPE_TABLE(ucc_cl_lib_config_table)}, {NULL}};
static ucc_config_field_t cl_basic_context_config_table[] = {{"", "", NULL, 0, UCC_CONFIG_TYPE_TABLE(ucc_cl_context_config_table)}, {NULL}};

static ucc_status_t basic_lib_init(const ucc_base_lib_params_t *p, const ucc_base_lib_config_t *config, ucc_base_lib_t **lib_p)
{
    ucc_cl_basic_lib_t *lib = (ucc_cl_basic_lib_t *)calloc(1, sizeof(*lib));
    ucc_status_t st;
    (void)p;
    if (!lib) return UCC_ERR_NO_MEMORY;
    st = ucc_cl_lib_init_base(&lib->super, &ucc_cl_basic, ucc_derived_of(config, ucc_cl_lib_config_t));
    if (st != UCC_OK) { free(lib); return st; }
    cl_debug(&lib->super, "initialized lib object: %p", (void *)lib);
    *lib_p = &lib->super.super;
    return UCC_OK;
}
static void basic_lib_finalize(ucc_base_lib_t *b) { ucc_cl_basic_lib_t *lib = ucc_derived_of(b, ucc_cl_basic_lib_t); ucc_cl_lib_cleanup_base(&lib->super); free(lib); }
static ucc_status_t basic_lib_get_attr(const ucc_base_lib_t *b, ucc_base_lib_attr_t *battr)
{
    const ucc_cl_basic_lib_t *lib = ucc_derived_of(b, ucc_cl_basic_lib_t);
    ucc_cl_lib_attr_t *attr = ucc_derived_of(battr, ucc_cl_lib_attr_t);
    /* capabilities = union over the TLs we may use; thread mode = weakest of them */
    attr->super.attr.thread_mode = UCC_THREAD_MULTIPLE; attr->super.attr.coll_types = 0; attr->super.flags = 0;
    for (unsigned i = 0; i < lib->super.tls.array.count; i++) {
        ucc_tl_iface_t *tl = ucc_derived_of(ucc_get_component(&ucc_global_config.tl_framework, lib->super.tls.array.names[i]), ucc_tl_iface_t);
        ucc_base_lib_attr_t a;
        if (!tl) continue;
        memset(&a, 0, sizeof(a)); a.attr.thread_mode = UCC_THREAD_MULTIPLE;
        if (tl->lib.get_attr) tl->lib.get_attr(NULL, &a);
        attr->super.attr.coll_types |= a.attr.coll_types;
        if (a.attr.thread_mode < attr->super.attr.thread_mode) attr->super.attr.thread_mode = a.attr.thread_mode;
        attr->super.flags |= a.flags;
    }
    attr->tls = (ucc_config_names_array_t *)&lib->super.tls.array; attr->tls_forced = lib->super.tls_forced;
    return UCC_OK;
}

static ucc_status_t basic_ctx_create(const ucc_base_context_params_t *p, const ucc_base_ctx_config_t *config, ucc_base_context_t **ctx_p)
{
    ucc_cl_lib_t *lib = ucc_derived_of(config->lib, ucc_cl_lib_t);
    ucc_cl_basic_context_t *ctx = (ucc_cl_basic_context_t *)calloc(1, sizeof(*ctx));
    if (!ctx) return UCC_ERR_NO_MEMORY;
    ctx->super.super.ucc_context = p->context; ctx->super.super.lib = config->lib;
    ctx->tune = config->score_str ? strdup(config->score_str) : NULL;
    ctx->tl_ctxs = (ucc_tl_context_t **)calloc(lib->tls.array.count + 1, sizeof(void *));
    for (unsigned i = 0; i < lib->tls.array.count; i++) {
        ucc_status_t st = ucc_tl_context_get(p->context, lib->tls.array.names[i], &ctx->tl_ctxs[ctx->n_tl_ctxs]);
        if (st != UCC_OK) { cl_debug(lib, "TL %s context is not available, skipping", lib->tls.array.names[i]); continue; }
        ctx->n_tl_ctxs++;
    }
    if (ctx->n_tl_ctxs == 0) { cl_error(lib, "no TL contexts are available"); free(ctx->tl_ctxs); free(ctx->tune); free(ctx); return UCC_ERR_NOT_FOUND; }
    *ctx_p = &ctx->super.super;
    return UCC_OK;
}
static void basic_ctx_destroy(ucc_base_context_t *b)
{
    ucc_cl_basic_context_t *ctx = ucc_derived_of(b, ucc_cl_basic_context_t);
    for (unsigned i = 0; i < ctx->n_tl_ctxs; i++) ucc_tl_context_put(ctx->tl_ctxs[i]);
    free(ctx->tl_ctxs); free(ctx->tune); free(ctx);
}
static ucc_status_t basic_ctx_get_attr(const ucc_base_context_t *b, ucc_base_ctx_attr_t *attr)
{
    const ucc_cl_basic_context_t *ctx = ucc_derived_of(b, ucc_cl_basic_context_t);
    if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN) attr->attr.ctx_addr_len = 0;
    attr->topo_required = 0;
    for (unsigned i = 0; i < ctx->n_tl_ctxs; i++) {
        ucc_base_ctx_attr_t a; memset(&a, 0, sizeof(a));
        UCC_TL_CTX_IFACE(ctx->tl_ctxs[i])->context.get_attr(&ctx->tl_ctxs[i]->super, &a);
        attr->topo_required |= a.topo_required;
    }
    return UCC_OK;
}

static ucc_status_t basic_team_create_post(ucc_base_context_t *b_ctx, const ucc_base_team_params_t *params, ucc_base_team_t **team_p)
{
    ucc_cl_basic_context_t *ctx = ucc_derived_of(b_ctx, ucc_cl_basic_context_t);
    ucc_cl_basic_team_t *team = (ucc_cl_basic_team_t *)calloc(1, sizeof(*team));
    ucc_status_t st;
    if (!team) return UCC_ERR_NO_MEMORY;
    team->super.super.context = b_ctx; team->super.super.params = *params;
    team->tl_teams = (ucc_tl_team_t **)calloc(ctx->n_tl_ctxs, sizeof(void *));
    st = ucc_team_multiple_req_alloc(&team->team_create_req, (int)ctx->n_tl_ctxs);
    if (st != UCC_OK) { free(team->tl_teams); free(team); return st; }
    for (unsigned i = 0; i < ctx->n_tl_ctxs; i++) {
        ucc_team_multiple_req_descr_t *d = &team->team_create_req->descs[i];
        d->param = *params; d->param.scope = UCC_CL_BASIC; d->param.scope_id = 0; d->param.id = params->id;
        d->param.team = params->team; d->param.map = params->map;
        d->ctx = ctx->tl_ctxs[i];
    }
    st = ucc_tl_team_create_multiple(team->team_create_req);
    if (st < 0) { ucc_team_multiple_req_free(team->team_create_req); free(team->tl_teams); free(team); return st; }
    *team_p = &team->super.super;
    return UCC_OK;
}

static ucc_status_t basic_team_create_test(ucc_base_team_t *b)
{
    ucc_cl_basic_team_t *team = ucc_derived_of(b, ucc_cl_basic_team_t);
    ucc_cl_basic_context_t *ctx = ucc_derived_of(b->context, ucc_cl_basic_context_t);
    ucc_base_lib_t *lib = b->context->lib;
    ucc_coll_score_t *score = NULL, *tl_score;
    ucc_status_t st = ucc_tl_team_create_multiple(team->team_create_req);
    if (st == UCC_INPROGRESS) return st;
    if (st < 0) goto fail;
    team->n_tl_teams = 0;
    for (int i = 0; i < team->team_create_req->n_teams; i++) {
        ucc_team_multiple_req_descr_t *d = &team->team_create_req->descs[i];
        if (d->status == UCC_OK && d->team) { team->tl_teams[team->n_tl_teams++] = d->team; cl_debug(lib, "initialized tl %s team", UCC_TL_CTX_IFACE(d->ctx)->super.name); }
        else cl_debug(lib, "failed to create tl %s team: %s", UCC_TL_CTX_IFACE(d->ctx)->super.name, ucc_status_string(d->status));
    }
    ucc_team_multiple_req_free(team->team_create_req); team->team_create_req = NULL;
    if (team->n_tl_teams == 0) { cl_error(lib, "no tl teams were created"); st = UCC_ERR_NO_RESOURCE; goto fail_noreq; }
    for (unsigned i = 0; i < team->n_tl_teams; i++) {
        st = UCC_TL_TEAM_IFACE(team->tl_teams[i])->team.get_scores(&team->tl_teams[i]->super, &tl_score);
        if (st != UCC_OK) { cl_warn(lib, "failed to get tl %s scores", UCC_TL_TEAM_IFACE(team->tl_teams[i])->super.name); continue; }
        if (!score) score = tl_score; else if ((st = ucc_coll_score_merge_in(&score, tl_score)) != UCC_OK) goto fail_noreq;
    }
    if (!score) { st = UCC_ERR_NO_RESOURCE; goto fail_noreq; }
    if (ctx->tune && ctx->tune[0]) {
        ucc_memory_type_t mt[] = {UCC_MEMORY_TYPE_HOST, UCC_MEMORY_TYPE_CUDA, UCC_MEMORY_TYPE_CUDA_MANAGED, UCC_MEMORY_TYPE_ROCM, UCC_MEMORY_TYPE_ROCM_MANAGED};
        ucc_coll_score_team_info_t info = {UCC_CL_BASIC_DEFAULT_SCORE, b->params.size, UCC_COLL_TYPE_ALL, mt, 5, NULL, NULL};
        if (ucc_coll_score_update_from_str(ctx->tune, &info, b, score) != UCC_OK) cl_warn(lib, "ignoring invalid CL_BASIC_TUNE \"%s\"", ctx->tune);
    }
    team->score = score;
    st = ucc_coll_score_dup(score, &score);
    if (st != UCC_OK) goto fail_noreq;
    st = ucc_coll_score_build_map(score, &team->score_map);
    if (st != UCC_OK) { ucc_coll_score_free(score); goto fail_noreq; }
    if (b->params.rank == 0 && ucc_global_log_component.log_level >= UCC_LOG_LEVEL_INFO) {
        /* the core prints the CL-level map (one CL_BASIC range per collective); this is what is behind it: TL, score, algorithm */
        ucc_info("----- cl/basic selection (size %u, scope %d) -----", b->params.size, b->params.scope);
        ucc_coll_score_map_print_info(team->score_map, UCC_LOG_LEVEL_INFO);
    }
    return UCC_OK;
fail:
    ucc_team_multiple_req_free(team->team_create_req);
fail_noreq:
    for (unsigned i = 0; i < team->n_tl_teams; i++) UCC_TL_TEAM_IFACE(team->tl_teams[i])->team.destroy(&team->tl_teams[i]->super);
    if (team->score) ucc_coll_score_free(team->score);
    free(team->tl_teams); free(team);
    return st;
}

static ucc_status_t basic_team_destroy(ucc_base_team_t *b)
{
    ucc_cl_basic_team_t *team = ucc_derived_of(b, ucc_cl_basic_team_t);
    ucc_status_t st;
    if (!team->team_create_req) {
        st = ucc_team_multiple_req_alloc(&team->team_create_req, (int)team->n_tl_teams);
        if (st != UCC_OK) return st;
        for (unsigned i = 0; i < team->n_tl_teams; i++) team->team_create_req->descs[i].team = team->tl_teams[i];
    }
    st = ucc_tl_team_destroy_multiple(team->team_create_req);
    if (st == UCC_INPROGRESS) return st;
    ucc_team_multiple_req_free(team->team_create_req);
    if (team->score_map) ucc_coll_score_free_map(team->score_map);
    if (team->score) ucc_coll_score_free(team->score);
    free(team->tl_teams); free(team);
    return st;
}

static ucc_status_t basic_coll_init(ucc_base_coll_args_t *b, ucc_base_team_t *team, ucc_coll_task_t **task)
{ return ucc_coll_init(ucc_derived_of(team, ucc_cl_basic_team_t)->score_map, b, task); }

/* what core sees: same ranges, every score rewritten to the CL's own so CL-level
 * competition (hier vs basic) is independent of TL scores; init goes through our map */
static ucc_status_t basic_team_get_scores(ucc_base_team_t *b, ucc_coll_score_t **score_p)
{
    ucc_cl_basic_team_t *team = ucc_derived_of(b, ucc_cl_basic_team_t);
    ucc_coll_score_t *out, *src = team->score;
    ucc_msg_range_t *r;
    ucc_status_t st = ucc_coll_score_alloc(&out);
    if (st != UCC_OK) return st;
    for (int c = 0; c < UCC_COLL_TYPE_NUM; c++) for (int m = 0; m < UCC_MEMORY_TYPE_LAST; m++)
        ucc_list_for_each(r, &src->scores[c][m], super.list_elem) {
            st = ucc_coll_score_add_range(out, (ucc_coll_type_t)UCC_BIT(c), (ucc_memory_type_t)m, r->start, r->end, UCC_CL_BASIC_DEFAULT_SCORE, basic_coll_init, b);
            if (st != UCC_OK) { ucc_coll_score_free(out); return st; }
        }
    *score_p = out;
    return UCC_OK;
}

ucc_cl_iface_t ucc_cl_basic = {
    .super = {.name = "basic", .score = UCC_CL_BASIC_DEFAULT_SCORE},
    .type = UCC_CL_BASIC,
    .cl_lib_config = {"CL_BASIC lib", "CL_BASIC_", cl_basic_lib_config_table, sizeof(ucc_cl_lib_config_t), {NULL, NULL}},
    .cl_context_config = {"CL_BASIC context", "CL_BASIC_", cl_basic_context_config_table, sizeof(ucc_cl_context_config_t), {NULL, NULL}},
    .lib = {basic_lib_init, basic_lib_finalize, basic_lib_get_attr, NULL},
    .context = {basic_ctx_create, NULL, basic_ctx_destroy, basic_ctx_get_attr, NULL, NULL, NULL},
    .team = {basic_team_create_post, basic_team_create_test, basic_team_destroy, basic_team_get_scores},
    .coll = {basic_coll_init},
};
UCC_COMPONENT_REGISTER_STATIC(cl, ucc_cl_basic)
static void UCC_CTOR cl_basic_cfg_register(void) { ucc_config_table_register(&ucc_cl_basic.cl_lib_config); ucc_config_table_register(&ucc_cl_basic.cl_context_config); }
