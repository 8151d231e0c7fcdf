/* cl/hier: hierarchical collective layer for multi-node teams (reference cl/hier, 3.8 K LoC).
 * Sub-groups NODE / NODE_LEADERS / NET / FULL each get their own TL teams and score map; a
 * hierarchical algorithm is a schedule of collectives on those sub-groups:
 *   allreduce  rab         node reduce -> leaders allreduce -> node bcast
 *   allreduce  split_rail  node reduce_scatter -> per-rail (NET) allreduce -> node allgather
 *   bcast      2step       leaders bcast + node bcast
 *   reduce     2step       node reduce + leaders reduce
 *   barrier    node fanin -> leaders barrier -> node fanout
 * Like the reference (cl_hier_team.c:56-59) the CL refuses single-node teams; whatever it cannot do
 * returns NOT_SUPPORTED at init so the score-map fallback hands the collective to cl/basic. */
#include "components/cl/ucc_cl.h"
#include "components/tl/ucc_tl.h"
#include "components/topo/ucc_topo.h"
#include "core/ucc_context.h"
#include "core/ucc_team.h"
#include "core/ucc_global_opts.h"
#include "schedule/ucc_schedule_pipelined.h"
#include "components/mc/ucc_mc.h"
#include "utils/ucc_string.h"
#include <strings.h>

#define UCC_CL_HIER_DEFAULT_SCORE 50
typedef enum { HIER_SBGP_NODE, HIER_SBGP_NODE_LEADERS, HIER_SBGP_NET, HIER_SBGP_FULL, HIER_SBGP_LAST } hier_sbgp_t;
static const ucc_sbgp_type_t hier_to_topo[HIER_SBGP_LAST] = {UCC_SBGP_NODE, UCC_SBGP_NODE_LEADERS, UCC_SBGP_NET, UCC_SBGP_FULL};
static const char *hier_sbgp_names[HIER_SBGP_LAST] = {"node", "node_leaders", "net", "full"};

typedef struct ucc_cl_hier_lib_config { ucc_cl_lib_config_t super; ucc_config_allow_list_t sbgp_tls[HIER_SBGP_LAST]; size_t a2av_node_thresh;
    ucc_pipeline_params_t allreduce_rab_pipeline, allreduce_split_rail_pipeline, bcast_2step_pipeline, reduce_2step_pipeline; } ucc_cl_hier_lib_config_t;
typedef struct ucc_cl_hier_context_config { ucc_cl_context_config_t super; } ucc_cl_hier_context_config_t;
typedef struct ucc_cl_hier_lib { ucc_cl_lib_t super; ucc_config_names_list_t sbgp_tls[HIER_SBGP_LAST]; size_t a2av_node_thresh;
    ucc_pipeline_params_t allreduce_rab_pipeline, allreduce_split_rail_pipeline, bcast_2step_pipeline, reduce_2step_pipeline; } ucc_cl_hier_lib_t;
typedef struct ucc_cl_hier_context { ucc_cl_context_t super; ucc_tl_context_t **tl_ctxs; unsigned n_tl_ctxs; char *tune; } ucc_cl_hier_context_t;
typedef struct hier_sbgp {
    int enabled; ucc_sbgp_t *sbgp; ucc_tl_team_t **tl_teams; unsigned n_tl_teams; ucc_coll_score_t *score; ucc_score_map_t *map;
    ucc_team_multiple_req_t *req;
} hier_sbgp_state_t;
typedef struct ucc_cl_hier_team { ucc_cl_team_t super; hier_sbgp_state_t sb[HIER_SBGP_LAST]; int cur; ucc_coll_score_t *score; } ucc_cl_hier_team_t;
extern ucc_cl_iface_t ucc_cl_hier;
#define HLIB(_t) ((_t)->super.super.context->lib)

static ucc_config_field_t cl_hier_lib_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_cl_hier_lib_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_cl_lib_config_table)},
    {"NODE_SBGP_TLS", "all", "TLs used inside a node", ucc_offsetof(ucc_cl_hier_lib_config_t, sbgp_tls[HIER_SBGP_NODE]), UCC_CONFIG_TYPE_ALLOW_LIST},
    {"NODE_LEADERS_SBGP_TLS", "all", "TLs used between node leaders", ucc_offsetof(ucc_cl_hier_lib_config_t, sbgp_tls[HIER_SBGP_NODE_LEADERS]), UCC_CONFIG_TYPE_ALLOW_LIST},
    {"NET_SBGP_TLS", "all", "TLs used on a rail (ranks with the same local index on every node)", ucc_offsetof(ucc_cl_hier_lib_config_t, sbgp_tls[HIER_SBGP_NET]), UCC_CONFIG_TYPE_ALLOW_LIST},
    {"FULL_SBGP_TLS", "all", "TLs used on the full team", ucc_offsetof(ucc_cl_hier_lib_config_t, sbgp_tls[HIER_SBGP_FULL]), UCC_CONFIG_TYPE_ALLOW_LIST},
    {"ALLREDUCE_RAB_PIPELINE", "n", "Pipelining of the RAB allreduce: thresh=<size>:fragsize=<size>:nfrags=<n>:pdepth=<n>:<parallel|ordered|sequential>",
     ucc_offsetof(ucc_cl_hier_lib_config_t, allreduce_rab_pipeline), UCC_CONFIG_TYPE_PIPELINE_PARAMS},
    {"ALLREDUCE_SPLIT_RAIL_PIPELINE", "n", "Pipelining of the split_rail allreduce (fragments are multiples of the node size)",
     ucc_offsetof(ucc_cl_hier_lib_config_t, allreduce_split_rail_pipeline), UCC_CONFIG_TYPE_PIPELINE_PARAMS},
    {"BCAST_2STEP_PIPELINE", "n", "Pipelining of the 2step bcast", ucc_offsetof(ucc_cl_hier_lib_config_t, bcast_2step_pipeline), UCC_CONFIG_TYPE_PIPELINE_PARAMS},
    {"REDUCE_2STEP_PIPELINE", "n", "Pipelining of the 2step reduce", ucc_offsetof(ucc_cl_hier_lib_config_t, reduce_2step_pipeline), UCC_CONFIG_TYPE_PIPELINE_PARAMS},
    {"ALLTOALLV_SPLIT_NODE_THRESH", "0", "node_split alltoall(v): blocks for node peers larger than this many bytes go through the NODE sub-team, smaller ones stay in the full exchange",
     ucc_offsetof(ucc_cl_hier_lib_config_t, a2av_node_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {NULL}};
static ucc_config_field_t cl_hier_context_config_table[] = {{"", "", NULL, 0, UCC_CONFIG_TYPE_TABLE(ucc_cl_context_config_table)}, {NULL}};

/* ---- lib / ctx ---- */
static ucc_status_t hier_lib_init(const ucc_base_lib_params_t *p, const ucc_base_lib_config_t *config, ucc_base_lib_t **lib_p)
{
    const ucc_cl_hier_lib_config_t *cfg = ucc_derived_of(config, ucc_cl_hier_lib_config_t);
    ucc_cl_hier_lib_t *lib = (ucc_cl_hier_lib_t *)calloc(1, sizeof(*lib));
    (void)p;
    if (!lib) return UCC_ERR_NO_MEMORY;
    if (ucc_cl_lib_init_base(&lib->super, &ucc_cl_hier, &cfg->super) != UCC_OK) { free(lib); return UCC_ERR_NO_MEMORY; }
    for (int i = 0; i < HIER_SBGP_LAST; i++) ucc_config_allow_list_process(&cfg->sbgp_tls[i], &lib->super.tls.array, &lib->sbgp_tls[i]);
    lib->allreduce_rab_pipeline = cfg->allreduce_rab_pipeline; lib->allreduce_split_rail_pipeline = cfg->allreduce_split_rail_pipeline;
    lib->bcast_2step_pipeline = cfg->bcast_2step_pipeline; lib->reduce_2step_pipeline = cfg->reduce_2step_pipeline; lib->a2av_node_thresh = cfg->a2av_node_thresh;
    *lib_p = &lib->super.super;
    return UCC_OK;
}
static void hier_lib_finalize(ucc_base_lib_t *b)
{ ucc_cl_hier_lib_t *lib = ucc_derived_of(b, ucc_cl_hier_lib_t); for (int i = 0; i < HIER_SBGP_LAST; i++) ucc_config_names_array_free(&lib->sbgp_tls[i].array); ucc_cl_lib_cleanup_base(&lib->super); free(lib); }
static ucc_status_t hier_lib_get_attr(const ucc_base_lib_t *b, ucc_base_lib_attr_t *battr)
{
    const ucc_cl_hier_lib_t *lib = ucc_derived_of(b, ucc_cl_hier_lib_t);
    ucc_cl_lib_attr_t *attr = ucc_derived_of(battr, ucc_cl_lib_attr_t);
    attr->super.attr.thread_mode = UCC_THREAD_MULTIPLE;
    attr->super.attr.coll_types = UCC_COLL_TYPE_ALLREDUCE | UCC_COLL_TYPE_BCAST | UCC_COLL_TYPE_REDUCE | UCC_COLL_TYPE_BARRIER | UCC_COLL_TYPE_ALLTOALL | UCC_COLL_TYPE_ALLTOALLV | UCC_COLL_TYPE_ALLGATHERV;
    attr->super.flags = UCC_BASE_LIB_FLAG_SERVICE_TEAM_REQUIRED | UCC_BASE_LIB_FLAG_TEAM_ID_REQUIRED;
    attr->tls = (ucc_config_names_array_t *)&lib->super.tls.array; attr->tls_forced = lib->super.tls_forced;
    return UCC_OK;
}
static ucc_status_t hier_ctx_create(const ucc_base_context_params_t *p, const ucc_base_ctx_config_t *config, ucc_base_context_t **ctx_p)
{
    ucc_cl_lib_t *lib = ucc_derived_of(config->lib, ucc_cl_lib_t);
    ucc_cl_hier_context_t *ctx = (ucc_cl_hier_context_t *)calloc(1, sizeof(*ctx));
    if (!ctx) return UCC_ERR_NO_MEMORY;
    ctx->super.super.ucc_context = p->context; ctx->super.super.lib = config->lib; ctx->tune = config->score_str ? strdup(config->score_str) : NULL;
    ctx->tl_ctxs = (ucc_tl_context_t **)calloc(lib->tls.array.count + 1, sizeof(void *));
    for (unsigned i = 0; i < lib->tls.array.count; i++) if (ucc_tl_context_get(p->context, lib->tls.array.names[i], &ctx->tl_ctxs[ctx->n_tl_ctxs]) == UCC_OK) ctx->n_tl_ctxs++;
    if (!ctx->n_tl_ctxs) { free(ctx->tl_ctxs); free(ctx->tune); free(ctx); return UCC_ERR_NOT_FOUND; }
    *ctx_p = &ctx->super.super;
    return UCC_OK;
}
static void hier_ctx_destroy(ucc_base_context_t *b)
{ ucc_cl_hier_context_t *ctx = ucc_derived_of(b, ucc_cl_hier_context_t); for (unsigned i = 0; i < ctx->n_tl_ctxs; i++) ucc_tl_context_put(ctx->tl_ctxs[i]); free(ctx->tl_ctxs); free(ctx->tune); free(ctx); }
static ucc_status_t hier_ctx_get_attr(const ucc_base_context_t *b, ucc_base_ctx_attr_t *attr)
{ (void)b; if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN) attr->attr.ctx_addr_len = 0; attr->topo_required = 1; return UCC_OK; }

/* ---- team ---- */
static ucc_status_t sbgp_post(ucc_cl_hier_team_t *team, int i)
{
    ucc_cl_hier_context_t *ctx = ucc_derived_of(team->super.super.context, ucc_cl_hier_context_t);
    ucc_cl_hier_lib_t *lib = ucc_derived_of(HLIB(team), ucc_cl_hier_lib_t);
    hier_sbgp_state_t *s = &team->sb[i];
    unsigned n = 0;
    ucc_status_t st = ucc_team_multiple_req_alloc(&s->req, (int)ctx->n_tl_ctxs);
    if (st != UCC_OK) return st;
    for (unsigned k = 0; k < ctx->n_tl_ctxs; k++) {
        ucc_team_multiple_req_descr_t *d;
        if (ucc_config_names_search(&lib->sbgp_tls[i].array, UCC_TL_CTX_IFACE(ctx->tl_ctxs[k])->super.name) < 0) continue;
        d = &s->req->descs[n++];
        d->ctx = ctx->tl_ctxs[k]; d->param = team->super.super.params;
        d->param.params.mask &= ~(uint64_t)UCC_TEAM_PARAM_FIELD_OOB; /* sub-teams wire up over the internal (service) OOB */
        d->param.scope = UCC_CL_HIER; d->param.scope_id = i; d->param.rank = s->sbgp->group_rank; d->param.size = s->sbgp->group_size; d->param.map = s->sbgp->map;
    }
    s->req->n_teams = (int)n;
    st = ucc_tl_team_create_multiple(s->req);
    return st < 0 ? st : UCC_OK;
}
static ucc_status_t sbgp_finish(ucc_cl_hier_team_t *team, int i)
{
    hier_sbgp_state_t *s = &team->sb[i];
    ucc_coll_score_t *tl_score;
    s->tl_teams = (ucc_tl_team_t **)calloc((size_t)s->req->n_teams + 1, sizeof(void *));
    for (int k = 0; k < s->req->n_teams; k++) if (s->req->descs[k].status == UCC_OK && s->req->descs[k].team) s->tl_teams[s->n_tl_teams++] = s->req->descs[k].team;
    ucc_team_multiple_req_free(s->req); s->req = NULL;
    if (!s->n_tl_teams) { s->enabled = 0; return UCC_OK; }
    for (unsigned k = 0; k < s->n_tl_teams; k++) {
        if (UCC_TL_TEAM_IFACE(s->tl_teams[k])->team.get_scores(&s->tl_teams[k]->super, &tl_score) != UCC_OK) continue;
        if (!s->score) s->score = tl_score; else UCC_CHECK_RET(ucc_coll_score_merge_in(&s->score, tl_score));
    }
    if (!s->score) { s->enabled = 0; return UCC_OK; }
    return ucc_coll_score_build_map(s->score, &s->map);
}

static ucc_status_t hier_team_create_post(ucc_base_context_t *b_ctx, const ucc_base_team_params_t *params, ucc_base_team_t **team_p)
{
    ucc_cl_hier_team_t *team;
    ucc_topo_t *topo = params->team->topo;
    if (!topo) { cl_debug(b_ctx->lib, "no topology information: cl/hier is not usable"); return UCC_ERR_NOT_SUPPORTED; }
    if (ucc_topo_is_single_node(topo)) { cl_debug(b_ctx->lib, "single node team: cl/hier is not needed"); return UCC_ERR_NOT_SUPPORTED; }
    team = (ucc_cl_hier_team_t *)calloc(1, sizeof(*team));
    if (!team) return UCC_ERR_NO_MEMORY;
    team->super.super.context = b_ctx; team->super.super.params = *params;
    for (int i = 0; i < HIER_SBGP_LAST; i++) {
        ucc_sbgp_t *sb = ucc_topo_get_sbgp(topo, hier_to_topo[i]);
        team->sb[i].sbgp = sb;
        /* a rank that is not a member (e.g. not a leader) simply does not have that sub-team */
        team->sb[i].enabled = (sb->status == UCC_SBGP_ENABLED && sb->group_size > 1 && sb->group_rank != UCC_RANK_INVALID);
    }
    team->cur = -1;
    *team_p = &team->super.super;
    return UCC_OK;
}
static ucc_status_t hier_team_create_test(ucc_base_team_t *b)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b, ucc_cl_hier_team_t);
    ucc_status_t st;
    for (;;) {
        if (team->cur >= 0 && team->sb[team->cur].req) {
            st = ucc_tl_team_create_multiple(team->sb[team->cur].req);
            if (st == UCC_INPROGRESS) return st;
            if (st < 0) return st;
            UCC_CHECK_RET(sbgp_finish(team, team->cur));
        }
        do { team->cur++; } while (team->cur < HIER_SBGP_LAST && !team->sb[team->cur].enabled);
        if (team->cur >= HIER_SBGP_LAST) break;
        st = sbgp_post(team, team->cur);
        if (st < 0) return st;
    }
    if (!team->sb[HIER_SBGP_NODE].sbgp || team->sb[HIER_SBGP_NODE].sbgp->status != UCC_SBGP_ENABLED) return UCC_ERR_NOT_SUPPORTED;
    return UCC_OK;
}
static ucc_status_t hier_team_destroy(ucc_base_team_t *b)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b, ucc_cl_hier_team_t);
    for (int i = 0; i < HIER_SBGP_LAST; i++) {
        hier_sbgp_state_t *s = &team->sb[i];
        for (unsigned k = 0; k < s->n_tl_teams; k++) { ucc_status_t st; do { st = UCC_TL_TEAM_IFACE(s->tl_teams[k])->team.destroy(&s->tl_teams[k]->super); } while (st == UCC_INPROGRESS); }
        if (s->map) ucc_coll_score_free_map(s->map); else if (s->score) ucc_coll_score_free(s->score);
        free(s->tl_teams);
    }
    free(team);
    return UCC_OK;
}

/* ------------------------------------------------------------------ */
/* hierarchical algorithms = schedules of sub-collectives              */
/* ------------------------------------------------------------------ */
typedef struct hier_schedule { ucc_schedule_t super; ucc_mc_buffer_header_t *scratch; void *arrays; } hier_schedule_t;
static ucc_status_t hier_schedule_finalize(ucc_coll_task_t *t)
{ hier_schedule_t *hs = (hier_schedule_t *)t; ucc_status_t st = ucc_schedule_finalize(t); if (hs->scratch) ucc_mc_free(hs->scratch); free(hs->arrays); free(t); return st; }
static ucc_status_t hier_schedule_post(ucc_coll_task_t *t) { return ucc_schedule_start(t); }

static ucc_status_t sub_coll(ucc_cl_hier_team_t *team, int sb, ucc_base_coll_args_t *proto, ucc_coll_args_t *args, ucc_coll_task_t **task)
{
    ucc_base_coll_args_t b = *proto;
    b.args = *args; b.asymm.scratch = NULL;
    if (!team->sb[sb].enabled || !team->sb[sb].map) return UCC_ERR_NOT_SUPPORTED;
    /* sub-collectives select by their own (sub-team) size: give the score map a team view of that size */
    return ucc_coll_init_as(team->sb[sb].map, &b, team->sb[sb].sbgp->group_rank, team->sb[sb].sbgp->group_size, task);
}
/* chain: tasks run one after another */
static ucc_status_t chain(hier_schedule_t *hs, ucc_coll_task_t **tasks, int n)
{
    for (int i = 0; i < n; i++) {
        UCC_CHECK_RET(ucc_schedule_add_task(&hs->super, tasks[i]));
        if (i == 0) UCC_CHECK_RET(ucc_task_subscribe_dep(&hs->super.super, tasks[0], UCC_EVENT_SCHEDULE_STARTED));
        else UCC_CHECK_RET(ucc_task_subscribe_dep(tasks[i - 1], tasks[i], UCC_EVENT_COMPLETED));
    }
    return UCC_OK;
}
static ucc_status_t hier_sched_alloc(ucc_base_coll_args_t *b, ucc_base_team_t *team, hier_schedule_t **hp)
{
    hier_schedule_t *hs = (hier_schedule_t *)calloc(1, sizeof(*hs));
    if (!hs) return UCC_ERR_NO_MEMORY;
    ucc_schedule_init(&hs->super, b, team);
    hs->super.super.post = hier_schedule_post; hs->super.super.finalize = hier_schedule_finalize;
    *hp = hs;
    return UCC_OK;
}
static void free_tasks(ucc_coll_task_t **t, int n) { for (int i = 0; i < n; i++) if (t[i]) t[i]->finalize(t[i]); }
static int is_leader(ucc_cl_hier_team_t *team) { ucc_sbgp_t *l = team->sb[HIER_SBGP_NODE_LEADERS].sbgp; return l && l->status == UCC_SBGP_ENABLED && l->group_rank != UCC_RANK_INVALID; }
static int node_multi(ucc_cl_hier_team_t *team) { return team->sb[HIER_SBGP_NODE].enabled; }

/* is the (team-rank) root its node's leader?  rooted 2-step algorithms need that */
static int root_is_leader(ucc_cl_hier_team_t *team, ucc_rank_t root, ucc_rank_t *root_leader_rank)
{
    /* all ranks must be able to evaluate this: the leaders list is global knowledge (topology) */
    ucc_rank_t *nl = NULL;
    if (ucc_topo_get_node_leaders(team->super.super.params.team->topo, &nl) != UCC_OK) return 0;
    if (nl[root] != root) return 0;
    if (root_leader_rank) { /* rank of root inside the leaders group = number of leaders with a smaller team rank */
        ucc_rank_t r = 0;
        for (ucc_rank_t i = 0; i < root; i++) if (nl[i] == i) r++;
        *root_leader_rank = r;
    }
    return 1;
}

/* ---- chains of up to three sub-collectives (rab, split_rail, 2step bcast / reduce) ----
 * An algorithm is an "args function": from the collective's arguments it derives the arguments and the sub-group of every step.  The same
 * function builds the schedule and, for the pipelined variants, re-derives the steps of a fragment (offset buffers, its count) each time the
 * fragment is re-armed - so a fragment of any algorithm is just the algorithm on a shorter vector.
 * Returns the number of steps of THIS rank (0: nothing to do here), < 0: not supported.  *need: bytes of scratch the rank wants (asked
 * with scratch == NULL first). */
#define HIER_CHAIN_MAX 3
typedef int (*hier_args_fn_t)(ucc_cl_hier_team_t *team, const ucc_coll_args_t *a, void *scratch, size_t *need, ucc_coll_args_t *sub, int *sb);

/* allreduce rab: node reduce to the leader -> leaders allreduce -> node bcast (reference allreduce/allreduce_rab.c:79-270) */
static int rab_args(ucc_cl_hier_team_t *team, const ucc_coll_args_t *a, void *scratch, size_t *need, ucc_coll_args_t *sub, int *sb)
{
    int n = 0, leader = is_leader(team), inplace = UCC_IS_INPLACE(*a);
    (void)scratch; (void)need;
    if (a->op == UCC_OP_AVG) return -1; /* averaging over sub-groups needs the global size: left to cl/basic */
    if (node_multi(team)) { /* node reduce to the leader (local rank 0) */
        sub[n] = *a; sub[n].coll_type = UCC_COLL_TYPE_REDUCE; sub[n].root = 0; sub[n].mask |= UCC_COLL_ARGS_FIELD_FLAGS;
        if (inplace && !leader) { sub[n].flags &= ~(uint64_t)UCC_COLL_ARGS_FLAG_IN_PLACE; sub[n].src.info = a->dst.info; }
        sb[n++] = HIER_SBGP_NODE;
    }
    if (leader && team->sb[HIER_SBGP_NODE_LEADERS].enabled) {
        sub[n] = *a; sub[n].mask |= UCC_COLL_ARGS_FIELD_FLAGS;
        if (node_multi(team)) sub[n].flags |= UCC_COLL_ARGS_FLAG_IN_PLACE; /* result of the node reduce is already in dst */
        sb[n++] = HIER_SBGP_NODE_LEADERS;
    }
    if (node_multi(team)) {
        sub[n] = *a; sub[n].coll_type = UCC_COLL_TYPE_BCAST; sub[n].root = 0; sub[n].src.info = a->dst.info; sub[n].mask |= UCC_COLL_ARGS_FIELD_FLAGS;
        sub[n].flags &= ~(uint64_t)UCC_COLL_ARGS_FLAG_IN_PLACE;
        sb[n++] = HIER_SBGP_NODE;
    }
    return n;
}

/* allreduce split_rail: node reduce_scatter -> PPN concurrent rail allreduces -> node allgather (reference allreduce/allreduce_split_rail.c) */
static int split_rail_args(ucc_cl_hier_team_t *team, const ucc_coll_args_t *a, void *scratch, size_t *need, ucc_coll_args_t *sub, int *sb)
{
    ucc_topo_t *topo = team->super.super.params.team->topo;
    ucc_sbgp_t *node = team->sb[HIER_SBGP_NODE].sbgp;
    size_t count = a->dst.info.count, dts = ucc_dt_size(a->dst.info.datatype), ppn, off, cnt;
    (void)scratch; (void)need;
    if (a->op == UCC_OP_AVG || !ucc_topo_isoppn(topo) || ucc_topo_min_ppn(topo) < 2 || !node_multi(team) || !team->sb[HIER_SBGP_NET].enabled) return -1;
    ppn = node->group_size;
    if (count < ppn || count % ppn) return -1; /* equal blocks keep reduce_scatter / allgather plain */
    cnt = count / ppn; off = cnt * node->group_rank;
    /* 1) node reduce_scatter: my block lands at its final place in dst */
    sub[0] = *a; sub[0].coll_type = UCC_COLL_TYPE_REDUCE_SCATTER; sub[0].mask |= UCC_COLL_ARGS_FIELD_FLAGS;
    if (!UCC_IS_INPLACE(*a)) { sub[0].flags &= ~(uint64_t)UCC_COLL_ARGS_FLAG_IN_PLACE; sub[0].dst.info.buffer = PTR_OFFSET(a->dst.info.buffer, off * dts); sub[0].dst.info.count = cnt; }
    sb[0] = HIER_SBGP_NODE;
    /* 2) rail allreduce of my block */
    sub[1] = *a; sub[1].mask |= UCC_COLL_ARGS_FIELD_FLAGS; sub[1].flags |= UCC_COLL_ARGS_FLAG_IN_PLACE;
    sub[1].dst.info.buffer = PTR_OFFSET(a->dst.info.buffer, off * dts); sub[1].dst.info.count = cnt; sub[1].src.info = sub[1].dst.info;
    sb[1] = HIER_SBGP_NET;
    /* 3) node allgather in place */
    sub[2] = *a; sub[2].coll_type = UCC_COLL_TYPE_ALLGATHER; sub[2].mask |= UCC_COLL_ARGS_FIELD_FLAGS; sub[2].flags |= UCC_COLL_ARGS_FLAG_IN_PLACE;
    sub[2].dst.info = a->dst.info; sub[2].src.info = a->dst.info; sub[2].src.info.count = cnt;
    sb[2] = HIER_SBGP_NODE;
    return 3;
}

/* bcast 2step: leaders bcast -> node bcast (reference bcast/bcast_2step.c); the root has to be its node's leader */
static int bcast_2step_args(ucc_cl_hier_team_t *team, const ucc_coll_args_t *a, void *scratch, size_t *need, ucc_coll_args_t *sub, int *sb)
{
    ucc_rank_t lroot = 0; int n = 0;
    (void)scratch; (void)need;
    if (UCC_COLL_ARGS_ACTIVE_SET(a) || !root_is_leader(team, (ucc_rank_t)a->root, &lroot)) return -1;
    if (is_leader(team) && team->sb[HIER_SBGP_NODE_LEADERS].enabled) { sub[n] = *a; sub[n].root = lroot; sb[n++] = HIER_SBGP_NODE_LEADERS; }
    if (node_multi(team)) { sub[n] = *a; sub[n].root = 0; sb[n++] = HIER_SBGP_NODE; }
    return n;
}

/* reduce 2step: node reduce -> leaders reduce (reference reduce/reduce_2step.c); leaders other than the root keep their node's partial result
 * in a scratch buffer */
static int reduce_2step_args(ucc_cl_hier_team_t *team, const ucc_coll_args_t *a, void *scratch, size_t *need, ucc_coll_args_t *sub, int *sb)
{
    ucc_rank_t lroot = 0, me = team->super.super.params.rank; int n = 0, root = (ucc_rank_t)a->root == me, leader = is_leader(team);
    if (a->op == UCC_OP_AVG || !root_is_leader(team, (ucc_rank_t)a->root, &lroot)) return -1;
    if (leader && !root && node_multi(team)) { size_t len = a->src.info.count * ucc_dt_size(a->src.info.datatype); *need = len ? len : 1; }
    if (node_multi(team)) {
        sub[n] = *a; sub[n].root = 0; sub[n].mask |= UCC_COLL_ARGS_FIELD_FLAGS;
        if (leader && !root) { sub[n].dst.info = a->src.info; sub[n].dst.info.buffer = scratch; sub[n].flags &= ~(uint64_t)UCC_COLL_ARGS_FLAG_IN_PLACE; }
        sb[n++] = HIER_SBGP_NODE;
    }
    if (leader && team->sb[HIER_SBGP_NODE_LEADERS].enabled) {
        sub[n] = *a; sub[n].root = lroot; sub[n].mask |= UCC_COLL_ARGS_FIELD_FLAGS;
        if (node_multi(team)) { if (root) sub[n].flags |= UCC_COLL_ARGS_FLAG_IN_PLACE; else sub[n].src.info.buffer = scratch; }
        sb[n++] = HIER_SBGP_NODE_LEADERS;
    }
    return n;
}

static ucc_status_t hier_chain_build(hier_args_fn_t fn, ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b_team, ucc_cl_hier_team_t);
    ucc_coll_args_t sub[HIER_CHAIN_MAX]; int sb[HIER_CHAIN_MAX], n;
    ucc_coll_task_t *tasks[HIER_CHAIN_MAX] = {NULL, NULL, NULL}; hier_schedule_t *hs; ucc_status_t st;
    ucc_mc_buffer_header_t *scratch = NULL; size_t need = 0;
    n = fn(team, &b->args, NULL, &need, sub, sb);
    if (n <= 0) return UCC_ERR_NOT_SUPPORTED;
    if (need) {
        if (ucc_mc_alloc(&scratch, need, b->args.src.info.mem_type) != UCC_OK) return UCC_ERR_NO_MEMORY;
        n = fn(team, &b->args, scratch->addr, &need, sub, sb);
    }
    for (int i = 0; i < n; i++) { st = sub_coll(team, sb[i], b, &sub[i], &tasks[i]); if (st != UCC_OK) goto err; }
    st = hier_sched_alloc(b, b_team, &hs); if (st != UCC_OK) goto err;
    st = chain(hs, tasks, n); if (st != UCC_OK) { free(hs); goto err; }
    hs->scratch = scratch; /* released with the schedule */
    *task_p = &hs->super.super;
    return UCC_OK;
err:
    if (scratch) ucc_mc_free(scratch);
    free_tasks(tasks, HIER_CHAIN_MAX);
    return st == UCC_ERR_NOT_IMPLEMENTED ? UCC_ERR_NOT_SUPPORTED : st;
}

/* pipelined chains: the vector is cut into fragments of `fcount` elements (the last one shorter), up to pdepth chains are in flight and
 * re-armed round robin on top of the pipelined schedule (reference: allreduce_rab.c:79-270, allreduce_split_rail.c, bcast_2step.c,
 * reduce_2step.c each carry their own frag_init / frag_setup pair; here one pair serves every args function) */
typedef struct hier_pipe { ucc_schedule_pipelined_t super; hier_args_fn_t fn; size_t total, dts, fcount; } hier_pipe_t;
/* the collective restricted to elements [off, off + cnt) */
static void hier_frag_args(const ucc_coll_args_t *a, size_t off_bytes, size_t cnt, ucc_coll_args_t *f)
{
    *f = *a;
    if (f->src.info.buffer) f->src.info.buffer = PTR_OFFSET(f->src.info.buffer, off_bytes);
    if (f->dst.info.buffer) f->dst.info.buffer = PTR_OFFSET(f->dst.info.buffer, off_bytes);
    f->src.info.count = cnt; f->dst.info.count = cnt;
}
static ucc_status_t hier_pipe_frag_init(ucc_base_coll_args_t *b, ucc_schedule_pipelined_t *sp, ucc_base_team_t *team, ucc_schedule_t **frag)
{
    hier_pipe_t *hp = (hier_pipe_t *)sp;
    ucc_base_coll_args_t fb = *b; ucc_coll_task_t *t; ucc_status_t st;
    hier_frag_args(&b->args, 0, hp->fcount, &fb.args); /* the largest fragment: sizes the scratch and picks the algorithms of the steps */
    st = hier_chain_build(hp->fn, &fb, team, &t); if (st != UCC_OK) return st;
    *frag = (ucc_schedule_t *)t;
    return UCC_OK;
}
static ucc_status_t hier_pipe_frag_setup(ucc_schedule_pipelined_t *sp, ucc_schedule_t *frag, int frag_num)
{
    hier_pipe_t *hp = (hier_pipe_t *)sp;
    hier_schedule_t *hs = (hier_schedule_t *)frag;
    ucc_cl_hier_team_t *team = ucc_derived_of(frag->super.team, ucc_cl_hier_team_t);
    size_t off = (size_t)frag_num * hp->fcount, cnt = ucc_min(hp->fcount, hp->total - off), need = 0;
    ucc_coll_args_t fa, sub[HIER_CHAIN_MAX]; int sb[HIER_CHAIN_MAX], n;
    hier_frag_args(&sp->super.super.bargs.args, off * hp->dts, cnt, &fa);
    n = hp->fn(team, &fa, hs->scratch ? hs->scratch->addr : NULL, &need, sub, sb);
    if (n < 0 || (unsigned)n != frag->n_tasks) return UCC_ERR_NOT_SUPPORTED; /* the steps of a rank do not depend on the fragment */
    for (int j = 0; j < n; j++) {
        ucc_coll_args_t *ta = &frag->tasks[j]->bargs.args;
        ta->src = sub[j].src; ta->dst = sub[j].dst;
        frag->tasks[j]->flags |= UCC_COLL_TASK_FLAG_ARGS_UPDATED; /* TLs rebuild what they derived from the arguments at init */
    }
    return UCC_OK;
}
static ucc_status_t hier_pipe_finalize(ucc_coll_task_t *t) { ucc_status_t st = ucc_schedule_pipelined_finalize(t); free(t); return st; }

/* `count` elements of `dts` bytes; fragments are multiples of `align` elements (split_rail: the node size) */
static ucc_status_t hier_chain_init(hier_args_fn_t fn, const ucc_pipeline_params_t *pp, size_t count, size_t dts, size_t align, int plain_dt,
                                    ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    hier_pipe_t *hp; size_t fcount; int n_total, depth; ucc_status_t st;
    if (!pp->n_frags || !pp->pdepth || pp->threshold == UCC_MEMUNITS_INF || pp->threshold == UCC_MEMUNITS_AUTO || count * dts < pp->threshold || !plain_dt || !dts)
        return hier_chain_build(fn, b, b_team, task_p);
    fcount = (pp->frag_size == UCC_MEMUNITS_INF || pp->frag_size == UCC_MEMUNITS_AUTO) ? ucc_div_round_up(count, pp->n_frags) : ucc_max(1, pp->frag_size / dts);
    fcount = ucc_div_round_up(fcount, align) * align;
    n_total = (int)ucc_div_round_up(count, fcount);
    if (n_total < 2) return hier_chain_build(fn, b, b_team, task_p);
    depth = (int)ucc_min(ucc_min(pp->pdepth, (unsigned)n_total), UCC_SCHEDULE_PIPELINED_MAX_FRAGS);
    hp = (hier_pipe_t *)calloc(1, sizeof(*hp));
    if (!hp) return UCC_ERR_NO_MEMORY;
    hp->fn = fn; hp->total = count; hp->dts = dts; hp->fcount = fcount;
    st = ucc_schedule_pipelined_init(b, b_team, hier_pipe_frag_init, hier_pipe_frag_setup, depth, n_total, pp->order, &hp->super);
    if (st != UCC_OK) { free(hp); return st == UCC_ERR_NOT_IMPLEMENTED ? UCC_ERR_NOT_SUPPORTED : st; }
    hp->super.super.super.finalize = hier_pipe_finalize;
    *task_p = &hp->super.super.super;
    return UCC_OK;
}
#define HIER_LIB_OF(_b_team) ucc_derived_of(HLIB(ucc_derived_of(_b_team, ucc_cl_hier_team_t)), ucc_cl_hier_lib_t)

static ucc_status_t hier_allreduce_rab(ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    return hier_chain_init(rab_args, &HIER_LIB_OF(b_team)->allreduce_rab_pipeline, b->args.dst.info.count, ucc_dt_size(b->args.dst.info.datatype), 1,
                           !UCC_DT_IS_GENERIC(b->args.dst.info.datatype), b, b_team, task_p);
}
static ucc_status_t hier_allreduce_split_rail(ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b_team, ucc_cl_hier_team_t);
    size_t ppn = node_multi(team) ? team->sb[HIER_SBGP_NODE].sbgp->group_size : 1;
    return hier_chain_init(split_rail_args, &HIER_LIB_OF(b_team)->allreduce_split_rail_pipeline, b->args.dst.info.count, ucc_dt_size(b->args.dst.info.datatype), ppn,
                           !UCC_DT_IS_GENERIC(b->args.dst.info.datatype), b, b_team, task_p);
}
static ucc_status_t hier_bcast_2step(ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    return hier_chain_init(bcast_2step_args, &HIER_LIB_OF(b_team)->bcast_2step_pipeline, b->args.src.info.count, ucc_dt_size(b->args.src.info.datatype), 1,
                           !UCC_DT_IS_GENERIC(b->args.src.info.datatype), b, b_team, task_p);
}
static ucc_status_t hier_reduce_2step(ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b_team, ucc_cl_hier_team_t);
    /* a root working in place describes the vector by dst only; everyone else by src */
    const ucc_coll_buffer_info_t *i = ((ucc_rank_t)b->args.root == team->super.super.params.rank && UCC_IS_INPLACE(b->args)) ? &b->args.dst.info : &b->args.src.info;
    return hier_chain_init(reduce_2step_args, &HIER_LIB_OF(b_team)->reduce_2step_pipeline, i->count, ucc_dt_size(i->datatype), 1, !UCC_DT_IS_GENERIC(i->datatype), b, b_team, task_p);
}

static ucc_status_t hier_barrier(ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b_team, ucc_cl_hier_team_t);
    ucc_coll_args_t *a = &b->args, sub;
    ucc_coll_task_t *tasks[3] = {NULL, NULL, NULL}; int n = 0; hier_schedule_t *hs; ucc_status_t st;
    if (node_multi(team)) { sub = *a; sub.coll_type = UCC_COLL_TYPE_FANIN; sub.root = 0; st = sub_coll(team, HIER_SBGP_NODE, b, &sub, &tasks[n]); if (st != UCC_OK) goto err; n++; }
    if (is_leader(team) && team->sb[HIER_SBGP_NODE_LEADERS].enabled) { sub = *a; st = sub_coll(team, HIER_SBGP_NODE_LEADERS, b, &sub, &tasks[n]); if (st != UCC_OK) goto err; n++; }
    if (node_multi(team)) { sub = *a; sub.coll_type = UCC_COLL_TYPE_FANOUT; sub.root = 0; st = sub_coll(team, HIER_SBGP_NODE, b, &sub, &tasks[n]); if (st != UCC_OK) goto err; n++; }
    if (!n) return UCC_ERR_NOT_SUPPORTED;
    st = hier_sched_alloc(b, b_team, &hs); if (st != UCC_OK) goto err;
    st = chain(hs, tasks, n); if (st != UCC_OK) { free(hs); goto err; }
    *task_p = &hs->super.super;
    return UCC_OK;
err:
    free_tasks(tasks, 3);
    return st == UCC_ERR_NOT_IMPLEMENTED ? UCC_ERR_NOT_SUPPORTED : st;
}


/* ---- alltoall / alltoallv node_split: the exchange is cut into the intra-node part (NODE sub-team) and the rest
 * (FULL sub-team with empty blocks for node peers); both halves run concurrently and write disjoint ranges ---- */
static ucc_status_t hier_a2av_node_split(ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b_team, ucc_cl_hier_team_t);
    ucc_coll_args_t *a = &b->args, sub;
    ucc_coll_task_t *tasks[2] = {NULL, NULL}; hier_schedule_t *hs = NULL; ucc_status_t st;
    ucc_sbgp_t *node = team->sb[HIER_SBGP_NODE].sbgp, *full = team->sb[HIER_SBGP_FULL].sbgp;
    ucc_rank_t N = team->super.super.params.size, nn, i; int is_v = a->coll_type == UCC_COLL_TYPE_ALLTOALLV;
    uint64_t *arr, *nsc, *nsd, *nrc, *nrd, *fsc, *fsd, *frc, *frd;
    if (UCC_IS_INPLACE(*a) || !team->sb[HIER_SBGP_FULL].enabled) return UCC_ERR_NOT_SUPPORTED; /* both conditions are the same on every rank */
    size_t thresh = HIER_LIB_OF(b_team)->a2av_node_thresh, sdts = ucc_dt_size(is_v ? a->src.info_v.datatype : a->src.info.datatype), rdts = ucc_dt_size(is_v ? a->dst.info_v.datatype : a->dst.info.datatype);
    if (thresh == UCC_MEMUNITS_AUTO) thresh = 0;
    nn = node_multi(team) ? node->group_size : 0; /* alone on my node: everything goes through the FULL exchange */
    arr = (uint64_t *)calloc(4 * (size_t)nn + 4 * (size_t)N, sizeof(uint64_t));
    if (!arr) return UCC_ERR_NO_MEMORY;
    nsc = arr; nsd = nsc + nn; nrc = nsd + nn; nrd = nrc + nn; fsc = nrd + nn; fsd = fsc + N; frc = fsd + N; frd = frc + N;
    for (i = 0; i < N; i++) { /* FULL sub-team rank i == team rank rank_map[i] (identity for the full group) */
        ucc_rank_t r = full->rank_map ? full->rank_map[i] : i;
        if (is_v) { fsc[i] = ucc_coll_args_get_count(a, a->src.info_v.counts, r); fsd[i] = ucc_coll_args_get_displacement(a, a->src.info_v.displacements, r);
                    frc[i] = ucc_coll_args_get_count(a, a->dst.info_v.counts, r); frd[i] = ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, r); }
        else { fsc[i] = a->src.info.count / N; fsd[i] = (uint64_t)r * fsc[i]; frc[i] = a->dst.info.count / N; frd[i] = (uint64_t)r * frc[i]; }
    }
    for (i = 0; i < nn; i++) {
        ucc_rank_t r = node->rank_map[i], fi = r; /* node member's index inside FULL */
        if (full->rank_map) for (fi = 0; fi < N && full->rank_map[fi] != r; fi++) ;
        nsd[i] = fsd[fi]; nrd[i] = frd[fi];
        /* ALLTOALLV_SPLIT_NODE_THRESH: only blocks above the threshold leave the full exchange (sender and receiver judge the same block;
         * reference alltoallv/alltoallv.c SET_FULL_COUNTS / SET_NODE_COUNTS) */
        if (fsc[fi] * sdts > thresh) { nsc[i] = fsc[fi]; fsc[fi] = 0; }
        if (frc[fi] * rdts > thresh) { nrc[i] = frc[fi]; frc[fi] = 0; }
    }
    sub = *a; sub.coll_type = UCC_COLL_TYPE_ALLTOALLV; sub.mask |= UCC_COLL_ARGS_FIELD_FLAGS; sub.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT;
    if (!is_v) { sub.src.info_v.buffer = a->src.info.buffer; sub.src.info_v.datatype = a->src.info.datatype; sub.src.info_v.mem_type = a->src.info.mem_type;
                 sub.dst.info_v.buffer = a->dst.info.buffer; sub.dst.info_v.datatype = a->dst.info.datatype; sub.dst.info_v.mem_type = a->dst.info.mem_type; }
    sub.src.info_v.counts = (ucc_count_t *)nsc; sub.src.info_v.displacements = (ucc_aint_t *)nsd; sub.dst.info_v.counts = (ucc_count_t *)nrc; sub.dst.info_v.displacements = (ucc_aint_t *)nrd;
    if (nn) { st = sub_coll(team, HIER_SBGP_NODE, b, &sub, &tasks[0]); if (st != UCC_OK) goto err; }
    sub.src.info_v.counts = (ucc_count_t *)fsc; sub.src.info_v.displacements = (ucc_aint_t *)fsd; sub.dst.info_v.counts = (ucc_count_t *)frc; sub.dst.info_v.displacements = (ucc_aint_t *)frd;
    st = sub_coll(team, HIER_SBGP_FULL, b, &sub, &tasks[1]); if (st != UCC_OK) goto err;
    st = hier_sched_alloc(b, b_team, &hs); if (st != UCC_OK) goto err;
    for (i = 0; i < 2; i++) { /* both start with the schedule */
        if (!tasks[i]) continue;
        st = ucc_schedule_add_task(&hs->super, tasks[i]); if (st != UCC_OK) goto err;
        st = ucc_task_subscribe_dep(&hs->super.super, tasks[i], UCC_EVENT_SCHEDULE_STARTED); if (st != UCC_OK) goto err;
    }
    hs->arrays = arr;
    *task_p = &hs->super.super;
    return UCC_OK;
err:
    free(hs); free(arr); free_tasks(tasks, 2);
    return st == UCC_ERR_NOT_IMPLEMENTED ? UCC_ERR_NOT_SUPPORTED : st;
}

/* ---- allgatherv gab: node gatherv -> leaders allgatherv -> node bcast.  Needs the packed layout
 * (displs are the running sum of counts) and nodes made of consecutive ranks; other layouts go to cl/basic ---- */
static ucc_status_t hier_allgatherv_gab(ucc_base_coll_args_t *b, ucc_base_team_t *b_team, ucc_coll_task_t **task_p)
{
    ucc_cl_hier_team_t *team = ucc_derived_of(b_team, ucc_cl_hier_team_t);
    ucc_coll_args_t *a = &b->args, sub;
    ucc_coll_task_t *tasks[3] = {NULL, NULL, NULL}; int n = 0; hier_schedule_t *hs = NULL; ucc_status_t st;
    ucc_sbgp_t *node = team->sb[HIER_SBGP_NODE].sbgp, *lead = ucc_topo_get_sbgp(team->super.super.params.team->topo, UCC_SBGP_NODE_LEADERS);
    ucc_rank_t N = team->super.super.params.size, me = team->super.super.params.rank, i, nl = 0, *leaders = NULL;
    size_t dts = ucc_dt_size(a->dst.info_v.datatype), total = 0; uint64_t *arr, *gc, *gd, *lc, *ld; int leader = is_leader(team), inplace = UCC_IS_INPLACE(*a);
    if (ucc_topo_get_node_leaders(team->super.super.params.team->topo, &leaders) != UCC_OK) return UCC_ERR_NOT_SUPPORTED;
    for (i = 0; i < N; i++) {
        if (ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, i) != total) return UCC_ERR_NOT_SUPPORTED;
        total += ucc_coll_args_get_count(a, a->dst.info_v.counts, i);
        if (leaders[i] == i) nl++; else if (i == 0 || leaders[i] != leaders[i - 1]) return UCC_ERR_NOT_SUPPORTED; /* node = consecutive ranks led by the first */
    }
    (void)lead;
    arr = (uint64_t *)calloc(2 * (size_t)N + 2 * (size_t)nl + 2, sizeof(uint64_t));
    if (!arr) return UCC_ERR_NO_MEMORY;
    gc = arr; gd = gc + N; lc = gd + N; ld = lc + nl;
    if (node_multi(team)) {
        for (i = 0; i < node->group_size; i++) { gc[i] = ucc_coll_args_get_count(a, a->dst.info_v.counts, node->rank_map[i]); gd[i] = ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, node->rank_map[i]); }
        sub = *a; sub.coll_type = UCC_COLL_TYPE_GATHERV; sub.root = 0; sub.mask |= UCC_COLL_ARGS_FIELD_FLAGS; sub.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT;
        sub.dst.info_v.counts = (ucc_count_t *)gc; sub.dst.info_v.displacements = (ucc_aint_t *)gd;
        if (inplace && !leader) { sub.flags &= ~(uint64_t)UCC_COLL_ARGS_FLAG_IN_PLACE; sub.src.info.buffer = PTR_OFFSET(a->dst.info_v.buffer, ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, me) * dts);
            sub.src.info.count = ucc_coll_args_get_count(a, a->dst.info_v.counts, me); sub.src.info.datatype = a->dst.info_v.datatype; sub.src.info.mem_type = a->dst.info_v.mem_type; }
        st = sub_coll(team, HIER_SBGP_NODE, b, &sub, &tasks[n]); if (st != UCC_OK) goto err; n++;
    }
    if (leader && team->sb[HIER_SBGP_NODE_LEADERS].enabled) {
        ucc_rank_t j = 0;
        for (i = 0; i < N; i++) { if (leaders[i] == i) { ld[j] = ucc_coll_args_get_displacement(a, a->dst.info_v.displacements, i); j++; } lc[j - 1] += ucc_coll_args_get_count(a, a->dst.info_v.counts, i); }
        sub = *a; sub.mask |= UCC_COLL_ARGS_FIELD_FLAGS; sub.flags |= UCC_COLL_ARGS_FLAG_COUNT_64BIT | UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT;
        sub.dst.info_v.counts = (ucc_count_t *)lc; sub.dst.info_v.displacements = (ucc_aint_t *)ld;
        if (node_multi(team) || inplace) sub.flags |= UCC_COLL_ARGS_FLAG_IN_PLACE; /* the node's data already sits in dst */
        st = sub_coll(team, HIER_SBGP_NODE_LEADERS, b, &sub, &tasks[n]); if (st != UCC_OK) goto err; n++;
    }
    if (node_multi(team)) {
        sub = *a; sub.coll_type = UCC_COLL_TYPE_BCAST; sub.root = 0; sub.mask |= UCC_COLL_ARGS_FIELD_FLAGS; sub.flags &= ~(uint64_t)UCC_COLL_ARGS_FLAG_IN_PLACE;
        sub.src.info.buffer = a->dst.info_v.buffer; sub.src.info.count = total; sub.src.info.datatype = a->dst.info_v.datatype; sub.src.info.mem_type = a->dst.info_v.mem_type;
        st = sub_coll(team, HIER_SBGP_NODE, b, &sub, &tasks[n]); if (st != UCC_OK) goto err; n++;
    }
    if (!n) { st = UCC_ERR_NOT_SUPPORTED; goto err; }
    st = hier_sched_alloc(b, b_team, &hs); if (st != UCC_OK) goto err;
    st = chain(hs, tasks, n); if (st != UCC_OK) goto err;
    hs->arrays = arr;
    *task_p = &hs->super.super;
    return UCC_OK;
err:
    free(hs); free(arr); free_tasks(tasks, 3);
    return st == UCC_ERR_NOT_IMPLEMENTED ? UCC_ERR_NOT_SUPPORTED : st;
}

/* ---- algorithm table / scores ---- */
typedef struct hier_alg { const char *name, *desc; ucc_base_coll_init_fn_t init; } hier_alg_t;
static const hier_alg_t hier_allreduce_algs[] = {{"rab", "intra-node reduce, followed by inter-node allreduce, followed by innode broadcast", hier_allreduce_rab},
    {"split_rail", "intra-node reduce_scatter, followed by PPN concurrent inter-node allreduces, followed by intra-node allgather", hier_allreduce_split_rail}, {NULL, NULL, NULL}};
static const hier_alg_t hier_bcast_algs[] = {{"2step", "leaders bcast followed by intra-node bcast", hier_bcast_2step}, {NULL, NULL, NULL}};
static const hier_alg_t hier_reduce_algs[] = {{"2step", "intra-node reduce followed by leaders reduce", hier_reduce_2step}, {NULL, NULL, NULL}};
static const hier_alg_t hier_barrier_algs[] = {{"knomial", "node fanin, leaders barrier, node fanout", hier_barrier}, {NULL, NULL, NULL}};
static const hier_alg_t hier_a2a_algs[] = {{"node_split", "splitting alltoall into two concurrent a2av calls withing the node and outside of it", hier_a2av_node_split}, {NULL, NULL, NULL}};
static const hier_alg_t hier_a2av_algs[] = {{"node_split", "splitting alltoallv into two concurrent a2av calls withing the node and outside of it", hier_a2av_node_split}, {NULL, NULL, NULL}};
static const hier_alg_t hier_agv_algs[] = {{"gab", "gatherv + allgatherv + bcast", hier_allgatherv_gab}, {NULL, NULL, NULL}};
static const hier_alg_t *hier_algs(ucc_coll_type_t ct)
{ switch (ct) { case UCC_COLL_TYPE_ALLTOALL: return hier_a2a_algs; case UCC_COLL_TYPE_ALLTOALLV: return hier_a2av_algs; case UCC_COLL_TYPE_ALLGATHERV: return hier_agv_algs; case UCC_COLL_TYPE_ALLREDUCE: return hier_allreduce_algs; case UCC_COLL_TYPE_BCAST: return hier_bcast_algs; case UCC_COLL_TYPE_REDUCE: return hier_reduce_algs; case UCC_COLL_TYPE_BARRIER: return hier_barrier_algs; default: return NULL; } }
static ucc_base_coll_alg_info_t hier_alg_info[UCC_COLL_TYPE_NUM][3];

static ucc_status_t hier_alg_id_to_init(int alg_id, const char *s, ucc_coll_type_t ct, ucc_memory_type_t mt, ucc_base_coll_init_fn_t *init)
{
    const hier_alg_t *algs = hier_algs(ct); int n = 0; (void)mt;
    if (!algs) return UCC_ERR_NOT_SUPPORTED;
    while (algs[n].name) n++;
    if (s) { alg_id = -1; for (int i = 0; i < n; i++) if (!strcasecmp(s, algs[i].name)) alg_id = i; }
    if (alg_id < 0 || alg_id >= n) return s ? UCC_ERR_NOT_SUPPORTED : UCC_ERR_INVALID_PARAM;
    *init = algs[alg_id].init;
    return UCC_OK;
}
static ucc_status_t hier_coll_init(ucc_base_coll_args_t *b, ucc_base_team_t *team, ucc_coll_task_t **task)
{ const hier_alg_t *algs = hier_algs(b->args.coll_type); return algs ? algs[0].init(b, team, task) : UCC_ERR_NOT_SUPPORTED; }

static ucc_status_t hier_team_get_scores(ucc_base_team_t *b_team, ucc_coll_score_t **score_p)
{
    ucc_cl_hier_context_t *ctx = ucc_derived_of(b_team->context, ucc_cl_hier_context_t);
    ucc_memory_type_t mt[3] = {UCC_MEMORY_TYPE_HOST, UCC_MEMORY_TYPE_CUDA, UCC_MEMORY_TYPE_CUDA_MANAGED};
    uint64_t colls = UCC_COLL_TYPE_ALLREDUCE | UCC_COLL_TYPE_BCAST | UCC_COLL_TYPE_REDUCE | UCC_COLL_TYPE_BARRIER | UCC_COLL_TYPE_ALLTOALL | UCC_COLL_TYPE_ALLTOALLV | UCC_COLL_TYPE_ALLGATHERV;
    ucc_coll_score_team_info_t info = {UCC_CL_HIER_DEFAULT_SCORE, b_team->params.size, colls, mt, 3, hier_coll_init, hier_alg_id_to_init};
    ucc_coll_score_t *score;
    ucc_status_t st = ucc_coll_score_build_default(b_team, UCC_CL_HIER_DEFAULT_SCORE, hier_coll_init, colls, mt, 3, &score);
    if (st != UCC_OK) return st;
    /* reference defaults: rab for small/medium, split_rail for large vectors */
    st = ucc_coll_score_update_from_str("allreduce:0-4k:@rab#allreduce:4k-inf:@split_rail", &info, b_team, score);
    if (st == UCC_OK && ctx->tune && ctx->tune[0] && ucc_coll_score_update_from_str(ctx->tune, &info, b_team, score) != UCC_OK)
        cl_warn(b_team->context->lib, "ignoring invalid CL_HIER_TUNE \"%s\"", ctx->tune);
    if (st != UCC_OK) { ucc_coll_score_free(score); return st; }
    *score_p = score;
    return UCC_OK;
}

ucc_cl_iface_t ucc_cl_hier = {
    .super = {.name = "hier", .score = UCC_CL_HIER_DEFAULT_SCORE},
    .type = UCC_CL_HIER,
    .cl_lib_config = {"CL_HIER lib", "CL_HIER_", cl_hier_lib_config_table, sizeof(ucc_cl_hier_lib_config_t), {NULL, NULL}},
    .cl_context_config = {"CL_HIER context", "CL_HIER_", cl_hier_context_config_table, sizeof(ucc_cl_hier_context_config_t), {NULL, NULL}},
    .lib = {hier_lib_init, hier_lib_finalize, hier_lib_get_attr, NULL},
    .context = {hier_ctx_create, NULL, hier_ctx_destroy, hier_ctx_get_attr, NULL, NULL, NULL},
    .team = {hier_team_create_post, hier_team_create_test, hier_team_destroy, hier_team_get_scores},
    .coll = {hier_coll_init},
};
UCC_COMPONENT_REGISTER_STATIC(cl, ucc_cl_hier)
static void UCC_CTOR cl_hier_register(void)
{
    ucc_config_table_register(&ucc_cl_hier.cl_lib_config); ucc_config_table_register(&ucc_cl_hier.cl_context_config);
    for (int c = 0; c < UCC_COLL_TYPE_NUM; c++) {
        const hier_alg_t *algs = hier_algs((ucc_coll_type_t)UCC_BIT(c)); int i = 0;
        if (!algs) continue;
        for (; algs[i].name && i < 2; i++) { hier_alg_info[c][i].id = (unsigned)i; hier_alg_info[c][i].name = algs[i].name; hier_alg_info[c][i].desc = algs[i].desc; }
        hier_alg_info[c][i].name = NULL; ucc_cl_hier.alg_info[c] = hier_alg_info[c];
    }
    (void)hier_sbgp_names;
}
