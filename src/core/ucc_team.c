/* Team creation state machine:
 *   ADDR_EXCHANGE -> SERVICE_TEAM -> ALLOC_ID -> CL_CREATE -> ACTIVE
 * Behaviour: reference core/ucc_team.c:93-646 (parameter validation, id
 * allocation by a service allreduce(BAND) over the free-id bitmap, CL teams
 * created one at a time, merged score map). */
#include "ucc_team.h"
#include "ucc_global_opts.h"
#include "ucc_service_coll.h"
#include "components/topo/ucc_topo.h"
#include "utils/ucc_string.h"

static ucc_status_t team_create_service_team(ucc_context_t *ctx, ucc_team_t *team);
static ucc_status_t team_alloc_id(ucc_team_t *team);

UCC_EXPORT ucc_status_t ucc_team_get_attr(ucc_team_h team, ucc_team_attr_t *attr)
{
    if (!team || !attr) return UCC_ERR_INVALID_PARAM;
    if (attr->mask & (UCC_TEAM_ATTR_FIELD_POST_ORDERING | UCC_TEAM_ATTR_FIELD_OUTSTANDING_CALLS | UCC_TEAM_ATTR_FIELD_EP_RANGE |
                      UCC_TEAM_ATTR_FIELD_SYNC_TYPE | UCC_TEAM_ATTR_FIELD_MEM_PARAMS | UCC_TEAM_ATTR_FIELD_EPS)) return UCC_ERR_NOT_SUPPORTED;
    if (attr->mask & UCC_TEAM_ATTR_FIELD_SIZE) attr->size = team->size;
    if (attr->mask & UCC_TEAM_ATTR_FIELD_EP) attr->ep = team->rank;
    return UCC_OK;
}

UCC_EXPORT ucc_status_t ucc_team_create_from_parent(uint64_t my_ep, uint32_t included, ucc_team_h parent_team, ucc_team_h *new_team)
{ (void)my_ep; (void)included; (void)parent_team; (void)new_team; return UCC_ERR_NOT_IMPLEMENTED; }

UCC_EXPORT ucc_status_t ucc_team_create_post(ucc_context_h *contexts, uint32_t num_contexts, const ucc_team_params_t *params, ucc_team_h *new_team)
{
    uint64_t team_size = 0;
    ucc_rank_t rank = UCC_RANK_INVALID;
    ucc_team_t *team;
    if (!contexts || !params || !new_team || num_contexts < 1) return UCC_ERR_INVALID_PARAM;
    if (num_contexts > 1) { ucc_error("team creation from multiple contexts is not supported yet"); return UCC_ERR_NOT_SUPPORTED; }
    if (params->mask & UCC_TEAM_PARAM_FIELD_TEAM_SIZE) team_size = params->team_size;
    if (params->mask & UCC_TEAM_PARAM_FIELD_OOB) {
        if (team_size > 0 && params->oob.n_oob_eps != team_size) { ucc_error("inconsistent team_sizes provided as params.team_size %llu and params.oob.n_oob_eps %llu",
            (unsigned long long)team_size, (unsigned long long)params->oob.n_oob_eps); return UCC_ERR_INVALID_PARAM; }
        team_size = params->oob.n_oob_eps; rank = params->oob.oob_ep;
    }
    if (params->mask & UCC_TEAM_PARAM_FIELD_EP_MAP) {
        if (team_size > 0 && params->ep_map.ep_num != team_size) { ucc_error("inconsistent team_sizes provided as team_size %llu and ep_map.ep_num %llu",
            (unsigned long long)team_size, (unsigned long long)params->ep_map.ep_num); return UCC_ERR_INVALID_PARAM; }
        team_size = params->ep_map.ep_num;
    }
    if (team_size < 1) { ucc_warn("minimal size of UCC team is 1, provided %llu", (unsigned long long)team_size); return UCC_ERR_INVALID_PARAM; }
    if ((params->mask & UCC_TEAM_PARAM_FIELD_EP) && (params->mask & UCC_TEAM_PARAM_FIELD_EP_RANGE) && params->ep_range == UCC_COLLECTIVE_EP_RANGE_CONTIG) {
        if (rank != UCC_RANK_INVALID && rank != (ucc_rank_t)params->ep) { ucc_error("inconsistent EP value (%llu) and oob.oob_ep (%u)", (unsigned long long)params->ep, rank); return UCC_ERR_INVALID_PARAM; }
        rank = (ucc_rank_t)params->ep;
    } else if (rank == UCC_RANK_INVALID) {
        ucc_error("either UCC_TEAM_PARAM_FIELD_EP with EP_RANGE_CONTIG or UCC_TEAM_PARAM_FIELD_OOB must be provided to define the team rank");
        return UCC_ERR_INVALID_PARAM;
    }
    if (rank >= team_size) { ucc_error("team rank %u is out of range [0,%llu)", rank, (unsigned long long)team_size); return UCC_ERR_INVALID_PARAM; }
    team = (ucc_team_t *)calloc(1, sizeof(*team));
    if (!team) return UCC_ERR_NO_MEMORY;
    team->contexts = (ucc_context_t **)malloc(sizeof(void *) * num_contexts);
    if (!team->contexts) { free(team); return UCC_ERR_NO_MEMORY; }
    memcpy(team->contexts, contexts, sizeof(void *) * num_contexts);
    team->num_contexts = num_contexts; team->size = (ucc_rank_t)team_size; team->rank = rank;
    team->bp.params = *params; team->bp.rank = rank; team->bp.size = team->size; team->bp.team = team; team->bp.scope = UCC_CL_LAST;
    team->bp.map = ucc_ep_map_create_full(team->size);
    if (params->mask & UCC_TEAM_PARAM_FIELD_OOB) team->oob = params->oob;
    if (params->mask & UCC_TEAM_PARAM_FIELD_ID) {
        if (params->id > UCC_TEAM_ID_MAX) { ucc_error("team id %llu exceeds the maximum %u", (unsigned long long)params->id, UCC_TEAM_ID_MAX); free(team->contexts); free(team); return UCC_ERR_INVALID_PARAM; }
        team->id = (uint16_t)(params->id | UCC_TEAM_ID_EXTERNAL_BIT);
    }
    team->cl_teams = (ucc_cl_team_t **)calloc((size_t)contexts[0]->n_cl_ctx + 1, sizeof(void *));
    team->last_team_create_posted = -1;
    team->status = UCC_INPROGRESS;
    team->state = team->size > 1 ? UCC_TEAM_ADDR_EXCHANGE : UCC_TEAM_ALLOC_ID;
    if (team->size == 1) {
        team->ctx_map = ucc_ep_map_create_full(1);
        if (contexts[0]->addr_storage.storage) { ucc_rank_t *a = (ucc_rank_t *)malloc(sizeof(*a)); a[0] = contexts[0]->rank; team->ctx_ranks = a;
            team->ctx_map = ucc_ep_map_from_array(&team->ctx_ranks, 1, contexts[0]->addr_storage.size, 0); }
    }
    *new_team = team;
    return UCC_OK;
}

/* learn which context rank each team member is (or exchange full addresses when the context has none) */
static ucc_status_t team_exchange(ucc_context_t *ctx, ucc_team_t *team)
{
    ucc_status_t st;
    if (!ctx->addr_storage.storage) {
        if (!(team->bp.params.mask & UCC_TEAM_PARAM_FIELD_OOB)) { ucc_error("team OOB is required when the context was created without OOB"); return UCC_ERR_INVALID_PARAM; }
        st = ucc_core_addr_exchange(ctx, &team->oob, &team->addr_storage);
        if (st == UCC_OK) team->ctx_map = ucc_ep_map_create_full(team->size);
        return st;
    }
    if (team->bp.params.mask & UCC_TEAM_PARAM_FIELD_EP_MAP) { team->ctx_map = team->bp.params.ep_map; return UCC_OK; }
    if (!(team->bp.params.mask & UCC_TEAM_PARAM_FIELD_OOB)) {
        if (team->size == ctx->addr_storage.size) { team->ctx_map = ucc_ep_map_create_full(team->size); return UCC_OK; }
        ucc_error("team of size %u over a context of size %u needs OOB or EP_MAP", team->size, ctx->addr_storage.size);
        return UCC_ERR_INVALID_PARAM;
    }
    if (!team->ctx_ranks) {
        team->ctx_ranks = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * team->size);
        if (!team->ctx_ranks) return UCC_ERR_NO_MEMORY;
        team->ctx_rank_send = ctx->rank;
        st = team->oob.allgather(&team->ctx_rank_send, team->ctx_ranks, sizeof(ucc_rank_t), team->oob.coll_info, &team->oob_req);
        if (st != UCC_OK) { ucc_error("failed to start oob allgather for ctx ranks exchange"); return st; }
    }
    st = team->oob.req_test(team->oob_req);
    if (st < 0) { team->oob.req_free(team->oob_req); ucc_error("oob req test failed during team addr exchange"); return st; }
    if (st == UCC_INPROGRESS) return st;
    team->oob.req_free(team->oob_req); team->oob_req = NULL;
    team->ctx_map = ucc_ep_map_from_array(&team->ctx_ranks, team->size, ctx->addr_storage.size, 0);
    return UCC_OK;
}

static ucc_status_t team_build_score_map(ucc_team_t *team)
{
    ucc_coll_score_t *score = NULL, *cl_score;
    ucc_status_t st;
    for (int i = 0; i < team->n_cl_teams; i++) {
        st = UCC_CL_TEAM_IFACE(team->cl_teams[i])->team.get_scores(&team->cl_teams[i]->super, &cl_score);
        if (st != UCC_OK) { ucc_error("failed to get cl %s scores", UCC_CL_TEAM_IFACE(team->cl_teams[i])->super.name); if (score) ucc_coll_score_free(score); return st; }
        if (!score) { score = cl_score; continue; }
        st = ucc_coll_score_merge_in(&score, cl_score);
        if (st != UCC_OK) return st;
    }
    return ucc_coll_score_build_map(score, &team->score_map);
}

static ucc_status_t team_create_cls(ucc_context_t *ctx, ucc_team_t *team)
{
    ucc_base_team_t *b_team;
    ucc_status_t st;
    int i = team->last_team_create_posted;
    if (i < 0) {
        /* first entry into this state: topology view of the team */
        if (ctx->topo || team->addr_storage.storage) {
            ucc_subset_t set; set.map = team->ctx_map; set.myrank = team->rank;
            if (!ctx->topo) { /* context without OOB: the team's own storage describes exactly the team */
                st = ucc_context_topo_init(&team->addr_storage, &team->own_ctx_topo);
                if (st != UCC_OK) return st;
            }
            st = ucc_topo_init(set, ctx->topo ? ctx->topo : team->own_ctx_topo, &team->topo);
            if (st != UCC_OK) { ucc_warn("failed to init team topo"); team->topo = NULL; }
        }
    }
    if (i >= 0 && team->cl_teams[team->n_cl_teams]) {
        ucc_cl_iface_t *iface = UCC_CL_CTX_IFACE(ctx->cl_ctx[i]);
        st = iface->team.create_test(&team->cl_teams[team->n_cl_teams]->super);
        if (st == UCC_INPROGRESS) return st;
        if (st == UCC_OK) team->n_cl_teams++;
        else { ucc_debug("failed to create CL %s team: %s", iface->super.name, ucc_status_string(st)); team->cl_teams[team->n_cl_teams] = NULL; }
    }
    for (i = team->last_team_create_posted + 1; i < ctx->n_cl_ctx; i++) {
        ucc_cl_iface_t *iface = UCC_CL_CTX_IFACE(ctx->cl_ctx[i]);
        team->last_team_create_posted = i;
        st = iface->team.create_post(&ctx->cl_ctx[i]->super, &team->bp, &b_team);
        if (st != UCC_OK) { ucc_debug("failed to create CL %s team: %s", iface->super.name, ucc_status_string(st)); continue; }
        team->cl_teams[team->n_cl_teams] = ucc_derived_of(b_team, ucc_cl_team_t);
        st = iface->team.create_test(b_team);
        if (st == UCC_INPROGRESS) return st;
        if (st == UCC_OK) team->n_cl_teams++;
        else { ucc_debug("failed to create CL %s team: %s", iface->super.name, ucc_status_string(st)); team->cl_teams[team->n_cl_teams] = NULL; }
    }
    if (team->n_cl_teams == 0) { ucc_error("no CL teams were created"); return UCC_ERR_NO_MESSAGE; }
    return UCC_OK;
}

static ucc_status_t team_create_test_single(ucc_context_t *ctx, ucc_team_t *team)
{
    ucc_status_t st = UCC_OK;
    switch (team->state) {
    case UCC_TEAM_ADDR_EXCHANGE:
        st = team_exchange(ctx, team);
        if (st != UCC_OK) break;
        team->state = UCC_TEAM_SERVICE_TEAM;
        /* fall through */
    case UCC_TEAM_SERVICE_TEAM:
        /* a team-private service team is only needed when the context has none */
        if (team->size > 1 && !ctx->service_team && !UCC_TEAM_ID_IS_EXTERNAL(team) ) {
            st = team_create_service_team(ctx, team);
            if (st == UCC_INPROGRESS) break;
            if (st != UCC_OK) { ucc_error("a service team is required to allocate the team id (or pass UCC_TEAM_PARAM_FIELD_ID)"); break; }
        }
        team->state = UCC_TEAM_ALLOC_ID;
        /* fall through */
    case UCC_TEAM_ALLOC_ID:
        st = team_alloc_id(team);
        if (st != UCC_OK) break;
        team->bp.id = team->id;
        team->state = UCC_TEAM_CL_CREATE;
        /* fall through */
    case UCC_TEAM_CL_CREATE:
        st = team_create_cls(ctx, team);
        if (st != UCC_OK) break;
        st = team_build_score_map(team);
        if (st != UCC_OK) break;
        team->state = UCC_TEAM_ACTIVE;
        if (team->rank == 0 && ucc_global_log_component.log_level >= UCC_LOG_LEVEL_INFO) {
            ucc_info("===== COLL_SCORE_MAP (team_id %u, size %u) =====", team->id, team->size);
            ucc_coll_score_map_print_info(team->score_map, UCC_LOG_LEVEL_INFO);
            ucc_info("================================================");
        }
        /* fall through */
    case UCC_TEAM_ACTIVE:
        st = UCC_OK;
        break;
    }
    team->status = st;
    return st;
}

UCC_EXPORT ucc_status_t ucc_team_create_test(ucc_team_h team)
{
    if (!team) return UCC_ERR_INVALID_PARAM;
    if (team->status == UCC_OK) return UCC_OK;
    if (team->status < 0) return team->status;
    return team_create_test_single(team->contexts[0], team);
}

/* ---- service team private to this team ---- */
static ucc_status_t team_create_service_team(ucc_context_t *ctx, ucc_team_t *team)
{
    ucc_status_t st;
    if (!team->service_team) {
        ucc_base_team_params_t bp = team->bp;
        ucc_base_team_t *b_team;
        ucc_tl_context_t *tl_ctx;
        const char *svc = "shm";
        st = ucc_tl_context_get(ctx, svc, &tl_ctx);
        if (st != UCC_OK) { ucc_warn("TL %s context is not available, service team can not be created", svc); return st; }
        bp.id = 0; bp.scope = UCC_CL_LAST; bp.scope_id = 1; bp.map = ucc_ep_map_create_full(team->size);
        st = UCC_TL_CTX_IFACE(tl_ctx)->team.create_post(&tl_ctx->super, &bp, &b_team);
        if (st != UCC_OK) { ucc_tl_context_put(tl_ctx); return st; }
        team->service_team = ucc_derived_of(b_team, ucc_tl_team_t);
        team->service_team_owned = 1;
    }
    st = UCC_TL_TEAM_IFACE(team->service_team)->team.create_test(&team->service_team->super);
    if (st < 0) { team->service_team = NULL; team->service_team_owned = 0; }
    return st;
}

/* ---- team id ---- */
static ucc_status_t team_alloc_id(ucc_team_t *team)
{
    ucc_context_t *ctx = team->contexts[0];
    uint64_t *local, *global;
    ucc_status_t st;
    int pos = -1;
    if (team->id != 0) return UCC_OK; /* user supplied */
    if (team->size == 1) { /* local decision is enough */
        for (uint32_t i = 0; i < ctx->ids_pool_size && pos < 0; i++) if (ctx->ids_pool[i]) { pos = (int)(i * 64 + ucc_ffs64(ctx->ids_pool[i])); ctx->ids_pool[i] &= ~(1ull << (pos % 64)); }
        if (pos <= 0) { ucc_warn("could not allocate team id: the pool is exhausted (TEAM_IDS_POOL_SIZE)"); return UCC_ERR_NO_RESOURCE; }
        team->id = (uint16_t)pos;
        return UCC_OK;
    }
    if (!team->id_bitmap_work) {
        team->id_bitmap_work = (uint64_t *)malloc(2 * sizeof(uint64_t) * ctx->ids_pool_size);
        if (!team->id_bitmap_work) return UCC_ERR_NO_MEMORY;
    }
    local = team->id_bitmap_work; global = local + ctx->ids_pool_size;
    if (!team->sreq) {
        ucc_subset_t subset; subset.map = ucc_ep_map_create_full(team->size); subset.myrank = team->rank;
        memcpy(local, ctx->ids_pool, sizeof(uint64_t) * ctx->ids_pool_size);
        st = ucc_service_allreduce(team, local, global, UCC_DT_UINT64, ctx->ids_pool_size, UCC_OP_BAND, subset, &team->sreq);
        if (st < 0) return st;
    }
    st = ucc_service_coll_test(team->sreq);
    if (st < 0) { ucc_service_coll_finalize(team->sreq); team->sreq = NULL; return st; }
    if (st != UCC_OK) return st;
    ucc_service_coll_finalize(team->sreq); team->sreq = NULL;
    for (uint32_t i = 0; i < ctx->ids_pool_size && pos < 0; i++) if (global[i]) pos = (int)(i * 64 + ucc_ffs64(global[i]));
    free(team->id_bitmap_work); team->id_bitmap_work = NULL;
    if (pos <= 0) { ucc_warn("could not allocate team id: no id is free on every rank"); return UCC_ERR_NO_RESOURCE; }
    ctx->ids_pool[pos / 64] &= ~(1ull << (pos % 64));
    team->id = (uint16_t)pos;
    return UCC_OK;
}

void ucc_team_release_id(ucc_team_t *team)
{
    ucc_context_t *ctx = team->contexts[0];
    if (team->id && !UCC_TEAM_ID_IS_EXTERNAL(team)) ctx->ids_pool[team->id / 64] |= 1ull << (team->id % 64);
}

/* ---- destroy ---- */
UCC_EXPORT ucc_status_t ucc_team_destroy(ucc_team_h team)
{
    ucc_status_t st;
    if (!team) return UCC_ERR_INVALID_PARAM;
    if (team->status != UCC_OK && team->state == UCC_TEAM_ACTIVE) {}
    if (team->status == UCC_INPROGRESS) { ucc_error("team %p is used before team_create is completed", (void *)team); return UCC_ERR_INVALID_PARAM; }
    for (int i = team->n_cl_teams - 1; i >= 0; i--) {
        if (!team->cl_teams[i]) continue;
        st = UCC_CL_TEAM_IFACE(team->cl_teams[i])->team.destroy(&team->cl_teams[i]->super);
        if (st == UCC_INPROGRESS) return st;
        if (st != UCC_OK) ucc_error("cl team destroy failed: %s", ucc_status_string(st));
        team->cl_teams[i] = NULL; team->n_cl_teams = i;
    }
    if (team->service_team && team->service_team_owned) {
        ucc_tl_context_t *tl_ctx = ucc_derived_of(team->service_team->super.context, ucc_tl_context_t);
        st = UCC_TL_TEAM_IFACE(team->service_team)->team.destroy(&team->service_team->super);
        if (st == UCC_INPROGRESS) return st;
        ucc_tl_context_put(tl_ctx);
        team->service_team = NULL;
    }
    ucc_team_release_id(team);
    if (team->topo) ucc_topo_cleanup(team->topo);
    if (team->own_ctx_topo) ucc_context_topo_cleanup(team->own_ctx_topo);
    if (team->score_map) ucc_coll_score_free_map(team->score_map);
    ucc_addr_storage_free(&team->addr_storage);
    free(team->ctx_ranks); free(team->id_bitmap_work); free(team->cl_teams); free(team->contexts); free(team);
    return UCC_OK;
}
