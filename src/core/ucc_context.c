/* Context: creates TL then CL contexts, the progress queue, exchanges
 * packed addresses over the user's OOB allgather, builds the context
 * topology and the context-wide service team.
 * Behaviour: reference core/ucc_context.c:377-1210. */
#include "ucc_context.h"
#include "ucc_team.h"
#include "ucc_global_opts.h"
#include "ucc_service_coll.h"
#include "components/topo/ucc_topo.h"
#include "utils/ucc_string.h"
#include <unistd.h>

static ucc_config_field_t ucc_context_config_table[] = {
    {"ESTIMATED_NUM_EPS", "0", "An optimization hint of how many endpoints will be created on this context",
     ucc_offsetof(ucc_context_config_t, estimated_num_eps), UCC_CONFIG_TYPE_UINT},
    {"LOCK_FREE_PROGRESS_Q", "0", "Enable lock free progress queue optimization (THREAD_MULTIPLE only)",
     ucc_offsetof(ucc_context_config_t, lock_free_progress_q), UCC_CONFIG_TYPE_UINT},
    {"ESTIMATED_NUM_PPN", "0", "An optimization hint of how many endpoints created on this context reside on the same node",
     ucc_offsetof(ucc_context_config_t, estimated_num_ppn), UCC_CONFIG_TYPE_UINT},
    {"TEAM_IDS_POOL_SIZE", "32", "Defines the size of the team_id_pool. The number of coexisting unique team ids for a single "
     "process is team_ids_pool_size*64. This parameter is relevant when internal team id allocation takes place.",
     ucc_offsetof(ucc_context_config_t, team_ids_pool_size), UCC_CONFIG_TYPE_UINT},
    {"INTERNAL_OOB", "1", "Use internal OOB (service team allgather) for team creation. If set to 0 the user-provided team OOB is used.",
     ucc_offsetof(ucc_context_config_t, internal_oob), UCC_CONFIG_TYPE_UINT},
    {"THROTTLE_PROGRESS", "1000", "Throttle UCC progress to every <n>th invocation when the progress queue is empty",
     ucc_offsetof(ucc_context_config_t, throttle_progress), UCC_CONFIG_TYPE_UINT},
    {"NET_DEVICES", "", "Specifies which network device(s) to use (passed to network TLs)",
     ucc_offsetof(ucc_context_config_t, net_devices), UCC_CONFIG_TYPE_STRING_ARRAY},
    {"NODE_LOCAL_ID", "auto", "Rank of the process on its node (auto: derived from the address exchange)",
     ucc_offsetof(ucc_context_config_t, node_local_id), UCC_CONFIG_TYPE_UINT},
    {"SERVICE_TL", "shm", "TL that implements the internal service collectives (team-id allocation, internal OOB)",
     ucc_offsetof(ucc_context_config_t, service_tl), UCC_CONFIG_TYPE_STRING},
    {NULL}};
UCC_CONFIG_REGISTER_TABLE(ucc_context_config_table, "UCC context", "", ucc_context_config_t)

/* ------------------------------------------------------------------ */
/* config                                                              */
/* ------------------------------------------------------------------ */
UCC_EXPORT ucc_status_t ucc_context_config_read(ucc_lib_h lib, const char *filename, ucc_context_config_h *config_p)
{
    ucc_context_config_t *cfg;
    ucc_status_t st;
    if (filename) { ucc_error("read from file is not implemented"); return UCC_ERR_NOT_IMPLEMENTED; }
    if (!lib || !config_p) return UCC_ERR_INVALID_PARAM;
    cfg = (ucc_context_config_t *)calloc(1, sizeof(*cfg));
    if (!cfg) return UCC_ERR_NO_MEMORY;
    cfg->lib = lib;
    st = ucc_config_parser_fill_opts_table(cfg, ucc_context_config_table, lib->full_prefix, "", 0);
    if (st != UCC_OK) { free(cfg); return st; }
    cfg->cl_cfgs = (ucc_cl_context_config_t **)calloc((size_t)lib->n_cl_libs_opened + 1, sizeof(void *));
    cfg->tl_cfgs = (ucc_tl_context_config_t **)calloc((size_t)lib->n_tl_libs_opened + 1, sizeof(void *));
    for (int i = 0; i < lib->n_cl_libs_opened; i++) {
        st = ucc_cl_context_config_read(lib->cl_libs[i], lib->full_prefix, &cfg->cl_cfgs[cfg->n_cl_cfg]);
        if (st != UCC_OK) { ucc_error("failed to read CL \"%s\" context configuration", lib->cl_libs[i]->iface->super.name); goto err; }
        cfg->n_cl_cfg++;
    }
    for (int i = 0; i < lib->n_tl_libs_opened; i++) {
        st = ucc_tl_context_config_read(lib->tl_libs[i], lib->full_prefix, &cfg->tl_cfgs[cfg->n_tl_cfg]);
        if (st != UCC_OK) { ucc_error("failed to read TL \"%s\" context configuration", lib->tl_libs[i]->iface->super.name); goto err; }
        cfg->n_tl_cfg++;
    }
    *config_p = cfg;
    return UCC_OK;
err:
    ucc_context_config_release(cfg);
    return st;
}

UCC_EXPORT void ucc_context_config_release(ucc_context_config_h cfg)
{
    if (!cfg) return;
    for (int i = 0; i < cfg->n_cl_cfg; i++) ucc_cl_context_config_release(cfg->cl_cfgs[i]->cl_lib->iface, cfg->cl_cfgs[i]);
    for (int i = 0; i < cfg->n_tl_cfg; i++) ucc_tl_context_config_release(cfg->tl_cfgs[i]->tl_lib->iface, cfg->tl_cfgs[i]);
    ucc_config_parser_release_opts(cfg, ucc_context_config_table);
    free(cfg->cl_cfgs); free(cfg->tl_cfgs); free(cfg);
}

UCC_EXPORT void ucc_context_config_print(const ucc_context_config_h cfg, FILE *stream, const char *title, ucc_config_print_flags_t flags)
{
    ucc_config_parser_print_opts(stream, title, cfg, ucc_context_config_table, "", cfg->lib->full_prefix, flags);
    for (int i = 0; i < cfg->n_cl_cfg; i++) {
        ucc_cl_iface_t *iface = cfg->cl_cfgs[i]->cl_lib->iface;
        ucc_config_parser_print_opts(stream, iface->cl_context_config.name, cfg->cl_cfgs[i], iface->cl_context_config.table,
                                     iface->cl_context_config.prefix, cfg->lib->full_prefix, flags);
    }
    for (int i = 0; i < cfg->n_tl_cfg; i++) {
        ucc_tl_iface_t *iface = cfg->tl_cfgs[i]->tl_lib->iface;
        ucc_config_parser_print_opts(stream, iface->tl_context_config.name, cfg->tl_cfgs[i], iface->tl_context_config.table,
                                     iface->tl_context_config.prefix, cfg->lib->full_prefix, flags);
    }
}

/* component is "cl/<name>", "tl/<name>" or NULL for the core table */
UCC_EXPORT ucc_status_t ucc_context_config_modify(ucc_context_config_h cfg, const char *component, const char *name, const char *value)
{
    if (!cfg || !name || !value) return UCC_ERR_INVALID_PARAM;
    if (!component) return ucc_config_parser_set_value(cfg, ucc_context_config_table, name, value);
    if (!strncmp(component, "cl/", 3)) {
        for (int i = 0; i < cfg->n_cl_cfg; i++)
            if (!strcmp(component + 3, cfg->cl_cfgs[i]->cl_lib->iface->super.name))
                return ucc_config_parser_set_value(cfg->cl_cfgs[i], cfg->cl_cfgs[i]->cl_lib->iface->cl_context_config.table, name, value);
        ucc_error("required CL %s is not part of the context config", component);
        return UCC_ERR_INVALID_PARAM;
    }
    if (!strncmp(component, "tl/", 3)) {
        for (int i = 0; i < cfg->n_tl_cfg; i++)
            if (!strcmp(component + 3, cfg->tl_cfgs[i]->tl_lib->iface->super.name))
                return ucc_config_parser_set_value(cfg->tl_cfgs[i], cfg->tl_cfgs[i]->tl_lib->iface->tl_context_config.table, name, value);
        ucc_error("required TL %s is not part of the context config", component);
        return UCC_ERR_INVALID_PARAM;
    }
    ucc_error("invalid component name \"%s\": expected cl/<name> or tl/<name>", component);
    return UCC_ERR_INVALID_PARAM;
}

/* ------------------------------------------------------------------ */
/* address packing / exchange                                          */
/* ------------------------------------------------------------------ */
static ucc_status_t ucc_context_pack_addr(ucc_context_t *ctx, void **addr_p, size_t *len_p)
{
    int n = ctx->n_cl_ctx + ctx->n_tl_ctx, k = 0;
    size_t *lens = (size_t *)calloc((size_t)n + 1, sizeof(size_t)), total, off;
    ucc_base_ctx_attr_t attr;
    ucc_context_addr_header_t *h;
    if (!lens) return UCC_ERR_NO_MEMORY;
    for (int i = 0; i < ctx->n_cl_ctx; i++, k++) {
        memset(&attr, 0, sizeof(attr)); attr.attr.mask = UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN;
        UCC_CL_CTX_IFACE(ctx->cl_ctx[i])->context.get_attr(&ctx->cl_ctx[i]->super, &attr);
        lens[k] = attr.attr.ctx_addr_len;
    }
    for (int i = 0; i < ctx->n_tl_ctx; i++, k++) {
        memset(&attr, 0, sizeof(attr)); attr.attr.mask = UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN;
        UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->context.get_attr(&ctx->tl_ctx[i]->super, &attr);
        lens[k] = attr.attr.ctx_addr_len;
    }
    total = UCC_CONTEXT_ADDR_HEADER_SIZE(n);
    for (int i = 0; i < n; i++) total += ucc_align_up(lens[i], 8);
    h = (ucc_context_addr_header_t *)calloc(1, total);
    if (!h) { free(lens); return UCC_ERR_NO_MEMORY; }
    h->ctx_id = ctx->id; h->host_info = ctx->host_info; h->n_components = n;
    off = UCC_CONTEXT_ADDR_HEADER_SIZE(n);
    k = 0;
    for (int i = 0; i < ctx->n_cl_ctx; i++, k++) {
        h->components[k].id = UCC_CL_CTX_IFACE(ctx->cl_ctx[i])->super.id; h->components[k].offset = (uint32_t)off; h->components[k].len = (uint32_t)lens[k];
        if (lens[k]) { memset(&attr, 0, sizeof(attr)); attr.attr.mask = UCC_CONTEXT_ATTR_FIELD_CTX_ADDR; attr.attr.ctx_addr = (char *)h + off;
            UCC_CL_CTX_IFACE(ctx->cl_ctx[i])->context.get_attr(&ctx->cl_ctx[i]->super, &attr); }
        off += ucc_align_up(lens[k], 8);
    }
    for (int i = 0; i < ctx->n_tl_ctx; i++, k++) {
        h->components[k].id = UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->super.id; h->components[k].offset = (uint32_t)off; h->components[k].len = (uint32_t)lens[k];
        if (lens[k]) { memset(&attr, 0, sizeof(attr)); attr.attr.mask = UCC_CONTEXT_ATTR_FIELD_CTX_ADDR; attr.attr.ctx_addr = (char *)h + off;
            UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->context.get_attr(&ctx->tl_ctx[i]->super, &attr); }
        off += ucc_align_up(lens[k], 8);
    }
    free(lens);
    *addr_p = h; *len_p = total;
    return UCC_OK;
}

void *ucc_addr_storage_component_addr(ucc_addr_storage_t *s, ucc_rank_t rank, unsigned long component_id, uint32_t *len)
{
    ucc_context_addr_header_t *h;
    if (!s->storage || rank >= s->size) return NULL;
    h = UCC_ADDR_STORAGE_RANK_HEADER(s, rank);
    for (int i = 0; i < h->n_components; i++)
        if (h->components[i].id == component_id) { if (len) *len = h->components[i].len; return (char *)h + h->components[i].offset; }
    return NULL;
}

void ucc_addr_storage_free(ucc_addr_storage_t *s) { free(s->storage); memset(s, 0, sizeof(*s)); }

/* state kept in storage->oob_req between calls */
typedef struct addr_exchange_state { int phase; void *req; size_t *lens; size_t my_len; void *my_addr; void *padded; } addr_exchange_state_t;

ucc_status_t ucc_core_addr_exchange(ucc_context_t *ctx, ucc_oob_coll_t *oob, ucc_addr_storage_t *s)
{
    addr_exchange_state_t *x = (addr_exchange_state_t *)s->oob_req;
    ucc_status_t st;
    if (!x) {
        x = (addr_exchange_state_t *)calloc(1, sizeof(*x));
        if (!x) return UCC_ERR_NO_MEMORY;
        s->oob_req = x; s->size = oob->n_oob_eps; s->rank = oob->oob_ep;
        if (!ctx->attr.ctx_addr) { st = ucc_context_pack_addr(ctx, &ctx->attr.ctx_addr, &ctx->attr.ctx_addr_len); if (st != UCC_OK) goto err; }
        x->my_addr = ctx->attr.ctx_addr; x->my_len = ctx->attr.ctx_addr_len;
        x->lens = (size_t *)calloc(s->size, sizeof(size_t));
        if (!x->lens) { st = UCC_ERR_NO_MEMORY; goto err; }
        st = oob->allgather(&x->my_len, x->lens, sizeof(size_t), oob->coll_info, &x->req);
        if (st != UCC_OK) { ucc_error("failed to start oob allgather"); goto err; }
        x->phase = 1;
    }
    if (x->phase == 1) {
        st = oob->req_test(x->req);
        if (st == UCC_INPROGRESS) return UCC_INPROGRESS;
        if (st < 0) goto err;
        oob->req_free(x->req); x->req = NULL;
        s->addr_len = 0;
        for (ucc_rank_t i = 0; i < s->size; i++) if (x->lens[i] > s->addr_len) s->addr_len = x->lens[i];
        s->addr_len = ucc_align_up(s->addr_len, 8);
        s->storage = calloc(s->size, s->addr_len); x->padded = calloc(1, s->addr_len);
        if (!s->storage || !x->padded) { st = UCC_ERR_NO_MEMORY; goto err; }
        memcpy(x->padded, x->my_addr, x->my_len);
        st = oob->allgather(x->padded, s->storage, s->addr_len, oob->coll_info, &x->req);
        if (st != UCC_OK) goto err;
        x->phase = 2;
    }
    if (x->phase == 2) {
        int same_node = 1;
        st = oob->req_test(x->req);
        if (st == UCC_INPROGRESS) return UCC_INPROGRESS;
        if (st < 0) goto err;
        oob->req_free(x->req);
        for (ucc_rank_t i = 0; i < s->size; i++)
            if (UCC_ADDR_STORAGE_RANK_HEADER(s, i)->ctx_id.pi.host_hash != ctx->id.pi.host_hash) same_node = 0;
        s->flags = same_node;
        free(x->lens); free(x->padded); free(x); s->oob_req = NULL;
        return UCC_OK;
    }
    return UCC_INPROGRESS;
err:
    if (x) { free(x->lens); free(x->padded); free(x); }
    s->oob_req = NULL; free(s->storage); s->storage = NULL;
    return st < 0 ? st : UCC_ERR_NO_MESSAGE;
}

/* ------------------------------------------------------------------ */
/* create                                                              */
/* ------------------------------------------------------------------ */
static ucc_status_t create_tl_contexts(ucc_context_t *ctx, ucc_context_config_t *cfg, ucc_base_context_params_t *bp)
{
    ucc_lib_info_t *lib = ctx->lib;
    ctx->tl_ctx = (ucc_tl_context_t **)calloc((size_t)cfg->n_tl_cfg + 1, sizeof(void *));
    if (!ctx->tl_ctx) return UCC_ERR_NO_MEMORY;
    ctx->n_tl_ctx = 0;
    for (int i = 0; i < cfg->n_tl_cfg; i++) {
        ucc_tl_lib_t *tl_lib = cfg->tl_cfgs[i]->tl_lib;
        ucc_base_context_t *b_ctx = NULL;
        ucc_status_t st;
        if (cfg->tl_cfgs[i]->super.score_str && !strcmp(cfg->tl_cfgs[i]->super.score_str, "0")) {
            ucc_debug("TL %s is disabled by TUNE=0", tl_lib->iface->super.name);
            continue;
        }
        st = tl_lib->iface->context.create(bp, &cfg->tl_cfgs[i]->super, &b_ctx);
        if (st != UCC_OK) {
            if (ucc_tl_is_required(lib, tl_lib->iface, 1)) ucc_warn("failed to create tl context for %s that was explicitly requested", tl_lib->iface->super.name);
            else ucc_debug("failed to create tl context for %s, skipping", tl_lib->iface->super.name);
            continue;
        }
        ctx->tl_ctx[ctx->n_tl_ctx] = ucc_derived_of(b_ctx, ucc_tl_context_t);
        ctx->tl_ctx[ctx->n_tl_ctx]->ref_count = 0;
        ctx->n_tl_ctx++;
    }
    if (ctx->n_tl_ctx == 0) { ucc_error("no tl contexts were created"); return UCC_ERR_NOT_FOUND; }
    ctx->all_tls.names = (char **)calloc((size_t)ctx->n_tl_ctx, sizeof(char *)); ctx->all_tls.count = 0;
    for (int i = 0; i < ctx->n_tl_ctx; i++) ctx->all_tls.names[ctx->all_tls.count++] = strdup(UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->super.name);
    return UCC_OK;
}

static ucc_status_t create_service_team(ucc_context_t *ctx, ucc_context_config_t *cfg)
{
    ucc_base_team_params_t bp;
    ucc_base_team_t *b_team = NULL;
    ucc_team_t *stub;
    ucc_status_t st = ucc_tl_context_get(ctx, cfg->service_tl, &ctx->service_ctx);
    if (st != UCC_OK) { ucc_warn("TL %s context is not available, service team can not be created", cfg->service_tl); return UCC_ERR_NOT_FOUND; }
    stub = (ucc_team_t *)calloc(1, sizeof(*stub));
    if (!stub) return UCC_ERR_NO_MEMORY;
    stub->contexts = (ucc_context_t **)malloc(sizeof(void *)); stub->contexts[0] = ctx; stub->num_contexts = 1;
    stub->size = ctx->addr_storage.size; stub->rank = ctx->rank; stub->id = 0;
    stub->ctx_map = ucc_ep_map_create_full(stub->size);
    stub->state = UCC_TEAM_ACTIVE;
    memset(&bp, 0, sizeof(bp));
    bp.params.mask = UCC_TEAM_PARAM_FIELD_EP | UCC_TEAM_PARAM_FIELD_EP_RANGE;
    bp.params.ep = ctx->rank; bp.params.ep_range = UCC_COLLECTIVE_EP_RANGE_CONTIG;
    bp.scope = UCC_CL_LAST; bp.scope_id = 0; bp.rank = ctx->rank; bp.size = stub->size; bp.id = 0; bp.team = stub;
    bp.map = ucc_ep_map_create_full(stub->size);
    st = UCC_TL_CTX_IFACE(ctx->service_ctx)->team.create_post(&ctx->service_ctx->super, &bp, &b_team);
    if (st != UCC_OK) { ucc_warn("tl %s failed to create service team", cfg->service_tl); goto err; }
    do {
        st = UCC_TL_CTX_IFACE(ctx->service_ctx)->team.create_test(b_team);
        ucc_context_progress(ctx);
    } while (st == UCC_INPROGRESS);
    if (st < 0) { ucc_warn("failed to create service tl team"); goto err; }
    ctx->service_team = ucc_derived_of(b_team, ucc_tl_team_t);
    return UCC_OK;
err:
    free(stub->contexts); free(stub);
    ucc_tl_context_put(ctx->service_ctx); ctx->service_ctx = NULL;
    return st;
}

ucc_status_t ucc_context_create_proc_info(ucc_lib_h lib, const ucc_context_params_t *params, const ucc_context_config_h config,
                                          ucc_context_h *context, ucc_proc_info_t *proc_info)
{
    static uint32_t ctx_seq_num = 0;
    ucc_base_context_params_t bp;
    ucc_context_t *ctx;
    ucc_status_t st;
    uint64_t topo_required = 0, svc_required = 0;
    if (!lib || !params || !config || !context) return UCC_ERR_INVALID_PARAM;
    ctx = (ucc_context_t *)calloc(1, sizeof(*ctx));
    if (!ctx) return UCC_ERR_NO_MEMORY;
    ctx->lib = lib; ctx->params = *params; ctx->thread_mode = lib->attr.thread_mode;
    ctx->proc_info = *proc_info; ctx->host_info = ucc_local_host; ctx->host_info.host_hash = proc_info->host_hash;
    ctx->id.pi = *proc_info; ctx->id.seq_num = ucc_atomic_fadd32(&ctx_seq_num, 1);
    ctx->throttle_progress = (int)config->throttle_progress;
    ucc_spinlock_init(&ctx->progress_lock);
    memset(&bp, 0, sizeof(bp));
    bp.params = *params; bp.context = ctx; bp.thread_mode = ctx->thread_mode; bp.prefix = lib->full_prefix;
    bp.estimated_num_eps = (int)config->estimated_num_eps; bp.estimated_num_ppn = (int)config->estimated_num_ppn;
    if ((params->mask & UCC_CONTEXT_PARAM_FIELD_OOB) && params->oob.n_oob_eps > 0) {
        ctx->rank = params->oob.oob_ep;
        if (!bp.estimated_num_eps) bp.estimated_num_eps = (int)params->oob.n_oob_eps;
    }
    if (getenv("UCC_DEBUGGER_WAIT")) { /* spin-for-gdb (reference utils/ucc_debug.h) */
        int target = atoi(getenv("UCC_DEBUGGER_WAIT"));
        volatile int wait = ((int)ctx->rank == target);
        if (wait) ucc_print("rank %d pid %d waits for debugger: set var wait=0", target, (int)getpid());
        while (wait) usleep(1000);
    }
    st = ucc_progress_queue_init(&ctx->pq, ctx->thread_mode, config->lock_free_progress_q);
    if (st != UCC_OK) goto err_free;
    st = create_tl_contexts(ctx, config, &bp);
    if (st != UCC_OK) goto err_pq;
    ctx->cl_ctx = (ucc_cl_context_t **)calloc((size_t)config->n_cl_cfg + 1, sizeof(void *));
    for (int i = 0; i < config->n_cl_cfg; i++) {
        ucc_cl_lib_t *cl_lib = config->cl_cfgs[i]->cl_lib;
        ucc_base_context_t *b_ctx = NULL;
        ucc_cl_lib_attr_t lattr; /* CL get_attr fills the CL-sized structure */
        st = cl_lib->iface->context.create(&bp, &config->cl_cfgs[i]->super, &b_ctx);
        if (st != UCC_OK) {
            if (lib->specific_cls_requested) { ucc_error("failed to create cl context for %s", cl_lib->iface->super.name); goto err_cl; }
            ucc_debug("failed to create cl context for %s, skipping", cl_lib->iface->super.name);
            continue;
        }
        memset(&lattr, 0, sizeof(lattr));
        cl_lib->iface->lib.get_attr(&cl_lib->super, &lattr.super);
        ctx->cl_flags |= lattr.super.flags;
        ctx->cl_ctx[ctx->n_cl_ctx++] = ucc_derived_of(b_ctx, ucc_cl_context_t);
    }
    if (ctx->n_cl_ctx == 0) { ucc_error("no CL context created in ucc_context_create"); st = UCC_ERR_NO_MESSAGE; goto err_cl; }
    for (int i = 0; i < ctx->n_tl_ctx; i++) {
        ucc_base_lib_attr_t lattr; memset(&lattr, 0, sizeof(lattr));
        if (UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->lib.get_attr) UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->lib.get_attr(ctx->tl_ctx[i]->super.lib, &lattr);
        svc_required |= lattr.flags;
    }
    ctx->ids_pool_size = config->team_ids_pool_size;
    ctx->ids_pool = (uint64_t *)malloc(sizeof(uint64_t) * ctx->ids_pool_size);
    if (!ctx->ids_pool) { st = UCC_ERR_NO_MEMORY; goto err_cl; }
    memset(ctx->ids_pool, 0xff, sizeof(uint64_t) * ctx->ids_pool_size);
    ctx->ids_pool[0] &= ~1ull; /* id 0 is reserved for the service team */

    if ((params->mask & UCC_CONTEXT_PARAM_FIELD_OOB) && params->oob.n_oob_eps > 1) {
        ucc_oob_coll_t oob = params->oob;
        do { st = ucc_core_addr_exchange(ctx, &oob, &ctx->addr_storage); } while (st == UCC_INPROGRESS);
        if (st != UCC_OK) { ucc_error("failed to exchange addresses during context creation"); goto err_ids; }
        ctx->rank = ctx->addr_storage.rank;
        for (int i = 0; i < ctx->n_cl_ctx; i++) {
            ucc_base_ctx_attr_t a; memset(&a, 0, sizeof(a));
            UCC_CL_CTX_IFACE(ctx->cl_ctx[i])->context.get_attr(&ctx->cl_ctx[i]->super, &a); topo_required |= (uint64_t)a.topo_required;
        }
        for (int i = 0; i < ctx->n_tl_ctx; i++) {
            ucc_base_ctx_attr_t a; memset(&a, 0, sizeof(a));
            UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->context.get_attr(&ctx->tl_ctx[i]->super, &a); topo_required |= (uint64_t)a.topo_required;
        }
        (void)topo_required; /* the context topology is cheap here: always build it */
        st = ucc_context_topo_init(&ctx->addr_storage, &ctx->topo);
        if (st != UCC_OK) { ucc_error("failed to init ctx topo"); goto err_ids; }
        if (config->internal_oob || (ctx->cl_flags & UCC_BASE_LIB_FLAG_CTX_SERVICE_TEAM_REQUIRED) || (svc_required & UCC_BASE_LIB_FLAG_CTX_SERVICE_TEAM_REQUIRED)) {
            st = create_service_team(ctx, config);
            if (st != UCC_OK && ((ctx->cl_flags | svc_required) & UCC_BASE_LIB_FLAG_CTX_SERVICE_TEAM_REQUIRED)) goto err_ids;
        }
    }
    /* epilog: let TLs finish work that needed everybody's address */
    for (int i = 0; i < ctx->n_tl_ctx; i++) {
        ucc_tl_iface_t *iface = UCC_TL_CTX_IFACE(ctx->tl_ctx[i]);
        if (!iface->context.create_epilog) continue;
        st = iface->context.create_epilog(&ctx->tl_ctx[i]->super);
        if (st != UCC_OK) {
            if (ucc_tl_is_required(lib, iface, 1)) { ucc_error("ctx create epilog for %s failed", iface->super.name); goto err_ids; }
            ucc_debug("ctx create epilog for %s failed, removing the TL", iface->super.name);
            iface->context.destroy(&ctx->tl_ctx[i]->super);
            for (int k = i + 1; k < ctx->n_tl_ctx; k++) ctx->tl_ctx[k - 1] = ctx->tl_ctx[k];
            ctx->n_tl_ctx--; i--;
        }
    }
    for (int i = 0; i < ctx->n_cl_ctx; i++) {
        ucc_cl_iface_t *iface = UCC_CL_CTX_IFACE(ctx->cl_ctx[i]);
        if (iface->context.create_epilog && (st = iface->context.create_epilog(&ctx->cl_ctx[i]->super)) != UCC_OK) goto err_ids;
    }
    ucc_debug("created ucc context %p for lib %s", (void *)ctx, lib->full_prefix);
    if (ucc_global_config.warn_unused_env_vars) ucc_config_parser_warn_unused_env_vars_once(); /* every CL/TL table is registered by now */
    *context = ctx;
    return UCC_OK;
err_ids:
    free(ctx->ids_pool);
err_cl:
    for (int i = 0; i < ctx->n_cl_ctx; i++) UCC_CL_CTX_IFACE(ctx->cl_ctx[i])->context.destroy(&ctx->cl_ctx[i]->super);
    free(ctx->cl_ctx);
    for (int i = 0; i < ctx->n_tl_ctx; i++) UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->context.destroy(&ctx->tl_ctx[i]->super);
    free(ctx->tl_ctx); ucc_config_names_array_free(&ctx->all_tls);
err_pq:
    ucc_progress_queue_finalize(ctx->pq);
err_free:
    ucc_addr_storage_free(&ctx->addr_storage); free(ctx->attr.ctx_addr); free(ctx);
    return st;
}

UCC_EXPORT ucc_status_t ucc_context_create(ucc_lib_h lib, const ucc_context_params_t *params, const ucc_context_config_h config,
                                           ucc_context_h *context)
{ return ucc_context_create_proc_info(lib, params, config, context, &ucc_local_proc); }

UCC_EXPORT ucc_status_t ucc_context_destroy(ucc_context_t *ctx)
{
    if (!ctx) return UCC_ERR_INVALID_PARAM;
    if (ctx->service_team) {
        ucc_team_t *stub = ctx->service_team->super.params.team;
        ucc_status_t st;
        do { st = UCC_TL_CTX_IFACE(ctx->service_ctx)->team.destroy(&ctx->service_team->super); ucc_context_progress(ctx); } while (st == UCC_INPROGRESS);
        if (st != UCC_OK) ucc_error("failed to destroy ctx service team");
        ucc_tl_context_put(ctx->service_ctx);
        free(stub->contexts); free(stub);
    }
    for (int i = 0; i < ctx->n_cl_ctx; i++) UCC_CL_CTX_IFACE(ctx->cl_ctx[i])->context.destroy(&ctx->cl_ctx[i]->super);
    for (int i = 0; i < ctx->n_tl_ctx; i++) {
        if (ctx->tl_ctx[i]->ref_count != 0) ucc_warn("tl ctx %s is still in use", UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->super.name);
        UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->context.destroy(&ctx->tl_ctx[i]->super);
    }
    if (ctx->topo) ucc_context_topo_cleanup(ctx->topo);
    ucc_progress_queue_finalize(ctx->pq);
    ucc_config_names_array_free(&ctx->all_tls);
    ucc_addr_storage_free(&ctx->addr_storage);
    free(ctx->progress_fns); free(ctx->ids_pool); free(ctx->attr.ctx_addr); free(ctx->cl_ctx); free(ctx->tl_ctx); free(ctx);
    return UCC_OK;
}

/* ------------------------------------------------------------------ */
/* progress                                                            */
/* ------------------------------------------------------------------ */
ucc_status_t ucc_context_progress_register(ucc_context_t *ctx, ucc_context_progress_fn_t fn, void *arg)
{
    ucc_spin_lock(&ctx->progress_lock);
    ctx->progress_fns = (ucc_context_progress_entry_t *)realloc(ctx->progress_fns, sizeof(ucc_context_progress_entry_t) * (size_t)(ctx->n_progress_fns + 1));
    ctx->progress_fns[ctx->n_progress_fns].fn = fn; ctx->progress_fns[ctx->n_progress_fns].arg = arg; ctx->n_progress_fns++;
    ucc_spin_unlock(&ctx->progress_lock);
    return UCC_OK;
}
void ucc_context_progress_deregister(ucc_context_t *ctx, ucc_context_progress_fn_t fn, void *arg)
{
    ucc_spin_lock(&ctx->progress_lock);
    for (int i = 0; i < ctx->n_progress_fns; i++)
        if (ctx->progress_fns[i].fn == fn && ctx->progress_fns[i].arg == arg) {
            ctx->progress_fns[i] = ctx->progress_fns[--ctx->n_progress_fns]; break;
        }
    ucc_spin_unlock(&ctx->progress_lock);
}

UCC_EXPORT ucc_status_t ucc_context_progress(ucc_context_h ctx)
{
    int is_empty = ucc_progress_queue_is_empty(ctx->pq);
    int st;
    /* idle contexts only poll the transports every THROTTLE_PROGRESS calls */
    if (ucc_likely(is_empty)) {
        if (ctx->throttle_progress > 0 && ++ctx->progress_counter < ctx->throttle_progress) return UCC_OK;
        ctx->progress_counter = 0;
    }
    for (int i = 0; i < ctx->n_progress_fns; i++) ctx->progress_fns[i].fn(ctx->progress_fns[i].arg);
    if (is_empty) return UCC_OK;
    st = ucc_progress_queue(ctx->pq);
    return st < 0 ? (ucc_status_t)st : UCC_OK;
}

/* ------------------------------------------------------------------ */
/* attr                                                                */
/* ------------------------------------------------------------------ */
UCC_EXPORT ucc_status_t ucc_context_get_attr(ucc_context_t *ctx, ucc_context_attr_t *attr)
{
    ucc_status_t st = UCC_OK;
    if (!ctx || !attr) return UCC_ERR_INVALID_PARAM;
    if (attr->mask & (UCC_CONTEXT_ATTR_FIELD_CTX_ADDR | UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN)) {
        if (!ctx->attr.ctx_addr) { st = ucc_context_pack_addr(ctx, &ctx->attr.ctx_addr, &ctx->attr.ctx_addr_len); if (st != UCC_OK) return st; }
        if (attr->mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR) attr->ctx_addr = ctx->attr.ctx_addr;
        if (attr->mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN) attr->ctx_addr_len = ctx->attr.ctx_addr_len;
    }
    if (attr->mask & UCC_CONTEXT_ATTR_FIELD_TYPE) attr->type = (ctx->params.mask & UCC_CONTEXT_PARAM_FIELD_TYPE) ? ctx->params.type : UCC_CONTEXT_EXCLUSIVE;
    if (attr->mask & UCC_CONTEXT_ATTR_FIELD_SYNC_TYPE) attr->sync_type = (ctx->params.mask & UCC_CONTEXT_PARAM_FIELD_SYNC_TYPE) ? ctx->params.sync_type : UCC_NO_SYNC_COLLECTIVES;
    if (attr->mask & UCC_CONTEXT_ATTR_FIELD_WORK_BUFFER_SIZE) {
        attr->global_work_buffer_size = 0;
        for (int i = 0; i < ctx->n_tl_ctx; i++) {
            ucc_base_ctx_attr_t a; memset(&a, 0, sizeof(a)); a.attr.mask = UCC_CONTEXT_ATTR_FIELD_WORK_BUFFER_SIZE;
            UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->context.get_attr(&ctx->tl_ctx[i]->super, &a);
            if (a.attr.global_work_buffer_size > attr->global_work_buffer_size) attr->global_work_buffer_size = a.attr.global_work_buffer_size;
        }
    }
    return st;
}
