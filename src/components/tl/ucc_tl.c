#include "ucc_tl.h"
#include "core/ucc_context.h"
#include "core/ucc_team.h"
#include "core/ucc_global_opts.h"
#include "utils/ucc_string.h"
#include "utils/ucc_parser.h"
#include "utils/arch/cpu.h"
#include "components/topo/ucc_topo.h"

ucc_config_field_t ucc_tl_lib_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_tl_lib_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_base_lib_config_table)}, {NULL}};
ucc_config_field_t ucc_tl_context_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_tl_context_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_base_ctx_config_table)}, {NULL}};

ucc_status_t ucc_tl_lib_config_read(ucc_tl_iface_t *iface, const char *full_prefix, ucc_tl_lib_config_t **cfg_p)
{
    ucc_tl_lib_config_t *cfg = (ucc_tl_lib_config_t *)calloc(1, iface->tl_lib_config.size);
    ucc_status_t st;
    if (!cfg) return UCC_ERR_NO_MEMORY;
    st = ucc_config_parser_fill_opts(cfg, &iface->tl_lib_config, full_prefix, 0);
    if (st != UCC_OK) { free(cfg); return st; }
    cfg->iface = iface; cfg->super.cfg_entry = &iface->tl_lib_config;
    *cfg_p = cfg;
    return UCC_OK;
}
ucc_status_t ucc_tl_context_config_read(ucc_tl_lib_t *tl_lib, const char *full_prefix, ucc_tl_context_config_t **cfg_p)
{
    ucc_tl_iface_t *iface = tl_lib->iface;
    ucc_tl_context_config_t *cfg = (ucc_tl_context_config_t *)calloc(1, iface->tl_context_config.size);
    ucc_status_t st;
    if (!cfg) return UCC_ERR_NO_MEMORY;
    st = ucc_config_parser_fill_opts(cfg, &iface->tl_context_config, full_prefix, 0);
    if (st != UCC_OK) { free(cfg); return st; }
    cfg->tl_lib = tl_lib; cfg->super.lib = &tl_lib->super; cfg->super.cfg_entry = &iface->tl_context_config;
    *cfg_p = cfg;
    return UCC_OK;
}
void ucc_tl_lib_config_release(ucc_tl_iface_t *iface, ucc_tl_lib_config_t *cfg)
{ if (!cfg) return; ucc_config_parser_release_opts(cfg, iface->tl_lib_config.table); free(cfg); }
void ucc_tl_context_config_release(ucc_tl_iface_t *iface, ucc_tl_context_config_t *cfg)
{ if (!cfg) return; ucc_config_parser_release_opts(cfg, iface->tl_context_config.table); free(cfg); }

void ucc_tl_lib_init_base(ucc_tl_lib_t *lib, ucc_tl_iface_t *iface, const ucc_tl_lib_config_t *cfg)
{
    lib->iface = iface;
    lib->super.log_component = cfg->super.log_component;
    snprintf(lib->super.log_component.name, sizeof(lib->super.log_component.name), "TL_%s", iface->super.name);
    for (char *c = lib->super.log_component.name; *c; c++) if (*c >= 'a' && *c <= 'z') *c -= 32;
    lib->super.use_tuning = cfg->super.use_tuning;
    lib->super.min_team_size = cfg->super.min_team_size;
}

ucc_status_t ucc_tl_context_get(ucc_context_t *ctx, const char *name, ucc_tl_context_t **tl_context)
{
    for (int i = 0; i < ctx->n_tl_ctx; i++) {
        ucc_tl_lib_t *lib = ucc_derived_of(ctx->tl_ctx[i]->super.lib, ucc_tl_lib_t);
        if (!strcmp(name, lib->iface->super.name)) { ctx->tl_ctx[i]->ref_count++; *tl_context = ctx->tl_ctx[i]; return UCC_OK; }
    }
    return UCC_ERR_NOT_FOUND;
}
ucc_status_t ucc_tl_context_put(ucc_tl_context_t *tl_context) { tl_context->ref_count--; return UCC_OK; }

ucc_status_t ucc_tl_team_size_check(ucc_tl_iface_t *iface, ucc_tl_lib_t *lib, ucc_rank_t size)
{
    ucc_base_lib_properties_t p = {0, 1, UCC_RANK_MAX};
    ucc_rank_t min;
    if (iface->lib.get_properties) iface->lib.get_properties(&p);
    min = (lib->super.min_team_size == UCC_UUNITS_AUTO) ? p.default_team_size : lib->super.min_team_size;
    if (min < p.min_team_size) min = p.min_team_size;
    if (size < min || size > p.max_team_size) return UCC_ERR_NOT_SUPPORTED;
    return UCC_OK;
}

ucc_status_t ucc_tl_apply_tune(ucc_tl_team_t *team, ucc_coll_score_t *score, const ucc_coll_score_team_info_t *info,
                               const char *default_select_str, const char *user_tune)
{
    ucc_base_lib_t *lib = team->super.context->lib;
    ucc_status_t st;
    if (default_select_str && default_select_str[0] && lib->use_tuning) {
        st = ucc_coll_score_update_from_str(default_select_str, info, &team->super, score);
        if (st != UCC_OK) { tl_error(lib, "failed to apply default selection string: %s", default_select_str); return st; }
    }
    if (user_tune && user_tune[0]) {
        st = ucc_coll_score_update_from_str(user_tune, info, &team->super, score);
        /* a bad user string is reported but does not disable the component (reference tl_cuda_team.c:620-627) */
        if (st == UCC_ERR_INVALID_PARAM || st == UCC_ERR_NOT_SUPPORTED) tl_warn(lib, "ignoring invalid TUNE string \"%s\"", user_tune);
        else if (st != UCC_OK) return st;
    }
    /* [section]s of the configuration file that are predicated on team facts (team_size / ppn / nnodes / sockets / cpu) may carry
     * a <TL>_TUNE of their own: it is applied last, per team (reference: ucc_add_team_sections, utils/ucc_parser.c:612-655,
     * called from tl_ucp_team.c:82-88) */
    if (ucc_global_config.file_cfg && team->super.params.team && team->super.params.scope != UCC_CL_LAST && lib->use_tuning) {
        ucc_team_t *core = team->super.params.team;
        ucc_tl_iface_t *iface = UCC_TL_TEAM_IFACE(team);
        ucc_file_section_filter_t f;
        char key[128]; const char *sec_tune;
        memset(&f, 0, sizeof(f));
        f.vendor = ucc_cpu_vendor_string(ucc_arch_get_cpu_vendor()); f.model = ucc_arch_get_cpu_model_string();
        f.team_size = (unsigned)team->super.params.size;
        if (core->topo && team->super.params.size == core->size) {
            f.ppn = (unsigned)ucc_topo_max_ppn(core->topo); f.nnodes = (unsigned)ucc_topo_nnodes(core->topo); f.sock = (unsigned)core->topo->max_socket_size;
        } else if (core->size == team->super.params.size && core->contexts[0]->addr_storage.flags) { f.ppn = f.team_size; f.nnodes = 1; }
        snprintf(key, sizeof(key), "UCC_%sTUNE", iface->tl_context_config.prefix ? iface->tl_context_config.prefix : "");
        sec_tune = ucc_file_config_lookup(ucc_global_config.file_cfg, key, &f);
        if (sec_tune && sec_tune[0] && (!user_tune || strcmp(sec_tune, user_tune))) {
            st = ucc_coll_score_update_from_str(sec_tune, info, &team->super, score);
            if (st == UCC_ERR_INVALID_PARAM || st == UCC_ERR_NOT_SUPPORTED) tl_warn(lib, "ignoring invalid TUNE string \"%s\" of a configuration file section", sec_tune);
            else if (st != UCC_OK) return st;
            else tl_debug(lib, "team size %u: applied section tuning \"%s\"", f.team_size, sec_tune);
        }
    }
    return UCC_OK;
}

ucc_status_t ucc_team_multiple_req_alloc(ucc_team_multiple_req_t **req, int n_teams)
{
    ucc_team_multiple_req_t *r = (ucc_team_multiple_req_t *)calloc(1, sizeof(*r) + sizeof(ucc_team_multiple_req_descr_t) * (size_t)(n_teams > 0 ? n_teams - 1 : 0));
    if (!r) return UCC_ERR_NO_MEMORY;
    r->n_teams = n_teams; r->last = -1;
    *req = r;
    return UCC_OK;
}
void ucc_team_multiple_req_free(ucc_team_multiple_req_t *req) { free(req); }

/* one team at a time: post i, test i until done, move on; a failing TL just records its status */
ucc_status_t ucc_tl_team_create_multiple(ucc_team_multiple_req_t *req)
{
    ucc_base_team_t *b_team;
    ucc_status_t st;
    int i = req->last;
    if (i >= 0 && req->descs[i].status == UCC_INPROGRESS) {
        st = UCC_TL_CTX_IFACE(req->descs[i].ctx)->team.create_test(&req->descs[i].team->super);
        if (st == UCC_INPROGRESS) return UCC_INPROGRESS;
        req->descs[i].status = st;
        if (st != UCC_OK) { req->descs[i].team = NULL; } /* create_test is responsible for releasing a failed team */
    }
    for (i = req->last + 1; i < req->n_teams; i++) {
        req->last = i;
        st = UCC_TL_CTX_IFACE(req->descs[i].ctx)->team.create_post(&req->descs[i].ctx->super, &req->descs[i].param, &b_team);
        if (st != UCC_OK) { req->descs[i].status = st; req->descs[i].team = NULL; continue; }
        req->descs[i].team = ucc_derived_of(b_team, ucc_tl_team_t);
        req->descs[i].status = UCC_INPROGRESS;
        return UCC_INPROGRESS;
    }
    return UCC_OK;
}

ucc_status_t ucc_tl_team_destroy_multiple(ucc_team_multiple_req_t *req)
{
    ucc_status_t st_all = UCC_OK;
    for (int i = 0; i < req->n_teams; i++) {
        ucc_status_t st;
        if (!req->descs[i].team) continue;
        st = UCC_TL_TEAM_IFACE(req->descs[i].team)->team.destroy(&req->descs[i].team->super);
        if (st == UCC_INPROGRESS) st_all = UCC_INPROGRESS;
        else { req->descs[i].team = NULL; if (st != UCC_OK && st_all == UCC_OK) st_all = st; }
    }
    return st_all;
}
