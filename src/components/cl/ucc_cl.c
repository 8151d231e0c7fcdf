#include "ucc_cl.h"
#include "core/ucc_global_opts.h"
#include "utils/ucc_string.h"
#include <strings.h>

ucc_config_field_t ucc_cl_lib_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_cl_lib_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_base_lib_config_table)},
    {"TLS", "all", "Comma separated list of TLs to be used by this CL (use \"all\" for every available TL, \"^name\" to exclude)",
     ucc_offsetof(ucc_cl_lib_config_t, tls), UCC_CONFIG_TYPE_ALLOW_LIST},
    {NULL}};
ucc_config_field_t ucc_cl_context_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_cl_context_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_base_ctx_config_table)}, {NULL}};

static const char *cl_names[] = {"basic", "hier", "doca_urom", "all", NULL};
const char *ucc_cl_name(ucc_cl_type_t t) { return t < UCC_CL_LAST ? cl_names[t] : "unknown"; }

ucc_status_t ucc_cl_lib_config_read(ucc_cl_iface_t *iface, const char *full_prefix, ucc_cl_lib_config_t **cfg_p)
{
    ucc_cl_lib_config_t *cfg = (ucc_cl_lib_config_t *)calloc(1, iface->cl_lib_config.size);
    ucc_status_t st;
    if (!cfg) return UCC_ERR_NO_MEMORY;
    st = ucc_config_parser_fill_opts(cfg, &iface->cl_lib_config, full_prefix, 0);
    if (st != UCC_OK) { free(cfg); return st; }
    cfg->iface = iface; cfg->super.cfg_entry = &iface->cl_lib_config;
    *cfg_p = cfg;
    return UCC_OK;
}
ucc_status_t ucc_cl_context_config_read(ucc_cl_lib_t *cl_lib, const char *full_prefix, ucc_cl_context_config_t **cfg_p)
{
    ucc_cl_iface_t *iface = cl_lib->iface;
    ucc_cl_context_config_t *cfg = (ucc_cl_context_config_t *)calloc(1, iface->cl_context_config.size);
    ucc_status_t st;
    if (!cfg) return UCC_ERR_NO_MEMORY;
    st = ucc_config_parser_fill_opts(cfg, &iface->cl_context_config, full_prefix, 0);
    if (st != UCC_OK) { free(cfg); return st; }
    cfg->cl_lib = cl_lib; cfg->super.lib = &cl_lib->super; cfg->super.cfg_entry = &iface->cl_context_config;
    *cfg_p = cfg;
    return UCC_OK;
}
void ucc_cl_lib_config_release(ucc_cl_iface_t *iface, ucc_cl_lib_config_t *cfg)
{ if (!cfg) return; ucc_config_parser_release_opts(cfg, iface->cl_lib_config.table); free(cfg); }
void ucc_cl_context_config_release(ucc_cl_iface_t *iface, ucc_cl_context_config_t *cfg)
{ if (!cfg) return; ucc_config_parser_release_opts(cfg, iface->cl_context_config.table); free(cfg); }

ucc_status_t ucc_cl_lib_init_base(ucc_cl_lib_t *lib, ucc_cl_iface_t *iface, const ucc_cl_lib_config_t *cfg)
{
    lib->iface = iface;
    lib->super.log_component = cfg->super.log_component;
    snprintf(lib->super.log_component.name, sizeof(lib->super.log_component.name), "CL_%s", iface->super.name);
    for (char *c = lib->super.log_component.name; *c; c++) if (*c >= 'a' && *c <= 'z') *c -= 32;
    lib->super.use_tuning = cfg->super.use_tuning;
    lib->super.min_team_size = cfg->super.min_team_size;
    lib->tls_forced = (cfg->tls.mode == UCC_CONFIG_ALLOW_LIST_ALLOW);
    return ucc_config_allow_list_process(&cfg->tls, &ucc_global_config.tl_framework.names, &lib->tls);
}
void ucc_cl_lib_cleanup_base(ucc_cl_lib_t *lib) { ucc_config_names_array_free(&lib->tls.array); }

ucc_status_t ucc_parse_cls_string(const char *cls_str, ucc_cl_type_t **cls_array, int *n_cls)
{
    char **t = ucc_str_split(cls_str, ",");
    unsigned n = ucc_str_split_count(t);
    ucc_cl_type_t *arr = (ucc_cl_type_t *)malloc(sizeof(ucc_cl_type_t) * (n ? n : 1));
    int k = 0;
    for (unsigned i = 0; i < n; i++) {
        int id = ucc_str_find_in_list(t[i], cl_names);
        if (id < 0) { ucc_error("unsupported CL name \"%s\" in UCC_CLS", t[i]); ucc_str_split_free(t); free(arr); return UCC_ERR_INVALID_PARAM; }
        if (id == UCC_CL_ALL) { k = 0; arr[k++] = UCC_CL_ALL; break; }
        arr[k++] = (ucc_cl_type_t)id;
    }
    ucc_str_split_free(t);
    *cls_array = arr; *n_cls = k;
    return UCC_OK;
}
