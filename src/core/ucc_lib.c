/* Library object: component selection and life cycle.
 * Behaviour follows reference core/ucc_lib.c:64-518 (CLS list, TLs pulled in by
 * CLs, thread-mode negotiation, graceful skip of components that fail). */
#include "ucc_lib.h"
#include "ucc_global_opts.h"
#include "components/mc/ucc_mc.h"
#include "components/ec/ucc_ec.h"
#include "components/topo/ucc_sysinfo.h"
#include "utils/ucc_sys.h"
#include "utils/ucc_string.h"
#include "schedule/ucc_schedule.h"

ucc_status_t ucc_constructor(void);
ucc_host_info_t ucc_local_host;

static ucc_config_field_t ucc_lib_config_table[] = {
    {"CLS", "basic", "Comma separated list of CL components to be used (\"all\" selects every available CL)",
     ucc_offsetof(ucc_lib_config_t, cls), UCC_CONFIG_TYPE_STRING_ARRAY},
    {NULL}};
UCC_CONFIG_REGISTER_TABLE(ucc_lib_config_table, "UCC", "", ucc_lib_config_t)

UCC_EXPORT ucc_status_t ucc_lib_config_read(const char *env_prefix, const char *filename, ucc_lib_config_h *config_p)
{
    ucc_lib_config_t *config;
    ucc_status_t st;
    char prefix[128];
    if (filename) { ucc_error("read from file is not implemented"); return UCC_ERR_NOT_IMPLEMENTED; }
    st = ucc_constructor();
    if (st != UCC_OK) return st;
    config = (ucc_lib_config_t *)calloc(1, sizeof(*config));
    if (!config) return UCC_ERR_NO_MEMORY;
    if (env_prefix && env_prefix[0]) snprintf(prefix, sizeof(prefix), "%s_UCC_", env_prefix);
    else snprintf(prefix, sizeof(prefix), "UCC_");
    config->full_prefix = strdup(prefix);
    st = ucc_config_parser_fill_opts_table(config, ucc_lib_config_table, config->full_prefix, "", 0);
    if (st != UCC_OK) { free(config->full_prefix); free(config); return st; }
    *config_p = config;
    return UCC_OK;
}
UCC_EXPORT void ucc_lib_config_release(ucc_lib_config_h config)
{ if (!config) return; ucc_config_parser_release_opts(config, ucc_lib_config_table); free(config->full_prefix); free(config); }
UCC_EXPORT void ucc_lib_config_print(const ucc_lib_config_h config, FILE *stream, const char *title, ucc_config_print_flags_t flags)
{ ucc_config_parser_print_opts(stream, title, config, ucc_lib_config_table, "", config->full_prefix, flags); }
UCC_EXPORT ucc_status_t ucc_lib_config_modify(ucc_lib_config_h config, const char *name, const char *value)
{ return ucc_config_parser_set_value(config, ucc_lib_config_table, name, value); }

UCC_EXPORT void ucc_get_version(unsigned *major, unsigned *minor, unsigned *release)
{ *major = UCC_API_MAJOR; *minor = UCC_API_MINOR; *release = 0; }
UCC_EXPORT const char *ucc_get_version_string(void) { return UCC_VERSION_STRING; }

static void sysinfo_collect(void)
{
    ucc_component_framework_t *fw = &ucc_global_config.sysinfo_framework;
    memset(&ucc_local_host, 0, sizeof(ucc_local_host));
    ucc_local_host.host_hash = ucc_local_proc.host_hash;
    ucc_local_host.current_gpu = -1;
    for (int i = 0; i < fw->n_components; i++) {
        ucc_sysinfo_base_t *s = ucc_derived_of(fw->components[i], ucc_sysinfo_base_t);
        if (s->init && s->init() != UCC_OK) continue;
        if (s->get_host_info) s->get_host_info(&ucc_local_host);
    }
}

/* ---- CL libs ---- */
static ucc_status_t cl_lib_open(ucc_lib_info_t *lib, ucc_cl_iface_t *iface, const ucc_lib_params_t *params, int forced,
                                ucc_thread_mode_t *supported_tm, uint64_t *coll_types)
{
    ucc_cl_lib_config_t *cfg = NULL;
    ucc_base_lib_t *b_lib = NULL;
    ucc_base_lib_params_t bp;
    ucc_cl_lib_attr_t attr;
    ucc_status_t st = ucc_cl_lib_config_read(iface, lib->full_prefix, &cfg);
    if (st != UCC_OK) { ucc_error("failed to read CL \"%s\" lib configuration", iface->super.name); return st; }
    bp.params = *params; bp.full_prefix = lib->full_prefix;
    st = iface->lib.init(&bp, &cfg->super, &b_lib);
    ucc_cl_lib_config_release(iface, cfg);
    if (st != UCC_OK) {
        if (forced) ucc_error("lib_init failed for component: %s", iface->super.name);
        else ucc_info("lib_init failed for component: %s, skipping", iface->super.name);
        return st;
    }
    memset(&attr, 0, sizeof(attr));
    attr.super.attr.mask = UCC_LIB_ATTR_FIELD_THREAD_MODE | UCC_LIB_ATTR_FIELD_COLL_TYPES;
    st = iface->lib.get_attr(b_lib, &attr.super);
    if (st != UCC_OK) { iface->lib.finalize(b_lib); return st; }
    if (attr.super.attr.thread_mode < params->thread_mode && (params->mask & UCC_LIB_PARAM_FIELD_THREAD_MODE)) {
        ucc_info("CL %s supports thread mode %s < requested %s, skipping", iface->super.name,
                 ucc_thread_mode_str(attr.super.attr.thread_mode), ucc_thread_mode_str(params->thread_mode));
        iface->lib.finalize(b_lib);
        return UCC_ERR_NOT_SUPPORTED;
    }
    if (attr.super.attr.thread_mode < *supported_tm) *supported_tm = attr.super.attr.thread_mode;
    *coll_types |= attr.super.attr.coll_types;
    lib->cl_attrs[lib->n_cl_libs_opened] = attr;
    lib->cl_libs[lib->n_cl_libs_opened++] = ucc_derived_of(b_lib, ucc_cl_lib_t);
    ucc_info("lib_prefix \"%s\": initialized component \"%s\" score %u", lib->full_prefix, iface->super.name, iface->super.score);
    return UCC_OK;
}

static ucc_status_t ucc_cl_lib_init(const ucc_lib_params_t *user_params, const ucc_lib_config_t *config, ucc_lib_info_t *lib)
{
    ucc_component_framework_t *fw = &ucc_global_config.cl_framework;
    ucc_lib_params_t params = *user_params;
    ucc_thread_mode_t supported_tm = UCC_THREAD_MULTIPLE;
    uint64_t coll_types = 0;
    int n_cls = (int)config->cls.count, all = 0;
    if (!(params.mask & UCC_LIB_PARAM_FIELD_THREAD_MODE)) { params.mask |= UCC_LIB_PARAM_FIELD_THREAD_MODE; params.thread_mode = UCC_THREAD_SINGLE; }
    if (n_cls == 0) { ucc_error("empty UCC_CLS list"); return UCC_ERR_INVALID_PARAM; }
    for (int i = 0; i < n_cls; i++) if (!strcasecmp(config->cls.names[i], "all")) all = 1;
    lib->specific_cls_requested = !all;
    lib->cl_libs  = (ucc_cl_lib_t **)calloc((size_t)fw->n_components, sizeof(void *));
    lib->cl_attrs = (ucc_cl_lib_attr_t *)calloc((size_t)fw->n_components, sizeof(ucc_cl_lib_attr_t));
    if (!lib->cl_libs || !lib->cl_attrs) return UCC_ERR_NO_MEMORY;
    lib->n_cl_libs_opened = 0;
    if (all) {
        for (int i = 0; i < fw->n_components; i++)
            cl_lib_open(lib, ucc_derived_of(fw->components[i], ucc_cl_iface_t), &params, 0, &supported_tm, &coll_types);
    } else {
        for (int i = 0; i < n_cls; i++) {
            ucc_cl_iface_t *iface = ucc_derived_of(ucc_get_component(fw, config->cls.names[i]), ucc_cl_iface_t);
            ucc_status_t st;
            if (!iface) { ucc_error("required CL component \"%s\" is not available", config->cls.names[i]); return UCC_ERR_NOT_FOUND; }
            st = cl_lib_open(lib, iface, &params, 1, &supported_tm, &coll_types);
            if (st != UCC_OK) return st;
        }
    }
    if (lib->n_cl_libs_opened == 0) { ucc_error("lib_init failed: no CLs left after initialization"); return UCC_ERR_NO_MESSAGE; }
    lib->attr.thread_mode = ucc_min(supported_tm, params.thread_mode);
    lib->attr.coll_types = coll_types;
    return UCC_OK;
}

int ucc_tl_is_required(ucc_lib_info_t *lib, ucc_tl_iface_t *tl_iface, int forced)
{
    for (int i = 0; i < lib->n_cl_libs_opened; i++) {
        ucc_cl_lib_t *cl = lib->cl_libs[i];
        if (forced && !cl->tls_forced) continue;
        if (ucc_config_names_search(&cl->tls.array, tl_iface->super.name) >= 0) return 1;
    }
    return 0;
}

static ucc_status_t ucc_tl_lib_init(const ucc_lib_params_t *user_params, ucc_lib_info_t *lib)
{
    ucc_component_framework_t *fw = &ucc_global_config.tl_framework;
    ucc_lib_params_t params = *user_params;
    params.mask |= UCC_LIB_PARAM_FIELD_THREAD_MODE; params.thread_mode = lib->attr.thread_mode;
    lib->tl_libs = (ucc_tl_lib_t **)calloc((size_t)fw->n_components, sizeof(void *));
    if (!lib->tl_libs) return UCC_ERR_NO_MEMORY;
    lib->n_tl_libs_opened = 0;
    for (int i = 0; i < fw->n_components; i++) {
        ucc_tl_iface_t *iface = ucc_derived_of(fw->components[i], ucc_tl_iface_t);
        ucc_tl_lib_config_t *cfg = NULL;
        ucc_base_lib_t *b_lib = NULL;
        ucc_base_lib_params_t bp;
        ucc_status_t st;
        if (!ucc_tl_is_required(lib, iface, 0)) continue;
        st = ucc_tl_lib_config_read(iface, lib->full_prefix, &cfg);
        if (st != UCC_OK) { ucc_warn("failed to read TL \"%s\" lib configuration", iface->super.name); continue; }
        bp.params = params; bp.full_prefix = lib->full_prefix;
        st = iface->lib.init(&bp, &cfg->super, &b_lib);
        ucc_tl_lib_config_release(iface, cfg);
        if (st != UCC_OK) {
            if (ucc_tl_is_required(lib, iface, 1)) ucc_warn("lib_init failed for TL \"%s\" that was explicitly requested", iface->super.name);
            else ucc_debug("lib_init failed for TL \"%s\", skipping", iface->super.name);
            continue;
        }
        lib->tl_libs[lib->n_tl_libs_opened++] = ucc_derived_of(b_lib, ucc_tl_lib_t);
        ucc_info("lib_prefix \"%s\": initialized tl \"%s\" score %u", lib->full_prefix, iface->super.name, iface->super.score);
    }
    if (lib->n_tl_libs_opened == 0) { ucc_error("no TL could be initialised"); return UCC_ERR_NOT_FOUND; }
    return UCC_OK;
}

UCC_EXPORT ucc_status_t ucc_init_version(unsigned api_major_version, unsigned api_minor_version, const ucc_lib_params_t *params,
                                         const ucc_lib_config_h config, ucc_lib_h *lib_p)
{
    unsigned major, minor, release;
    ucc_lib_info_t *lib;
    ucc_status_t st;
    ucc_mc_params_t mc_params; ucc_ec_params_t ec_params;
    if (!params || !config || !lib_p) return UCC_ERR_INVALID_PARAM;
    *lib_p = NULL;
    st = ucc_constructor();
    if (st != UCC_OK) return st;
    mc_params.thread_mode = (params->mask & UCC_LIB_PARAM_FIELD_THREAD_MODE) ? params->thread_mode : UCC_THREAD_SINGLE;
    ec_params.thread_mode = mc_params.thread_mode;
    if ((st = ucc_mc_init(&mc_params)) != UCC_OK) return st;
    if ((st = ucc_ec_init(&ec_params)) != UCC_OK) { ucc_mc_finalize(); return st; }
    sysinfo_collect();
    ucc_get_version(&major, &minor, &release);
    if (api_major_version != major || (api_major_version == major && api_minor_version > minor))
        ucc_warn("UCC version is incompatible, required: %u.%u, actual: %u.%u.%u", api_major_version, api_minor_version, major, minor, release);
    lib = (ucc_lib_info_t *)calloc(1, sizeof(*lib));
    if (!lib) { st = UCC_ERR_NO_MEMORY; goto err; }
    lib->full_prefix = strdup(config->full_prefix);
    st = ucc_cl_lib_init(params, config, lib);
    if (st != UCC_OK) goto err_lib;
    st = ucc_tl_lib_init(params, lib);
    if (st != UCC_OK) goto err_lib;
    lib->attr.mask = UCC_LIB_ATTR_FIELD_THREAD_MODE | UCC_LIB_ATTR_FIELD_COLL_TYPES | UCC_LIB_ATTR_FIELD_REDUCTION_TYPES | UCC_LIB_ATTR_FIELD_SYNC_TYPE;
    lib->attr.reduction_types = (1u << UCC_OP_LAST) - 1;
    lib->attr.sync_type = UCC_NO_SYNC_COLLECTIVES;
    st = ucc_mpool_init(&lib->stub_tasks_mp, 0, sizeof(ucc_coll_task_t), 0, 64, 8, (unsigned)-1, NULL, lib->attr.thread_mode, "stub_tasks");
    if (st != UCC_OK) goto err_lib;
    *lib_p = lib;
    return UCC_OK;
err_lib:
    for (int i = 0; i < lib->n_tl_libs_opened; i++) lib->tl_libs[i]->iface->lib.finalize(&lib->tl_libs[i]->super);
    for (int i = 0; i < lib->n_cl_libs_opened; i++) lib->cl_libs[i]->iface->lib.finalize(&lib->cl_libs[i]->super);
    free(lib->tl_libs); free(lib->cl_libs); free(lib->cl_attrs); free(lib->full_prefix); free(lib);
err:
    ucc_ec_finalize(); ucc_mc_finalize();
    return st;
}

UCC_EXPORT ucc_status_t ucc_lib_get_attr(ucc_lib_h lib, ucc_lib_attr_t *attr)
{
    if (!lib || !attr) return UCC_ERR_INVALID_PARAM;
    if (attr->mask & UCC_LIB_ATTR_FIELD_THREAD_MODE) attr->thread_mode = lib->attr.thread_mode;
    if (attr->mask & UCC_LIB_ATTR_FIELD_COLL_TYPES) attr->coll_types = lib->attr.coll_types;
    if (attr->mask & UCC_LIB_ATTR_FIELD_REDUCTION_TYPES) attr->reduction_types = lib->attr.reduction_types;
    if (attr->mask & UCC_LIB_ATTR_FIELD_SYNC_TYPE) attr->sync_type = lib->attr.sync_type;
    return UCC_OK;
}

UCC_EXPORT ucc_status_t ucc_finalize(ucc_lib_h lib)
{
    ucc_status_t st1, st2;
    if (!lib) return UCC_ERR_INVALID_PARAM;
    ucc_mpool_cleanup(&lib->stub_tasks_mp, 1);
    for (int i = 0; i < lib->n_cl_libs_opened; i++) lib->cl_libs[i]->iface->lib.finalize(&lib->cl_libs[i]->super);
    for (int i = 0; i < lib->n_tl_libs_opened; i++) lib->tl_libs[i]->iface->lib.finalize(&lib->tl_libs[i]->super);
    free(lib->cl_libs); free(lib->tl_libs); free(lib->cl_attrs); free(lib->full_prefix); free(lib);
    st1 = ucc_ec_finalize(); st2 = ucc_mc_finalize();
    return st1 != UCC_OK ? st1 : st2;
}
