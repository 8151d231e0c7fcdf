/* One-time process-wide initialisation (role of reference core/ucc_constructor.c:100-218). */
#include "ucc_global_opts.h"
#include "ucc_lib.h"
#include "utils/ucc_sys.h"
#include "utils/ucc_proc_info.h"
#include "utils/profile/ucc_profile.h"
#include <pthread.h>
#include <unistd.h>

static pthread_mutex_t ucc_constructor_mutex = PTHREAD_MUTEX_INITIALIZER;

/* $UCC_CONFIG_FILE -> $HOME/ucc.conf -> <prefix>/share/ucc.conf */
static void load_config_file(void)
{
    char path[1024];
    const char *home;
    if (ucc_global_config.cfg_filename && ucc_global_config.cfg_filename[0]) {
        if (ucc_parse_file_config(ucc_global_config.cfg_filename, &ucc_global_config.file_cfg) != UCC_OK)
            ucc_warn("failed to read config file %s", ucc_global_config.cfg_filename);
        return;
    }
    home = getenv("HOME");
    if (home) {
        snprintf(path, sizeof(path), "%s/ucc.conf", home);
        if (ucc_parse_file_config(path, &ucc_global_config.file_cfg) == UCC_OK) return;
    }
    if (ucc_global_config.install_path) {
        snprintf(path, sizeof(path), "%s/share/ucc.conf", ucc_global_config.install_path);
        if (ucc_parse_file_config(path, &ucc_global_config.file_cfg) == UCC_OK) return;
    }
}

static void init_lib_paths(void)
{
    char buf[4200];
    const char *dir = ucc_sys_dirname_of_lib(); /* .../lib */
    snprintf(buf, sizeof(buf), "%s/ucc", dir);
    ucc_global_config.component_path = strdup(buf);
    snprintf(buf, sizeof(buf), "%s/..", dir);
    ucc_global_config.install_path = strdup(buf);
}

ucc_status_t ucc_constructor(void)
{
    ucc_status_t st = UCC_OK;
    pthread_mutex_lock(&ucc_constructor_mutex);
    if (ucc_global_config.initialized) goto out;
    { extern void ucc_debug_install_handlers(void); ucc_debug_install_handlers(); }
    st = ucc_config_parser_fill_opts_table(&ucc_global_config, ucc_global_config_table, "UCC_", "", 1);
    if (st != UCC_OK) { fprintf(stderr, "ucc: failed to parse global options\n"); goto out; }
    ucc_global_log_component.log_level = ucc_global_config.log_component.log_level;
    ucc_global_config.coll_trace.log_level = ucc_global_config.coll_trace_level;
    ucc_log_init();
    init_lib_paths();
    load_config_file();
    if (ucc_global_config.file_cfg) {
        /* re-read the global table now that file values are available (env still wins) */
        ucc_config_parser_release_opts(&ucc_global_config, ucc_global_config_table);
        ucc_config_parser_fill_opts_table(&ucc_global_config, ucc_global_config_table, "UCC_", "", 1);
        ucc_global_log_component.log_level = ucc_global_config.log_component.log_level;
        ucc_global_config.coll_trace.log_level = ucc_global_config.coll_trace_level;
    }
    if (ucc_components_load("cl", &ucc_global_config.cl_framework) != UCC_OK) { ucc_error("no CL components found"); st = UCC_ERR_NOT_FOUND; goto out; }
    if (ucc_component_check_scores_uniq(&ucc_global_config.cl_framework) != UCC_OK) { st = UCC_ERR_INVALID_PARAM; goto out; }
    if (ucc_components_load("tl", &ucc_global_config.tl_framework) != UCC_OK) { ucc_error("no TL components found"); st = UCC_ERR_NOT_FOUND; goto out; }
    if (ucc_component_check_scores_uniq(&ucc_global_config.tl_framework) != UCC_OK) { st = UCC_ERR_INVALID_PARAM; goto out; }
    if (ucc_components_load("mc", &ucc_global_config.mc_framework) != UCC_OK) { ucc_error("no MC components found"); st = UCC_ERR_NOT_FOUND; goto out; }
    if (ucc_components_load("ec", &ucc_global_config.ec_framework) != UCC_OK) { ucc_error("no EC components found"); st = UCC_ERR_NOT_FOUND; goto out; }
    ucc_components_load("sysinfo", &ucc_global_config.sysinfo_framework); /* optional */
    if (ucc_local_proc_info_init() != UCC_OK) { st = UCC_ERR_NO_MESSAGE; goto out; }
    ucc_profile_init(ucc_global_config.profile_mode_str, ucc_global_config.profile_file, ucc_global_config.profile_log_size);
    ucc_global_config.initialized = 1;
out:
    pthread_mutex_unlock(&ucc_constructor_mutex);
    return st;
}

static void UCC_DTOR ucc_destructor(void)
{
    if (!ucc_global_config.initialized) return;
    ucc_profile_cleanup();
    ucc_log_cleanup();
}
