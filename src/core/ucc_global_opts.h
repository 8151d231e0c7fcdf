/* Process-wide options parsed once from UCC_* (role of reference core/ucc_global_opts.c:34-121). */
#ifndef UCC_GLOBAL_OPTS_H_
#define UCC_GLOBAL_OPTS_H_
#include "utils/ucc_log.h"
#include "utils/ucc_parser.h"
#include "utils/ucc_component.h"

typedef enum { UCC_PROFILE_MODE_NONE = 0, UCC_PROFILE_MODE_ACCUM = 1, UCC_PROFILE_MODE_LOG = 2 } ucc_profile_mode_t;

typedef struct ucc_global_config {
    ucc_log_component_config_t log_component;   /* LOG_LEVEL */
    ucc_log_level_t            coll_trace_level;/* COLL_TRACE */
    ucc_log_component_config_t coll_trace;
    int                        warn_unused_env_vars;
    char                      *component_path;  /* resolved <libdir>/ucc */
    char                      *install_path;
    int                        initialized;
    unsigned                   profile_mode;    /* bitmask parsed from PROFILE_MODE */
    char                      *profile_mode_str;
    char                      *profile_file;
    size_t                     profile_log_size;
    char                      *cfg_filename;
    ucc_file_config_t         *file_cfg;
    char                      *log_file;
    size_t                     log_file_size;
    unsigned                   log_file_rotate;
    size_t                     log_buffer_size;
    size_t                     log_data_size;
    int                        log_print_enable;
    ucc_log_level_t            log_level_trigger;
    int                        check_asymmetric_dt;
    char                      *module_dir;      /* MODULE_DIR override for the dlopen loader */
    ucc_component_framework_t  cl_framework, tl_framework, mc_framework, ec_framework, sysinfo_framework;
} ucc_global_config_t;

extern ucc_global_config_t ucc_global_config;
extern ucc_config_field_t  ucc_global_config_table[];
extern const char         *ucc_log_level_cfg_names[];
#endif
