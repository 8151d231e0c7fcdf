#include "ucc_global_opts.h"

ucc_global_config_t ucc_global_config = {
    .log_component    = {UCC_LOG_LEVEL_WARN, "UCC"},
    .coll_trace       = {UCC_LOG_LEVEL_WARN, "UCC_COLL"},
    .log_file_size    = (size_t)-1,
    .log_level_trigger = UCC_LOG_LEVEL_FATAL,
    .log_print_enable = 1,
};

const char *ucc_log_level_cfg_names[] = {"fatal", "error", "warn", "diag", "info", "debug", "trace", "req",
                                         "data", "async", "func", "poll", NULL};

ucc_config_field_t ucc_global_config_table[] = {
    {"LOG_LEVEL", "warn",
     "UCC logging level. Messages with a level higher or equal to the selected will be printed.\n"
     "Possible values are: fatal, error, warn, diag, info, debug, trace, req, data, async, func, poll.",
     ucc_offsetof(ucc_global_config_t, log_component.log_level), UCC_CONFIG_TYPE_ENUM(ucc_log_level_cfg_names)},
    {"COLL_TRACE", "warn",
     "UCC collective logging level. Higher level will result in more verbose collective info.\n"
     "Possible values are: fatal, error, warn, info, debug, trace.",
     ucc_offsetof(ucc_global_config_t, coll_trace_level), UCC_CONFIG_TYPE_ENUM(ucc_log_level_cfg_names)},
    {"WARN_UNUSED_ENV_VARS", "y", "Issue a warning about UCC_* environment variables that no configuration table uses.",
     ucc_offsetof(ucc_global_config_t, warn_unused_env_vars), UCC_CONFIG_TYPE_BOOL},
    {"PROFILE_MODE", "",
     "Profile collection modes. If none is specified, profiling is disabled.\n"
     " - log   - Record all timestamps.\n - accum - Accumulate measurements per location.",
     ucc_offsetof(ucc_global_config_t, profile_mode_str), UCC_CONFIG_TYPE_STRING},
    {"PROFILE_FILE", "ucc_%h_%p.prof",
     "File name to dump profiling data to. Substitutions: %h: host, %p: pid.",
     ucc_offsetof(ucc_global_config_t, profile_file), UCC_CONFIG_TYPE_STRING},
    {"PROFILE_LOG_SIZE", "4m", "Maximal size of profiling log. New records will replace old records.",
     ucc_offsetof(ucc_global_config_t, profile_log_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"CONFIG_FILE", "", "Location of configuration file (default: $HOME/ucc.conf, then <prefix>/share/ucc.conf)",
     ucc_offsetof(ucc_global_config_t, cfg_filename), UCC_CONFIG_TYPE_STRING},
    {"LOG_FILE", "",
     "If not empty, UCC will print log messages to the specified file instead of stdout.\n"
     "Substitutions: %p - process ID, %h - host name. 'stdout'/'stderr' select the stream.",
     ucc_offsetof(ucc_global_config_t, log_file), UCC_CONFIG_TYPE_STRING},
    {"LOG_FILE_SIZE", "inf", "The maximal size of log file. On reaching it the log is rotated.",
     ucc_offsetof(ucc_global_config_t, log_file_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"LOG_FILE_ROTATE", "0", "The maximal number of backup log files kept on rotation.",
     ucc_offsetof(ucc_global_config_t, log_file_rotate), UCC_CONFIG_TYPE_UINT},
    {"LOG_BUFFER", "1024", "Buffer size for a single log message / stdio buffer of the log file.",
     ucc_offsetof(ucc_global_config_t, log_buffer_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"LOG_DATA_SIZE", "0", "How much packet payload to print, at most, in data mode.",
     ucc_offsetof(ucc_global_config_t, log_data_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"LOG_PRINT_ENABLE", "y", "Enable output of ucc_print(). This option is intended for use by the library developers.",
     ucc_offsetof(ucc_global_config_t, log_print_enable), UCC_CONFIG_TYPE_BOOL},
    {"LOG_LEVEL_TRIGGER", "fatal", "Log level to trigger error handling (abort at the first message of this level).",
     ucc_offsetof(ucc_global_config_t, log_level_trigger), UCC_CONFIG_TYPE_ENUM(ucc_log_level_cfg_names)},
    {"CHECK_ASYMMETRIC_DT", "n",
     "Check that datatypes and memory types are consistent across ranks for rooted v-collectives "
     "(adds a service allreduce).",
     ucc_offsetof(ucc_global_config_t, check_asymmetric_dt), UCC_CONFIG_TYPE_BOOL},
    {"MODULE_DIR", "", "Directory with libucc_<framework>_<name>.so plugin modules (default: <libdir>/ucc).",
     ucc_offsetof(ucc_global_config_t, module_dir), UCC_CONFIG_TYPE_STRING},
    {NULL}};
UCC_CONFIG_REGISTER_TABLE(ucc_global_config_table, "UCC global", "", ucc_global_config_t)
