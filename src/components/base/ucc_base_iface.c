#include "ucc_base_iface.h"
#include "core/ucc_global_opts.h"

ucc_config_field_t ucc_base_lib_config_table[] = {
    {"LOG_LEVEL", "warn", "UCC logging level of the component. Messages with a level higher or equal to the selected will be printed.\n"
     "Possible values are: fatal, error, warn, diag, info, debug, trace, req, data, async, func, poll.",
     ucc_offsetof(ucc_base_lib_config_t, log_component.log_level), UCC_CONFIG_TYPE_ENUM(ucc_log_level_cfg_names)},
    {"USE_TUNING", "y", "Use the component's built-in per-message-size algorithm selection (y) or plain default scores (n).",
     ucc_offsetof(ucc_base_lib_config_t, use_tuning), UCC_CONFIG_TYPE_BOOL},
    {"MIN_TEAM_SIZE", "auto", "Minimal team size the component accepts (auto: component default).",
     ucc_offsetof(ucc_base_lib_config_t, min_team_size), UCC_CONFIG_TYPE_UINT},
    {NULL}};

ucc_config_field_t ucc_base_ctx_config_table[] = {
    {"TUNE", "",
     "Collective tuning modifier: token1#token2#...#tokenN with token = coll_type:msg_range:mem_type:[team_size]:score:@alg.\n"
     "Every qualifier is optional but either score or @alg must be given.\n"
     "  coll_type = comma separated collective names\n"
     "  msg_range = start-end[,start-end..] with K/M/G suffixes and 'inf'\n"
     "  mem_type  = host|cuda|cuda_managed|rocm|rocm_managed (comma separated)\n"
     "  team_size = [lo-hi,...]\n  score     = 0..inf (0 disables, inf forces)\n  alg       = @<id|name>\n"
     "The value \"0\" disables the component.",
     ucc_offsetof(ucc_base_ctx_config_t, score_str), UCC_CONFIG_TYPE_STRING},
    {NULL}};
