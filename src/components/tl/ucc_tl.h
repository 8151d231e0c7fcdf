/* Transport layer framework (contract: reference tl/ucc_tl.h:50-82, tl/ucc_tl.c). */
#ifndef UCC_TL_H_
#define UCC_TL_H_
#include "components/base/ucc_base_iface.h"

typedef struct ucc_tl_iface ucc_tl_iface_t;
typedef struct ucc_tl_lib_config { ucc_base_lib_config_t super; ucc_tl_iface_t *iface; } ucc_tl_lib_config_t;
typedef struct ucc_tl_context_config { ucc_base_ctx_config_t super; struct ucc_tl_lib *tl_lib; } ucc_tl_context_config_t;
extern ucc_config_field_t ucc_tl_lib_config_table[];
extern ucc_config_field_t ucc_tl_context_config_table[];

typedef struct ucc_tl_lib { ucc_base_lib_t super; ucc_tl_iface_t *iface; } ucc_tl_lib_t;
typedef struct ucc_tl_context { ucc_base_context_t super; int ref_count; } ucc_tl_context_t;
typedef struct ucc_tl_team { ucc_base_team_t super; } ucc_tl_team_t;

/* service collectives: used by core for team-id allocation, internal OOB, dt checks */
typedef struct ucc_tl_service_coll {
    ucc_status_t (*allreduce)(ucc_base_team_t *team, void *sbuf, void *rbuf, ucc_datatype_t dt, size_t count,
                              ucc_reduction_op_t op, ucc_subset_t subset, ucc_coll_task_t **task);
    ucc_status_t (*allgather)(ucc_base_team_t *team, void *sbuf, void *rbuf, size_t msgsize, ucc_subset_t subset, ucc_coll_task_t **task);
    ucc_status_t (*bcast)(ucc_base_team_t *team, void *buf, size_t msgsize, ucc_rank_t root, ucc_subset_t subset, ucc_coll_task_t **task);
    void         (*update_id)(ucc_base_team_t *team, uint16_t id);
} ucc_tl_service_coll_t;

/* optional plugins adding algorithms to a TL (reference tl/ucc_tl.h:64-69) */
typedef struct ucc_tl_coll_plugin_iface {
    ucc_component_iface_t          super;
    ucc_score_t                    default_score;
    ucc_config_global_list_entry_t config;
    ucc_status_t (*get_scores)(ucc_base_team_t *team, ucc_coll_score_t **score);
    uint32_t                       id;
} ucc_tl_coll_plugin_iface_t;

struct ucc_tl_iface {
    ucc_component_iface_t          super;
    ucc_config_global_list_entry_t tl_lib_config;
    ucc_config_global_list_entry_t tl_context_config;
    ucc_base_lib_iface_t           lib;
    ucc_base_context_iface_t       context;
    ucc_base_team_iface_t          team;
    ucc_base_coll_iface_t          coll;
    ucc_tl_service_coll_t          scoll;
    ucc_component_framework_t      coll_plugins;
    const ucc_base_coll_alg_info_t *alg_info[UCC_COLL_TYPE_NUM];
};

#define UCC_TL_CTX_IFACE(_ctx) (ucc_derived_of((_ctx)->super.lib, ucc_tl_lib_t)->iface)
#define UCC_TL_TEAM_IFACE(_team) UCC_TL_CTX_IFACE(ucc_derived_of((_team)->super.context, ucc_tl_context_t))
/* the macros below take a TL-specific team whose first member is ucc_tl_team_t */
#define UCC_TL_TEAM_LIB(_team) ((_team)->super.super.context->lib)
#define UCC_TL_TEAM_RANK(_team) ((_team)->super.super.params.rank)
#define UCC_TL_TEAM_SIZE(_team) ((_team)->super.super.params.size)
#define UCC_TL_TEAM_MAP(_team) ((_team)->super.super.params.map)
#define UCC_TL_CORE_CTX(_team) ((_team)->super.super.context->ucc_context)
#define UCC_TL_CORE_TEAM(_team) ((_team)->super.super.params.team)
#define tl_log(_lib, _lvl, _f, ...) ucc_log_component(_lvl, &((ucc_base_lib_t *)(_lib))->log_component, _f, ##__VA_ARGS__)
#define tl_error(_lib, _f, ...) tl_log(_lib, UCC_LOG_LEVEL_ERROR, _f, ##__VA_ARGS__)
#define tl_warn(_lib, _f, ...)  tl_log(_lib, UCC_LOG_LEVEL_WARN, _f, ##__VA_ARGS__)
#define tl_info(_lib, _f, ...)  tl_log(_lib, UCC_LOG_LEVEL_INFO, _f, ##__VA_ARGS__)
#define tl_debug(_lib, _f, ...) tl_log(_lib, UCC_LOG_LEVEL_DEBUG, _f, ##__VA_ARGS__)
#define tl_trace(_lib, _f, ...) tl_log(_lib, UCC_LOG_LEVEL_TRACE, _f, ##__VA_ARGS__)

ucc_status_t ucc_tl_lib_config_read(ucc_tl_iface_t *iface, const char *full_prefix, ucc_tl_lib_config_t **cfg);
ucc_status_t ucc_tl_context_config_read(ucc_tl_lib_t *tl_lib, const char *full_prefix, ucc_tl_context_config_t **cfg);
void         ucc_tl_lib_config_release(ucc_tl_iface_t *iface, ucc_tl_lib_config_t *cfg);
void         ucc_tl_context_config_release(ucc_tl_iface_t *iface, ucc_tl_context_config_t *cfg);
ucc_status_t ucc_tl_context_get(ucc_context_t *ctx, const char *name, ucc_tl_context_t **tl_context);
ucc_status_t ucc_tl_context_put(ucc_tl_context_t *tl_context);
/* fills base lib fields from config */
void         ucc_tl_lib_init_base(ucc_tl_lib_t *lib, ucc_tl_iface_t *iface, const ucc_tl_lib_config_t *cfg);
/* checks that the team is within [min,max] team size of the TL */
ucc_status_t ucc_tl_team_size_check(ucc_tl_iface_t *iface, ucc_tl_lib_t *lib, ucc_rank_t size);
/* common tail of every TL get_scores(): applies USE_TUNING default string and the user's TUNE */
ucc_status_t ucc_tl_apply_tune(ucc_tl_team_t *team, ucc_coll_score_t *score, const ucc_coll_score_team_info_t *info,
                               const char *default_select_str, const char *user_tune);

/* sequential non-blocking creation of several TL teams (used by CLs) */
typedef struct ucc_team_multiple_req_descr { ucc_tl_context_t *ctx; ucc_tl_team_t *team; ucc_base_team_params_t param; ucc_status_t status; uint64_t id; } ucc_team_multiple_req_descr_t;
typedef struct ucc_team_multiple_req { int n_teams; int last; ucc_team_multiple_req_descr_t descs[1]; } ucc_team_multiple_req_t;
ucc_status_t ucc_team_multiple_req_alloc(ucc_team_multiple_req_t **req, int n_teams);
ucc_status_t ucc_tl_team_create_multiple(ucc_team_multiple_req_t *req);
ucc_status_t ucc_tl_team_destroy_multiple(ucc_team_multiple_req_t *req);
void         ucc_team_multiple_req_free(ucc_team_multiple_req_t *req);
#endif
