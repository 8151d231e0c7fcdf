/* Base objects and interfaces shared by collective layers (CL) and
 * transport layers (TL): lib / context / team / coll-init.
 * Contract: reference components/base/ucc_base_iface.h:49-272. */
#ifndef UCC_BASE_IFACE_H_
#define UCC_BASE_IFACE_H_
#include <ucc/api/ucc.h>
#include "utils/ucc_component.h"
#include "utils/ucc_parser.h"
#include "utils/ucc_log.h"
#include "utils/ucc_coll_utils.h"
#include "utils/ucc_proc_info.h"
#include "schedule/ucc_schedule.h"
#include "coll_score/ucc_coll_score.h"

typedef struct ucc_context ucc_context_t;
typedef struct ucc_lib_info ucc_lib_info_t;

/* ---- config shared by all CL/TL ---- */
typedef struct ucc_base_lib_config {
    ucc_config_global_list_entry_t *cfg_entry;
    ucc_log_component_config_t      log_component;
    int                             use_tuning;
    unsigned                        min_team_size;
} ucc_base_lib_config_t;
typedef struct ucc_base_ctx_config {
    ucc_config_global_list_entry_t *cfg_entry;
    struct ucc_base_lib            *lib;
    char                           *score_str; /* TUNE */
} ucc_base_ctx_config_t;
extern ucc_config_field_t ucc_base_lib_config_table[];
extern ucc_config_field_t ucc_base_ctx_config_table[];

/* ---- params / attrs ---- */
enum {
    UCC_BASE_LIB_FLAG_TEAM_ID_REQUIRED          = UCC_BIT(0),
    UCC_BASE_LIB_FLAG_SERVICE_TEAM_REQUIRED     = UCC_BIT(1),
    UCC_BASE_LIB_FLAG_CTX_SERVICE_TEAM_REQUIRED = UCC_BIT(2),
};
typedef struct ucc_base_lib_params { ucc_lib_params_t params; const char *full_prefix; } ucc_base_lib_params_t;
typedef struct ucc_base_lib_attr { ucc_lib_attr_t attr; uint64_t flags; ucc_rank_t min_team_size, max_team_size; } ucc_base_lib_attr_t;
typedef struct ucc_base_lib_properties { ucc_rank_t default_team_size, min_team_size, max_team_size; } ucc_base_lib_properties_t;

typedef struct ucc_base_context_params {
    ucc_context_params_t params;
    int                  estimated_num_eps;
    int                  estimated_num_ppn;
    ucc_thread_mode_t    thread_mode;
    const char          *prefix;
    ucc_context_t       *context;
} ucc_base_context_params_t;
typedef struct ucc_base_ctx_attr { ucc_context_attr_t attr; int topo_required; } ucc_base_ctx_attr_t;

typedef struct ucc_base_team_params {
    ucc_team_params_t params;
    int               scope;    /* ucc_cl_type_t of the creator, or UCC_CL_LAST for core */
    int               scope_id;
    ucc_rank_t        rank;
    ucc_rank_t        size;
    uint16_t          id;
    ucc_team_t       *team;     /* core team */
    ucc_ep_map_t      map;      /* base-team rank -> core-team rank */
} ucc_base_team_params_t;

/* ---- objects ---- */
typedef struct ucc_base_lib {
    ucc_log_component_config_t log_component;
    int                        use_tuning;
    unsigned                   min_team_size;
} ucc_base_lib_t;
typedef struct ucc_base_context { ucc_context_t *ucc_context; ucc_base_lib_t *lib; } ucc_base_context_t;
struct ucc_base_team { ucc_base_context_t *context; ucc_base_team_params_t params; };
static inline const char *ucc_base_team_name(const ucc_base_team_t *t) { return t->context->lib->log_component.name; }

typedef struct ucc_base_coll_alg_info { unsigned id; const char *name; const char *desc; } ucc_base_coll_alg_info_t;

/* ---- interfaces ---- */
typedef struct ucc_base_lib_iface {
    ucc_status_t (*init)(const ucc_base_lib_params_t *params, const ucc_base_lib_config_t *config, ucc_base_lib_t **lib);
    void         (*finalize)(ucc_base_lib_t *lib);
    ucc_status_t (*get_attr)(const ucc_base_lib_t *lib, ucc_base_lib_attr_t *attr);
    ucc_status_t (*get_properties)(ucc_base_lib_properties_t *prop);
} ucc_base_lib_iface_t;

typedef struct ucc_base_context_iface {
    ucc_status_t (*create)(const ucc_base_context_params_t *params, const ucc_base_ctx_config_t *config, ucc_base_context_t **ctx);
    ucc_status_t (*create_epilog)(ucc_base_context_t *ctx);
    void         (*destroy)(ucc_base_context_t *ctx);
    ucc_status_t (*get_attr)(const ucc_base_context_t *ctx, ucc_base_ctx_attr_t *attr);
    ucc_status_t (*mem_map)(const ucc_base_context_t *ctx, ucc_mem_map_mode_t mode, void *memh, void *tl_h);
    ucc_status_t (*mem_unmap)(const ucc_base_context_t *ctx, ucc_mem_map_mode_t mode, void *tl_h);
    ucc_status_t (*memh_pack)(const ucc_base_context_t *ctx, ucc_mem_map_mode_t mode, void *tl_h, void **pack_buffer);
} ucc_base_context_iface_t;

typedef struct ucc_base_team_iface {
    ucc_status_t (*create_post)(ucc_base_context_t *ctx, const ucc_base_team_params_t *params, ucc_base_team_t **team);
    ucc_status_t (*create_test)(ucc_base_team_t *team);
    ucc_status_t (*destroy)(ucc_base_team_t *team);
    ucc_status_t (*get_scores)(ucc_base_team_t *team, ucc_coll_score_t **score);
} ucc_base_team_iface_t;

typedef struct ucc_base_coll_iface { ucc_base_coll_init_fn_t init; } ucc_base_coll_iface_t;

/* logging through a base lib */
#define ucc_base_log(_lib, _lvl, _fmt, ...) ucc_log_component(_lvl, &(_lib)->log_component, _fmt, ##__VA_ARGS__)
#endif
