/* Collective layer framework (contract: reference cl/ucc_cl.h:62-72, cl/ucc_cl.c). */
#ifndef UCC_CL_H_
#define UCC_CL_H_
#include "components/base/ucc_base_iface.h"
#include "components/cl/ucc_cl_type.h"

typedef struct ucc_cl_iface ucc_cl_iface_t;
typedef struct ucc_cl_lib_config { ucc_base_lib_config_t super; ucc_cl_iface_t *iface; ucc_config_allow_list_t tls; } ucc_cl_lib_config_t;
typedef struct ucc_cl_context_config { ucc_base_ctx_config_t super; struct ucc_cl_lib *cl_lib; } ucc_cl_context_config_t;
extern ucc_config_field_t ucc_cl_lib_config_table[];
extern ucc_config_field_t ucc_cl_context_config_table[];

typedef struct ucc_cl_lib { ucc_base_lib_t super; ucc_cl_iface_t *iface; ucc_config_names_list_t tls; int tls_forced; } ucc_cl_lib_t;
typedef struct ucc_cl_context { ucc_base_context_t super; } ucc_cl_context_t;
typedef struct ucc_cl_team { ucc_base_team_t super; } ucc_cl_team_t;
typedef struct ucc_cl_lib_attr { ucc_base_lib_attr_t super; ucc_config_names_array_t *tls; int tls_forced; } ucc_cl_lib_attr_t;

struct ucc_cl_iface {
    ucc_component_iface_t          super;
    ucc_cl_type_t                  type;
    ucc_config_global_list_entry_t cl_lib_config;
    ucc_config_global_list_entry_t cl_context_config;
    ucc_base_lib_iface_t           lib;
    ucc_base_context_iface_t       context;
    ucc_base_team_iface_t          team;
    ucc_base_coll_iface_t          coll;
    const ucc_base_coll_alg_info_t *alg_info[UCC_COLL_TYPE_NUM];
};

#define UCC_CL_CTX_IFACE(_ctx) (ucc_derived_of((_ctx)->super.lib, ucc_cl_lib_t)->iface)
#define UCC_CL_TEAM_IFACE(_team) UCC_CL_CTX_IFACE(ucc_derived_of((_team)->super.context, ucc_cl_context_t))
#define UCC_CL_TEAM_LIB(_team) ((_team)->super.context->lib)
#define UCC_CL_TEAM_RANK(_team) ((_team)->super.params.rank)
#define UCC_CL_TEAM_SIZE(_team) ((_team)->super.params.size)
#define UCC_CL_CORE_CTX(_team) ((_team)->super.context->ucc_context)
#define cl_log(_lib, _lvl, _f, ...) ucc_log_component(_lvl, &((ucc_base_lib_t *)(_lib))->log_component, _f, ##__VA_ARGS__)
#define cl_error(_lib, _f, ...) cl_log(_lib, UCC_LOG_LEVEL_ERROR, _f, ##__VA_ARGS__)
#define cl_warn(_lib, _f, ...)  cl_log(_lib, UCC_LOG_LEVEL_WARN, _f, ##__VA_ARGS__)
#define cl_info(_lib, _f, ...)  cl_log(_lib, UCC_LOG_LEVEL_INFO, _f, ##__VA_ARGS__)
#define cl_debug(_lib, _f, ...) cl_log(_lib, UCC_LOG_LEVEL_DEBUG, _f, ##__VA_ARGS__)

ucc_status_t ucc_cl_lib_config_read(ucc_cl_iface_t *iface, const char *full_prefix, ucc_cl_lib_config_t **cfg);
ucc_status_t ucc_cl_context_config_read(ucc_cl_lib_t *cl_lib, const char *full_prefix, ucc_cl_context_config_t **cfg);
void         ucc_cl_lib_config_release(ucc_cl_iface_t *iface, ucc_cl_lib_config_t *cfg);
void         ucc_cl_context_config_release(ucc_cl_iface_t *iface, ucc_cl_context_config_t *cfg);
ucc_status_t ucc_cl_lib_init_base(ucc_cl_lib_t *lib, ucc_cl_iface_t *iface, const ucc_cl_lib_config_t *cfg);
void         ucc_cl_lib_cleanup_base(ucc_cl_lib_t *lib);
const char  *ucc_cl_name(ucc_cl_type_t t);
ucc_status_t ucc_parse_cls_string(const char *cls_str, ucc_cl_type_t **cls_array, int *n_cls);
#endif
