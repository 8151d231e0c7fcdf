/* Library object (contract: reference core/ucc_lib.c:291-518). */
#ifndef UCC_LIB_H_
#define UCC_LIB_H_
#include <ucc/api/ucc.h>
#include "utils/ucc_parser.h"
#include "utils/ucc_mpool.h"
#include "components/cl/ucc_cl.h"
#include "components/tl/ucc_tl.h"

typedef struct ucc_lib_config {
    char                    *full_prefix;
    ucc_config_names_array_t cls;
} ucc_lib_config_t;

typedef struct ucc_lib_info {
    char              *full_prefix;
    int                n_cl_libs_opened;
    int                n_tl_libs_opened;
    ucc_cl_lib_t     **cl_libs;
    ucc_tl_lib_t     **tl_libs;
    ucc_lib_attr_t     attr;
    int                specific_cls_requested;
    ucc_cl_lib_attr_t *cl_attrs;
    ucc_mpool_t        stub_tasks_mp;
} ucc_lib_info_t;

int ucc_tl_is_required(ucc_lib_info_t *lib, ucc_tl_iface_t *tl_iface, int forced);
extern ucc_host_info_t ucc_local_host; /* filled by sysinfo components at ucc_init */
#endif
