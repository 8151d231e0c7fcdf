/* sysinfo plugins fill the host description (GPUs, NVLink matrix, NICs).
 * Role of reference components/topo/{cuda,ib} sysinfo modules. */
#ifndef UCC_SYSINFO_H_
#define UCC_SYSINFO_H_
#include "utils/ucc_component.h"
#include "utils/ucc_proc_info.h"
typedef struct ucc_sysinfo_base {
    ucc_component_iface_t super;
    ucc_status_t (*init)(void);
    ucc_status_t (*get_host_info)(ucc_host_info_t *info);
    ucc_status_t (*finalize)(void);
} ucc_sysinfo_base_t;
#endif
