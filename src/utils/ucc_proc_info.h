/* Process / host identity used by topology code.  Fields follow the facts
 * the reference records (utils/ucc_proc_info.h:31-101) — host hash, socket,
 * numa, pid, per-host GPU table with NVLink capabilities and the NVLink
 * adjacency matrix — re-laid-out for this library. */
#ifndef UCC_PROC_INFO_H_
#define UCC_PROC_INFO_H_
#include "ucc_compiler_def.h"
#include <sys/types.h>

#define UCC_MAX_HOST_GPUS 16
#define UCC_MAX_HOST_NICS 16
typedef uint64_t ucc_host_id_t;
typedef uint8_t  ucc_socket_id_t;
typedef uint8_t  ucc_numa_id_t;
#define UCC_SOCKET_ID_INVALID ((ucc_socket_id_t)-1)
#define UCC_NUMA_ID_INVALID   ((ucc_numa_id_t)-1)

typedef struct ucc_proc_info {
    ucc_host_id_t   host_hash;
    ucc_socket_id_t socket_id;
    ucc_numa_id_t   numa_id;
    ucc_host_id_t   host_id;   /* dense index assigned by ctx topo */
    pid_t           pid;
} ucc_proc_info_t;

enum { UCC_GPU_CAP_NVLINK = 1, UCC_GPU_CAP_NVSWITCH = 2, UCC_GPU_CAP_FABRIC = 4, UCC_GPU_CAP_MULTICAST = 8 };
typedef struct ucc_pci_id { uint16_t domain; uint8_t bus, device, function; } ucc_pci_id_t;
typedef struct ucc_gpu_info {
    ucc_pci_id_t pci;
    uint32_t     caps;
    uint32_t     fabric_clique_id;
    uint64_t     fabric_partition;  /* hash of cluster uuid */
    uint64_t     uuid_hash;
    uint32_t     n_nvlinks;
    uint32_t     sm_count;
} ucc_gpu_info_t;
typedef struct ucc_nic_info { char name[16]; uint64_t guid; uint8_t port, numa; } ucc_nic_info_t;

typedef struct ucc_host_info {
    ucc_host_id_t  host_hash;
    uint8_t        n_gpus, n_nics;
    uint16_t       visible_gpus;   /* bit i: gpu i visible to this process (CUDA_VISIBLE_DEVICES) */
    int8_t         current_gpu;    /* device this process drives, -1 if none */
    ucc_gpu_info_t gpus[UCC_MAX_HOST_GPUS];
    ucc_nic_info_t nics[UCC_MAX_HOST_NICS];
    uint8_t        nvlink_matrix[UCC_MAX_HOST_GPUS][UCC_MAX_HOST_GPUS]; /* #links (direct or via switch) */
} ucc_host_info_t;

extern ucc_proc_info_t ucc_local_proc;
ucc_status_t ucc_local_proc_info_init(void);
#define UCC_PROC_INFO_EQUAL(_a, _b) ((_a).host_hash == (_b).host_hash && (_a).pid == (_b).pid)
#endif
