/* Context object: TL/CL contexts, progress queue, address storage built by
 * the OOB exchange, context topology, service team.
 * Contract: reference core/ucc_context.{h,c}. */
#ifndef UCC_CONTEXT_H_
#define UCC_CONTEXT_H_
#include <ucc/api/ucc.h>
#include "ucc_lib.h"
#include "ucc_progress_queue.h"
#include "utils/ucc_proc_info.h"

typedef struct ucc_context_id { ucc_proc_info_t pi; uint32_t seq_num; } ucc_context_id_t;
#define UCC_CTX_ID_EQUAL(_a, _b) (UCC_PROC_INFO_EQUAL((_a).pi, (_b).pi) && (_a).seq_num == (_b).seq_num)

typedef struct ucc_addr_storage {
    void             *storage;   /* size * addr_len bytes, rank-ordered */
    void             *oob_req;
    size_t            addr_len;
    ucc_rank_t        size, rank;
    int               flags;     /* 1: all ranks share one node (single-node job) */
} ucc_addr_storage_t;

/* packed address = header | component table | blobs */
typedef struct ucc_context_addr_header {
    ucc_context_id_t ctx_id;
    ucc_host_info_t  host_info;
    int              n_components;
    struct { unsigned long id; uint32_t offset; uint32_t len; } components[1];
} ucc_context_addr_header_t;
#define UCC_CONTEXT_ADDR_HEADER_SIZE(_n) (sizeof(ucc_context_addr_header_t) + ((_n) > 0 ? (_n) - 1 : 0) * sizeof(((ucc_context_addr_header_t *)0)->components[0]))
#define UCC_ADDR_STORAGE_RANK_HEADER(_s, _r) ((ucc_context_addr_header_t *)((char *)(_s)->storage + (size_t)(_r) * (_s)->addr_len))

typedef struct ucc_context_config {
    ucc_lib_info_t           *lib;
    ucc_cl_context_config_t **cl_cfgs;
    ucc_tl_context_config_t **tl_cfgs;
    int                       n_cl_cfg, n_tl_cfg;
    uint32_t                  team_ids_pool_size;
    uint32_t                  estimated_num_eps, estimated_num_ppn;
    uint32_t                  lock_free_progress_q;
    uint32_t                  internal_oob;
    uint32_t                  throttle_progress;
    ucc_config_names_array_t  net_devices;
    uint32_t                  node_local_id;
    char                     *service_tl;
} ucc_context_config_t;

typedef ucc_status_t (*ucc_context_progress_fn_t)(void *progress_arg);
typedef struct ucc_context_progress_entry { ucc_context_progress_fn_t fn; void *arg; } ucc_context_progress_entry_t;

typedef struct ucc_context_topo ucc_context_topo_t;
typedef struct ucc_tl_team ucc_tl_team_t;

typedef struct ucc_context {
    ucc_lib_info_t        *lib;
    ucc_context_params_t   params;
    ucc_context_attr_t     attr;     /* cached packed address */
    ucc_thread_mode_t      thread_mode;
    ucc_cl_context_t     **cl_ctx;
    ucc_tl_context_t     **tl_ctx;
    ucc_tl_context_t      *service_ctx;
    int                    n_cl_ctx, n_tl_ctx;
    ucc_config_names_array_t all_tls;
    ucc_progress_queue_t  *pq;
    ucc_context_id_t       id;
    ucc_rank_t             rank;     /* rank in the context OOB */
    ucc_addr_storage_t     addr_storage;
    ucc_tl_team_t         *service_team;
    uint64_t              *ids_pool; /* bitmap of free team ids (1 = free) */
    uint32_t               ids_pool_size;
    int                    throttle_progress, progress_counter;
    ucc_context_progress_entry_t *progress_fns;
    int                    n_progress_fns;
    ucc_spinlock_t         progress_lock;
    ucc_context_topo_t    *topo;
    uint64_t               cl_flags;
    ucc_proc_info_t        proc_info; /* possibly injected by tests */
    ucc_host_info_t        host_info;
    uint32_t               coll_seq;  /* seq_num source for coll trace */
} ucc_context_t;

/* internal entry used by the test harness to inject a synthetic process placement */
ucc_status_t ucc_context_create_proc_info(ucc_lib_h lib, const ucc_context_params_t *params, const ucc_context_config_h config,
                                          ucc_context_h *context, ucc_proc_info_t *proc_info);
ucc_status_t ucc_context_progress_register(ucc_context_t *ctx, ucc_context_progress_fn_t fn, void *progress_arg);
void         ucc_context_progress_deregister(ucc_context_t *ctx, ucc_context_progress_fn_t fn, void *progress_arg);
/* non-blocking two-stage allgather of packed addresses; call until != INPROGRESS */
ucc_status_t ucc_core_addr_exchange(ucc_context_t *context, ucc_oob_coll_t *oob, ucc_addr_storage_t *addr_storage);
void         ucc_addr_storage_free(ucc_addr_storage_t *s);
/* component blob of `rank` (ctx rank) or NULL */
void        *ucc_addr_storage_component_addr(ucc_addr_storage_t *s, ucc_rank_t rank, unsigned long component_id, uint32_t *len);
void *ucc_mem_map_tl_handle(ucc_mem_map_mem_h memh, const ucc_base_context_t *tl_ctx); /* core/ucc_mem_map.c */
#endif
