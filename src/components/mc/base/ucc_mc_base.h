/* Memory component interface: alloc/free (pooled), memcpy, memset, pointer
 * classification.  One MC per memory type (cpu, cuda).
 * Role of reference mc/base/ucc_mc_base.h:104-127. */
#ifndef UCC_MC_BASE_H_
#define UCC_MC_BASE_H_
#include <ucc/api/ucc.h>
#include "utils/ucc_component.h"
#include "utils/ucc_parser.h"

typedef struct ucc_mem_attr {
    uint64_t          field_mask;
    ucc_memory_type_t mem_type;
    void             *base_address;
    size_t            alloc_length;
} ucc_mem_attr_t;
enum { UCC_MEM_ATTR_FIELD_MEM_TYPE = UCC_BIT(0), UCC_MEM_ATTR_FIELD_BASE_ADDRESS = UCC_BIT(1), UCC_MEM_ATTR_FIELD_ALLOC_LENGTH = UCC_BIT(2) };

typedef struct ucc_mc_buffer_header {
    ucc_memory_type_t mt;
    int               from_pool;
    void             *addr;
} ucc_mc_buffer_header_t;

typedef struct ucc_mc_params { ucc_thread_mode_t thread_mode; } ucc_mc_params_t;
typedef struct ucc_mc_attr { uint64_t field_mask; size_t fast_alloc_size; } ucc_mc_attr_t;
enum { UCC_MC_ATTR_FIELD_THREAD_MODE = UCC_BIT(0), UCC_MC_ATTR_FIELD_FAST_ALLOC_SIZE = UCC_BIT(1) };

typedef struct ucc_mc_config { ucc_log_component_config_t log_component; } ucc_mc_config_t;
extern ucc_config_field_t ucc_mc_config_table[];

typedef struct ucc_mc_ops {
    ucc_status_t (*mem_query)(const void *ptr, ucc_mem_attr_t *mem_attr);
    ucc_status_t (*mem_alloc)(ucc_mc_buffer_header_t **h_ptr, size_t size, ucc_memory_type_t mt);
    ucc_status_t (*mem_free)(ucc_mc_buffer_header_t *h_ptr);
    ucc_status_t (*memcpy)(void *dst, const void *src, size_t len, ucc_memory_type_t dst_mem, ucc_memory_type_t src_mem);
    ucc_status_t (*memset)(void *dst, int value, size_t len);
    ucc_status_t (*flush)(void);
} ucc_mc_ops_t;

typedef struct ucc_mc_base {
    ucc_component_iface_t           super;
    uint32_t                        ref_cnt;
    ucc_ee_type_t                   ee_type;
    ucc_memory_type_t               type;
    ucc_mc_config_t                *config;
    ucc_config_global_list_entry_t  config_table;
    ucc_status_t (*init)(const ucc_mc_params_t *mc_params);
    ucc_status_t (*get_attr)(ucc_mc_attr_t *mc_attr);
    ucc_status_t (*finalize)(void);
    ucc_mc_ops_t                    ops;
} ucc_mc_base_t;
#endif
