/* ucc_b200 — opaque handles and scalar typedefs of the UCC-compatible API.
 * Contract: reference src/ucc/api/ucc_def.h:25-174 (names, widths, dt encoding). */
#ifndef UCC_DEF_H_
#define UCC_DEF_H_

#include <ucc/api/ucc_status.h>
#include <stddef.h>
#include <stdint.h>

#define UCC_BIT(i)  (1ul << (i))
#define UCC_MASK(i) (UCC_BIT(i) - 1)

/* object handles */
typedef struct ucc_lib_info       *ucc_lib_h;
typedef struct ucc_context        *ucc_context_h;
typedef struct ucc_team           *ucc_team_h;
typedef struct ucc_ee             *ucc_ee_h;
typedef struct ucc_mem_handle     *ucc_mem_h;
typedef struct ucc_lib_config     *ucc_lib_config_h;
typedef struct ucc_context_config *ucc_context_config_h;

/* A collective request: the only user-visible field is its status. */
typedef struct ucc_coll_req {
    ucc_status_t status;
} ucc_coll_req_t;
typedef struct ucc_coll_req *ucc_coll_req_h;

/* completion callback; may only call ucc_collective_finalize */
typedef struct ucc_coll_callback {
    void (*cb)(void *data, ucc_status_t status);
    void  *data;
} ucc_coll_callback_t;

typedef uint64_t ucc_count_t;
typedef uint64_t ucc_aint_t;
typedef uint16_t ucc_coll_id_t;
typedef void    *ucc_p2p_conn_t;
typedef void    *ucc_context_addr_h;
typedef size_t   ucc_context_addr_len_t;

typedef enum {
    UCC_CONFIG_PRINT_CONFIG = UCC_BIT(0),
    UCC_CONFIG_PRINT_HEADER = UCC_BIT(1),
    UCC_CONFIG_PRINT_DOC    = UCC_BIT(2),
    UCC_CONFIG_PRINT_HIDDEN = UCC_BIT(3)
} ucc_config_print_flags_t;

/* datatype handle encoding: low 3 bits = class, rest = id (predefined) or
 * pointer (generic, 8-byte aligned) */
typedef struct ucc_dt_generic ucc_dt_generic_t;
typedef enum {
    UCC_DATATYPE_PREDEFINED = 0,
    UCC_DATATYPE_GENERIC    = UCC_BIT(0),
    UCC_DATATYPE_SHIFT      = 3,
    UCC_DATATYPE_CLASS_MASK = UCC_MASK(UCC_DATATYPE_SHIFT)
} ucc_dt_type_t;

#define UCC_PREDEFINED_DT(_id) \
    (ucc_datatype_t)((((uint64_t)(_id)) << UCC_DATATYPE_SHIFT) | (UCC_DATATYPE_PREDEFINED))

#endif
