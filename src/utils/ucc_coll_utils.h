/* Collective-argument helpers: ep maps, subsets, counts/displacements,
 * message size used for score lookup, name tables.
 * Reference behaviour: utils/ucc_coll_utils.c:109-432 (re-implemented). */
#ifndef UCC_COLL_UTILS_H_
#define UCC_COLL_UTILS_H_
#include <ucc/api/ucc.h>
#include "ucc_compiler_def.h"
#include "ucc_math.h"
#include "core/ucc_dt.h"

#define UCC_COLL_TYPE_NUM 16
#define UCC_COLL_TYPE_ALL ((1u << UCC_COLL_TYPE_NUM) - 1)
#define UCC_EP_MAP_REVERSE_INTERNAL ((ucc_ep_map_type_t)5) /* ep = ep_num-1-rank */

#define UCC_IS_INPLACE(_a)    (((_a).mask & UCC_COLL_ARGS_FIELD_FLAGS) && ((_a).flags & UCC_COLL_ARGS_FLAG_IN_PLACE))
#define UCC_IS_PERSISTENT(_a) (((_a).mask & UCC_COLL_ARGS_FIELD_FLAGS) && ((_a).flags & UCC_COLL_ARGS_FLAG_PERSISTENT))
#define UCC_COLL_IS_COUNT_64(_a) (((_a)->mask & UCC_COLL_ARGS_FIELD_FLAGS) && ((_a)->flags & UCC_COLL_ARGS_FLAG_COUNT_64BIT))
#define UCC_COLL_IS_DISPL_64(_a) (((_a)->mask & UCC_COLL_ARGS_FIELD_FLAGS) && ((_a)->flags & UCC_COLL_ARGS_FLAG_DISPLACEMENTS_64BIT))
#define UCC_COLL_ARGS_ACTIVE_SET(_a) ((_a)->mask & UCC_COLL_ARGS_FIELD_ACTIVE_SET)
#define UCC_COLL_ARGS_HAS_TIMEOUT(_a) (((_a)->mask & UCC_COLL_ARGS_FIELD_FLAGS) && ((_a)->flags & UCC_COLL_ARGS_FLAG_TIMEOUT))
#define UCC_MEM_TYPE_MASK_FULL ((1u << UCC_MEMORY_TYPE_LAST) - 1)

typedef struct ucc_subset { ucc_ep_map_t map; ucc_rank_t myrank; } ucc_subset_t;

static inline int ucc_coll_type_index(ucc_coll_type_t ct) { return (int)ucc_ilog2((uint64_t)ct); }
static inline int ucc_coll_is_rooted(ucc_coll_type_t ct)
{ return !!(ct & (UCC_COLL_TYPE_BCAST | UCC_COLL_TYPE_FANIN | UCC_COLL_TYPE_FANOUT | UCC_COLL_TYPE_GATHER | UCC_COLL_TYPE_GATHERV |
                  UCC_COLL_TYPE_REDUCE | UCC_COLL_TYPE_SCATTER | UCC_COLL_TYPE_SCATTERV)); }
static inline int ucc_coll_has_reduction(ucc_coll_type_t ct)
{ return !!(ct & (UCC_COLL_TYPE_ALLREDUCE | UCC_COLL_TYPE_REDUCE | UCC_COLL_TYPE_REDUCE_SCATTER | UCC_COLL_TYPE_REDUCE_SCATTERV)); }

ucc_rank_t   ucc_ep_map_eval(ucc_ep_map_t map, ucc_rank_t rank);
ucc_ep_map_t ucc_ep_map_from_array(ucc_rank_t **array, ucc_rank_t size, ucc_rank_t full_size, int need_free);
ucc_ep_map_t ucc_ep_map_from_array_64(uint64_t **array, ucc_rank_t size, ucc_rank_t full_size, int need_free);
ucc_ep_map_t ucc_ep_map_create_full(ucc_rank_t size);
ucc_ep_map_t ucc_ep_map_create_reverse(ucc_rank_t size);
ucc_status_t ucc_ep_map_create_inverse(ucc_ep_map_t map, ucc_ep_map_t *inv, int need_free_array);
ucc_status_t ucc_ep_map_create_nested(ucc_ep_map_t *base, ucc_ep_map_t *sub, ucc_ep_map_t *out);
void         ucc_ep_map_destroy(ucc_ep_map_t *map);      /* frees ARRAY storage */
void         ucc_ep_map_destroy_nested(ucc_ep_map_t *map);
ucc_rank_t   ucc_ep_map_local_rank(ucc_ep_map_t map, ucc_rank_t ep); /* inverse lookup; UCC_RANK_INVALID if absent */
int          ucc_ep_map_is_identity(const ucc_ep_map_t *map);
static inline ucc_rank_t ucc_subset_size(const ucc_subset_t *s) { return (ucc_rank_t)s->map.ep_num; }

static inline uint64_t ucc_coll_args_get_count(const ucc_coll_args_t *a, const ucc_count_t *counts, ucc_rank_t i)
{ return UCC_COLL_IS_COUNT_64(a) ? ((const uint64_t *)counts)[i] : (uint64_t)((const uint32_t *)counts)[i]; }
static inline uint64_t ucc_coll_args_get_displacement(const ucc_coll_args_t *a, const ucc_aint_t *d, ucc_rank_t i)
{ return UCC_COLL_IS_DISPL_64(a) ? ((const uint64_t *)d)[i] : (uint64_t)((const uint32_t *)d)[i]; }
size_t ucc_coll_args_get_total_count(const ucc_coll_args_t *a, const ucc_count_t *counts, ucc_rank_t size);
size_t ucc_coll_args_get_max_count(const ucc_coll_args_t *a, const ucc_count_t *counts, ucc_rank_t size);
/* message size in bytes used for score-map lookup; UCC_MSG_MAX-style 0 for asymmetric v-colls */
size_t ucc_coll_args_msgsize(const ucc_coll_args_t *a, ucc_rank_t rank, ucc_rank_t size);
ucc_memory_type_t ucc_coll_args_mem_type(const ucc_coll_args_t *a, ucc_rank_t rank);

const char *ucc_coll_type_str(ucc_coll_type_t ct);
ucc_coll_type_t ucc_coll_type_from_str(const char *s); /* UCC_COLL_TYPE_LAST on error */
const char *ucc_mem_type_str(ucc_memory_type_t mt);
ucc_memory_type_t ucc_mem_type_from_str(const char *s); /* UCC_MEMORY_TYPE_LAST on error */
const char *ucc_reduction_op_str(ucc_reduction_op_t op);
ucc_reduction_op_t ucc_reduction_op_from_str(const char *s);
const char *ucc_thread_mode_str(ucc_thread_mode_t tm);
void ucc_coll_args_str(const ucc_coll_args_t *a, ucc_rank_t rank, ucc_rank_t size, char *buf, size_t len);
void ucc_mem_type_mask_str(uint32_t mask, char *buf, size_t len);
void ucc_coll_types_mask_str(uint64_t mask, char *buf, size_t len);
#endif
