#include "ucc_coll_utils.h"
#include "ucc_string.h"
#include <stdio.h>
#include <strings.h>

/* ---------------- ep maps ---------------- */
ucc_rank_t ucc_ep_map_eval(ucc_ep_map_t map, ucc_rank_t rank)
{
    switch ((int)map.type) {
    case UCC_EP_MAP_FULL: return rank;
    case UCC_EP_MAP_STRIDED: return (ucc_rank_t)((int64_t)map.strided.start + (int64_t)rank * map.strided.stride);
    case UCC_EP_MAP_ARRAY:
        return map.array.elem_size == 8 ? (ucc_rank_t)((uint64_t *)map.array.map)[rank] : ((uint32_t *)map.array.map)[rank];
    case UCC_EP_MAP_CB: return (ucc_rank_t)map.cb.cb(rank, map.cb.cb_ctx);
    case UCC_EP_MAP_REVERSE_INTERNAL: return (ucc_rank_t)(map.ep_num - 1 - rank);
    default: return UCC_RANK_INVALID;
    }
}
ucc_ep_map_t ucc_ep_map_create_full(ucc_rank_t size)
{ ucc_ep_map_t m; memset(&m, 0, sizeof(m)); m.type = UCC_EP_MAP_FULL; m.ep_num = size; return m; }
ucc_ep_map_t ucc_ep_map_create_reverse(ucc_rank_t size)
{ ucc_ep_map_t m; memset(&m, 0, sizeof(m)); m.type = UCC_EP_MAP_REVERSE_INTERNAL; m.ep_num = size; return m; }

#define DEFINE_FROM_ARRAY(_name, _T, _esz)                                                              \
    ucc_ep_map_t _name(_T **array, ucc_rank_t size, ucc_rank_t full_size, int need_free)                 \
    {                                                                                                    \
        ucc_ep_map_t m; int strided = 1; int64_t stride = size > 1 ? (int64_t)(*array)[1] - (int64_t)(*array)[0] : 1; \
        memset(&m, 0, sizeof(m)); m.ep_num = size;                                                       \
        for (ucc_rank_t i = 2; i < size && strided; i++)                                                 \
            if ((int64_t)(*array)[i] - (int64_t)(*array)[i - 1] != stride) strided = 0;                  \
        if (size > 0 && strided) {                                                                       \
            if (stride == 1 && (*array)[0] == 0 && size == full_size) m.type = UCC_EP_MAP_FULL;          \
            else { m.type = UCC_EP_MAP_STRIDED; m.strided.start = (*array)[0]; m.strided.stride = stride; } \
            if (need_free) { free(*array); *array = NULL; }                                              \
        } else { m.type = UCC_EP_MAP_ARRAY; m.array.map = *array; m.array.elem_size = _esz; }            \
        return m;                                                                                        \
    }
DEFINE_FROM_ARRAY(ucc_ep_map_from_array, ucc_rank_t, 4)
DEFINE_FROM_ARRAY(ucc_ep_map_from_array_64, uint64_t, 8)

ucc_status_t ucc_ep_map_create_inverse(ucc_ep_map_t map, ucc_ep_map_t *inv, int need_free_array)
{
    ucc_rank_t  n = (ucc_rank_t)map.ep_num, max = 0;
    ucc_rank_t *arr;
    for (ucc_rank_t i = 0; i < n; i++) { ucc_rank_t r = ucc_ep_map_eval(map, i); if (r > max) max = r; }
    arr = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * ((size_t)max + 1));
    if (!arr) return UCC_ERR_NO_MEMORY;
    for (ucc_rank_t i = 0; i <= max; i++) arr[i] = UCC_RANK_INVALID;
    for (ucc_rank_t i = 0; i < n; i++) arr[ucc_ep_map_eval(map, i)] = i;
    *inv = ucc_ep_map_from_array(&arr, max + 1, max + 1, need_free_array);
    return UCC_OK;
}
ucc_status_t ucc_ep_map_create_nested(ucc_ep_map_t *base, ucc_ep_map_t *sub, ucc_ep_map_t *out)
{
    ucc_rank_t  n = (ucc_rank_t)sub->ep_num;
    ucc_rank_t *arr = (ucc_rank_t *)malloc(sizeof(ucc_rank_t) * (n ? n : 1));
    ucc_rank_t  full = 0;
    if (!arr) return UCC_ERR_NO_MEMORY;
    for (ucc_rank_t i = 0; i < n; i++) arr[i] = ucc_ep_map_eval(*base, ucc_ep_map_eval(*sub, i));
    for (ucc_rank_t i = 0; i < (ucc_rank_t)base->ep_num; i++) { ucc_rank_t r = ucc_ep_map_eval(*base, i); if (r + 1 > full) full = r + 1; }
    *out = ucc_ep_map_from_array(&arr, n, full, 1);
    return UCC_OK;
}
void ucc_ep_map_destroy(ucc_ep_map_t *map)
{ if (map->type == UCC_EP_MAP_ARRAY) { free(map->array.map); map->array.map = NULL; } }
void ucc_ep_map_destroy_nested(ucc_ep_map_t *map) { ucc_ep_map_destroy(map); }
ucc_rank_t ucc_ep_map_local_rank(ucc_ep_map_t map, ucc_rank_t ep)
{
    ucc_rank_t n = (ucc_rank_t)map.ep_num;
    if (map.type == UCC_EP_MAP_FULL) return ep < n ? ep : UCC_RANK_INVALID;
    if (map.type == UCC_EP_MAP_STRIDED && map.strided.stride != 0) {
        int64_t d = (int64_t)ep - (int64_t)map.strided.start;
        if (d % map.strided.stride) return UCC_RANK_INVALID;
        d /= map.strided.stride;
        return (d >= 0 && d < (int64_t)n) ? (ucc_rank_t)d : UCC_RANK_INVALID;
    }
    for (ucc_rank_t i = 0; i < n; i++) if (ucc_ep_map_eval(map, i) == ep) return i;
    return UCC_RANK_INVALID;
}
int ucc_ep_map_is_identity(const ucc_ep_map_t *map)
{
    if (map->type == UCC_EP_MAP_FULL) return 1;
    if (map->type == UCC_EP_MAP_STRIDED) return map->strided.start == 0 && map->strided.stride == 1;
    for (ucc_rank_t i = 0; i < (ucc_rank_t)map->ep_num; i++) if (ucc_ep_map_eval(*map, i) != i) return 0;
    return 1;
}

/* ---------------- counts / sizes ---------------- */
size_t ucc_coll_args_get_total_count(const ucc_coll_args_t *a, const ucc_count_t *counts, ucc_rank_t size)
{ size_t t = 0; for (ucc_rank_t i = 0; i < size; i++) t += ucc_coll_args_get_count(a, counts, i); return t; }
size_t ucc_coll_args_get_max_count(const ucc_coll_args_t *a, const ucc_count_t *counts, ucc_rank_t size)
{ size_t m = 0; for (ucc_rank_t i = 0; i < size; i++) { size_t c = ucc_coll_args_get_count(a, counts, i); if (c > m) m = c; } return m; }

size_t ucc_coll_args_msgsize(const ucc_coll_args_t *a, ucc_rank_t rank, ucc_rank_t size)
{
    int is_root = (ucc_rank_t)a->root == rank;
    switch (a->coll_type) {
    case UCC_COLL_TYPE_BARRIER: case UCC_COLL_TYPE_FANIN: case UCC_COLL_TYPE_FANOUT: return 0;
    case UCC_COLL_TYPE_BCAST: return a->src.info.count * ucc_dt_size(a->src.info.datatype);
    case UCC_COLL_TYPE_ALLREDUCE: case UCC_COLL_TYPE_ALLTOALL: case UCC_COLL_TYPE_ALLGATHER:
        return a->dst.info.count * ucc_dt_size(a->dst.info.datatype);
    case UCC_COLL_TYPE_REDUCE_SCATTER:
        /* dst.count is the per-rank result unless in-place (then it is the total) */
        return UCC_IS_INPLACE(*a) ? a->dst.info.count * ucc_dt_size(a->dst.info.datatype)
                                  : a->dst.info.count * ucc_dt_size(a->dst.info.datatype) * size;
    case UCC_COLL_TYPE_REDUCE:
        return is_root ? a->dst.info.count * ucc_dt_size(a->dst.info.datatype) : a->src.info.count * ucc_dt_size(a->src.info.datatype);
    case UCC_COLL_TYPE_GATHER:
        return is_root ? a->dst.info.count * ucc_dt_size(a->dst.info.datatype)
                       : a->src.info.count * ucc_dt_size(a->src.info.datatype) * size;
    case UCC_COLL_TYPE_SCATTER:
        return is_root ? a->src.info.count * ucc_dt_size(a->src.info.datatype)
                       : a->dst.info.count * ucc_dt_size(a->dst.info.datatype) * size;
    case UCC_COLL_TYPE_ALLGATHERV: case UCC_COLL_TYPE_REDUCE_SCATTERV:
        return ucc_coll_args_get_total_count(a, a->dst.info_v.counts, size) * ucc_dt_size(a->dst.info_v.datatype);
    case UCC_COLL_TYPE_ALLTOALLV: case UCC_COLL_TYPE_GATHERV: case UCC_COLL_TYPE_SCATTERV:
    default: return 0; /* asymmetric: local knowledge is not enough for a consistent choice */
    }
}
ucc_memory_type_t ucc_coll_args_mem_type(const ucc_coll_args_t *a, ucc_rank_t rank)
{
    int is_root = (ucc_rank_t)a->root == rank;
    switch (a->coll_type) {
    case UCC_COLL_TYPE_BARRIER: case UCC_COLL_TYPE_FANIN: case UCC_COLL_TYPE_FANOUT: return UCC_MEMORY_TYPE_HOST;
    case UCC_COLL_TYPE_BCAST: return a->src.info.mem_type;
    case UCC_COLL_TYPE_ALLREDUCE: case UCC_COLL_TYPE_ALLTOALL: case UCC_COLL_TYPE_ALLGATHER: case UCC_COLL_TYPE_REDUCE_SCATTER:
        return a->dst.info.mem_type;
    case UCC_COLL_TYPE_ALLGATHERV: case UCC_COLL_TYPE_ALLTOALLV: case UCC_COLL_TYPE_REDUCE_SCATTERV: return a->dst.info_v.mem_type;
    case UCC_COLL_TYPE_REDUCE: case UCC_COLL_TYPE_GATHER: return is_root ? a->dst.info.mem_type : a->src.info.mem_type;
    case UCC_COLL_TYPE_GATHERV: return is_root ? a->dst.info_v.mem_type : a->src.info.mem_type;
    case UCC_COLL_TYPE_SCATTER: return is_root ? a->src.info.mem_type : a->dst.info.mem_type;
    case UCC_COLL_TYPE_SCATTERV: return is_root ? a->src.info_v.mem_type : a->dst.info.mem_type;
    default: return UCC_MEMORY_TYPE_UNKNOWN;
    }
}

/* ---------------- names ---------------- */
static const char *coll_names[UCC_COLL_TYPE_NUM] = {
    "allgather", "allgatherv", "allreduce", "alltoall", "alltoallv", "barrier", "bcast", "fanin", "fanout",
    "gather", "gatherv", "reduce", "reduce_scatter", "reduce_scatterv", "scatter", "scatterv"};
const char *ucc_coll_type_str(ucc_coll_type_t ct)
{ int i = ucc_coll_type_index(ct); return (ct && ucc_is_pow2((uint64_t)ct) && i < UCC_COLL_TYPE_NUM) ? coll_names[i] : "unknown"; }
ucc_coll_type_t ucc_coll_type_from_str(const char *s)
{ for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) if (!strcasecmp(s, coll_names[i])) return (ucc_coll_type_t)UCC_BIT(i); return UCC_COLL_TYPE_LAST; }
static const char *mt_names[] = {"host", "cuda", "cuda_managed", "rocm", "rocm_managed", "unknown"};
const char *ucc_mem_type_str(ucc_memory_type_t mt) { return mt <= UCC_MEMORY_TYPE_LAST ? mt_names[mt] : "invalid"; }
ucc_memory_type_t ucc_mem_type_from_str(const char *s)
{
    if (!strcasecmp(s, "cpu")) return UCC_MEMORY_TYPE_HOST;
    if (!strcasecmp(s, "cudamanaged") || !strcasecmp(s, "cuda-mng")) return UCC_MEMORY_TYPE_CUDA_MANAGED;
    if (!strcasecmp(s, "rocmmanaged")) return UCC_MEMORY_TYPE_ROCM_MANAGED;
    for (int i = 0; i < UCC_MEMORY_TYPE_LAST; i++) if (!strcasecmp(s, mt_names[i])) return (ucc_memory_type_t)i;
    return UCC_MEMORY_TYPE_LAST;
}
static const char *op_names[UCC_OP_LAST] = {"sum", "prod", "max", "min", "land", "lor", "lxor", "band", "bor", "bxor", "maxloc", "minloc", "avg"};
const char *ucc_reduction_op_str(ucc_reduction_op_t op) { return op < UCC_OP_LAST ? op_names[op] : "unknown"; }
ucc_reduction_op_t ucc_reduction_op_from_str(const char *s)
{ for (int i = 0; i < UCC_OP_LAST; i++) if (!strcasecmp(s, op_names[i])) return (ucc_reduction_op_t)i; return UCC_OP_LAST; }
const char *ucc_thread_mode_str(ucc_thread_mode_t tm)
{ return tm == UCC_THREAD_SINGLE ? "single" : tm == UCC_THREAD_FUNNELED ? "funneled" : tm == UCC_THREAD_MULTIPLE ? "multiple" : "unknown"; }

void ucc_mem_type_mask_str(uint32_t mask, char *buf, size_t len)
{
    size_t o = 0; buf[0] = 0;
    for (int i = 0; i < UCC_MEMORY_TYPE_LAST && o < len; i++) if (mask & (1u << i)) o += snprintf(buf + o, len - o, "%s%s", o ? "," : "", mt_names[i]);
}
void ucc_coll_types_mask_str(uint64_t mask, char *buf, size_t len)
{
    size_t o = 0; buf[0] = 0;
    for (int i = 0; i < UCC_COLL_TYPE_NUM && o < len; i++) if (mask & UCC_BIT(i)) o += snprintf(buf + o, len - o, "%s%s", o ? "," : "", coll_names[i]);
}

void ucc_coll_args_str(const ucc_coll_args_t *a, ucc_rank_t rank, ucc_rank_t size, char *buf, size_t len)
{
    size_t o = 0;
    ucc_coll_type_t ct = a->coll_type;
    int has_src = 1, has_dst = 1, vsrc = 0, vdst = 0;
    o += snprintf(buf + o, len - o, "%s", ucc_coll_type_str(ct));
    if (ucc_coll_has_reduction(ct)) o += snprintf(buf + o, len - o, " %s", ucc_reduction_op_str(a->op));
    if (ucc_coll_is_rooted(ct)) o += snprintf(buf + o, len - o, " root %lu", (unsigned long)a->root);
    if (UCC_IS_INPLACE(*a)) o += snprintf(buf + o, len - o, " inplace");
    if (UCC_IS_PERSISTENT(*a)) o += snprintf(buf + o, len - o, " persistent");
    if (ct & (UCC_COLL_TYPE_BARRIER | UCC_COLL_TYPE_FANIN | UCC_COLL_TYPE_FANOUT)) return;
    if (ct == UCC_COLL_TYPE_BCAST) has_dst = 0;
    if (ct & (UCC_COLL_TYPE_ALLTOALLV | UCC_COLL_TYPE_SCATTERV)) vsrc = 1;
    if (ct & (UCC_COLL_TYPE_ALLTOALLV | UCC_COLL_TYPE_ALLGATHERV | UCC_COLL_TYPE_GATHERV | UCC_COLL_TYPE_REDUCE_SCATTERV)) vdst = 1;
    if (ucc_coll_is_rooted(ct) && (ucc_rank_t)a->root != rank) {
        if (ct & (UCC_COLL_TYPE_REDUCE | UCC_COLL_TYPE_GATHER | UCC_COLL_TYPE_GATHERV)) has_dst = 0;
        if (ct & (UCC_COLL_TYPE_SCATTER | UCC_COLL_TYPE_SCATTERV)) has_src = 0;
    }
    if (UCC_IS_INPLACE(*a) && !(ucc_coll_is_rooted(ct) && (ucc_rank_t)a->root != rank) && ct != UCC_COLL_TYPE_BCAST) has_src = 0;
    if (has_src && o < len) {
        if (vsrc) o += snprintf(buf + o, len - o, ": src={%p, total %zu, %s, %s}", a->src.info_v.buffer,
                                ucc_coll_args_get_total_count(a, a->src.info_v.counts, size), ucc_datatype_str(a->src.info_v.datatype),
                                ucc_mem_type_str(a->src.info_v.mem_type));
        else o += snprintf(buf + o, len - o, ": src={%p, %lu, %s, %s}", a->src.info.buffer, (unsigned long)a->src.info.count,
                           ucc_datatype_str(a->src.info.datatype), ucc_mem_type_str(a->src.info.mem_type));
    }
    if (has_dst && o < len) {
        if (vdst) o += snprintf(buf + o, len - o, "%s dst={%p, total %zu, %s, %s}", has_src ? "," : ":", a->dst.info_v.buffer,
                                ucc_coll_args_get_total_count(a, a->dst.info_v.counts, size), ucc_datatype_str(a->dst.info_v.datatype),
                                ucc_mem_type_str(a->dst.info_v.mem_type));
        else o += snprintf(buf + o, len - o, "%s dst={%p, %lu, %s, %s}", has_src ? "," : ":", a->dst.info.buffer,
                           (unsigned long)a->dst.info.count, ucc_datatype_str(a->dst.info.datatype), ucc_mem_type_str(a->dst.info.mem_type));
    }
}
