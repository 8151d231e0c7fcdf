/* mc/cpu: host memory component — malloc + pool of fixed-size buffers for
 * the frequent small scratch allocations (role of reference mc/cpu/mc_cpu.c). */
#include "components/mc/base/ucc_mc_base.h"
#include "utils/ucc_mpool.h"
#include "utils/ucc_math.h"

typedef struct ucc_mc_cpu_config { ucc_mc_config_t super; size_t mpool_elem_size; unsigned mpool_max_elems; } ucc_mc_cpu_config_t;
typedef struct ucc_mc_cpu { ucc_mc_base_t super; ucc_mpool_t mpool; int mpool_init; ucc_thread_mode_t tm; ucc_spinlock_t init_lock; } ucc_mc_cpu_t;
extern ucc_mc_cpu_t ucc_mc_cpu;
#define MC_CPU_CFG ((ucc_mc_cpu_config_t *)ucc_mc_cpu.super.config)

static ucc_config_field_t ucc_mc_cpu_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_mc_cpu_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_mc_config_table)},
    {"MPOOL_ELEM_SIZE", "1Mb", "The size of each element in mc cpu mpool", ucc_offsetof(ucc_mc_cpu_config_t, mpool_elem_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"MPOOL_MAX_ELEMS", "8", "The max amount of elements in mc cpu mpool", ucc_offsetof(ucc_mc_cpu_config_t, mpool_max_elems), UCC_CONFIG_TYPE_UINT},
    {NULL}};

static ucc_status_t mc_cpu_init(const ucc_mc_params_t *p) { ucc_mc_cpu.tm = p->thread_mode; ucc_mc_cpu.mpool_init = 0; ucc_spinlock_init(&ucc_mc_cpu.init_lock); return UCC_OK; }
static ucc_status_t mc_cpu_get_attr(ucc_mc_attr_t *a)
{ if (a->field_mask & UCC_MC_ATTR_FIELD_FAST_ALLOC_SIZE) a->fast_alloc_size = MC_CPU_CFG->mpool_max_elems ? MC_CPU_CFG->mpool_elem_size : 0; return UCC_OK; }

static ucc_status_t mc_cpu_alloc_plain(ucc_mc_buffer_header_t **h_ptr, size_t size)
{
    ucc_mc_buffer_header_t *h = (ucc_mc_buffer_header_t *)malloc(sizeof(*h) + 64 + size);
    if (!h) return UCC_ERR_NO_MEMORY;
    h->from_pool = 0; h->mt = UCC_MEMORY_TYPE_HOST;
    h->addr = (void *)ucc_align_up_pow2((uintptr_t)(h + 1), (uintptr_t)64);
    *h_ptr = h;
    return UCC_OK;
}
static void pool_obj_init(ucc_mpool_t *mp, void *obj, void *chunk)
{
    ucc_mc_buffer_header_t *h = (ucc_mc_buffer_header_t *)obj;
    (void)mp; (void)chunk;
    h->from_pool = 1; h->mt = UCC_MEMORY_TYPE_HOST; h->addr = (void *)ucc_align_up_pow2((uintptr_t)(h + 1), (uintptr_t)64);
}
static ucc_status_t mc_cpu_mem_alloc(ucc_mc_buffer_header_t **h_ptr, size_t size, ucc_memory_type_t mt)
{
    (void)mt;
    if (MC_CPU_CFG->mpool_max_elems && size <= MC_CPU_CFG->mpool_elem_size) {
        if (!ucc_mc_cpu.mpool_init) {
            static ucc_mpool_ops_t ops = {NULL, NULL, pool_obj_init, NULL};
            ucc_spin_lock(&ucc_mc_cpu.init_lock);
            if (!ucc_mc_cpu.mpool_init) {
                ucc_mpool_init(&ucc_mc_cpu.mpool, 0, sizeof(ucc_mc_buffer_header_t) + 64 + MC_CPU_CFG->mpool_elem_size, 0, 64, 1,
                               MC_CPU_CFG->mpool_max_elems, &ops, ucc_mc_cpu.tm, "mc cpu mpool");
                ucc_mc_cpu.mpool_init = 1;
            }
            ucc_spin_unlock(&ucc_mc_cpu.init_lock);
        }
        ucc_mc_buffer_header_t *h = (ucc_mc_buffer_header_t *)ucc_mpool_get(&ucc_mc_cpu.mpool);
        if (h) { *h_ptr = h; return UCC_OK; }
    }
    return mc_cpu_alloc_plain(h_ptr, size);
}
static ucc_status_t mc_cpu_mem_free(ucc_mc_buffer_header_t *h) { if (h->from_pool) ucc_mpool_put(h); else free(h); return UCC_OK; }
static ucc_status_t mc_cpu_memcpy(void *dst, const void *src, size_t len, ucc_memory_type_t d, ucc_memory_type_t s)
{ (void)d; (void)s; memcpy(dst, src, len); return UCC_OK; }
static ucc_status_t mc_cpu_memset(void *dst, int v, size_t len) { memset(dst, v, len); return UCC_OK; }
static ucc_status_t mc_cpu_mem_query(const void *ptr, ucc_mem_attr_t *a)
{
    a->mem_type = UCC_MEMORY_TYPE_HOST;
    if (a->field_mask & UCC_MEM_ATTR_FIELD_BASE_ADDRESS) a->base_address = (void *)ptr;
    if (a->field_mask & UCC_MEM_ATTR_FIELD_ALLOC_LENGTH) a->alloc_length = 1;
    return UCC_OK;
}
static ucc_status_t mc_cpu_finalize(void) { if (ucc_mc_cpu.mpool_init) { ucc_mpool_cleanup(&ucc_mc_cpu.mpool, 1); ucc_mc_cpu.mpool_init = 0; } return UCC_OK; }

ucc_mc_cpu_t ucc_mc_cpu = {
    .super = {.super = {.name = "cpu", .score = 1}, .ref_cnt = 0, .ee_type = UCC_EE_CPU_THREAD, .type = UCC_MEMORY_TYPE_HOST,
              .config_table = {"CPU memory component", "MC_CPU_", ucc_mc_cpu_config_table, sizeof(ucc_mc_cpu_config_t), {NULL, NULL}},
              .init = mc_cpu_init, .get_attr = mc_cpu_get_attr, .finalize = mc_cpu_finalize,
              .ops = {mc_cpu_mem_query, mc_cpu_mem_alloc, mc_cpu_mem_free, mc_cpu_memcpy, mc_cpu_memset, NULL}}};
UCC_COMPONENT_REGISTER_STATIC(mc, ucc_mc_cpu)
static void UCC_CTOR mc_cpu_cfg_register(void) { ucc_config_table_register(&ucc_mc_cpu.super.config_table); }
