/* mc/cuda: device memory component — pointer classification, pooled scratch
 * allocations, synchronous memcpy/memset on a private stream.
 * Role of reference mc/cuda/mc_cuda.c (693 LoC); loaded as a plugin module so
 * the core library has no CUDA link dependency. */
#include "components/mc/base/ucc_mc_base.h"
#include "utils/ucc_mpool.h"
#include "utils/ucc_math.h"
#include "utils/cuda/ucc_cuda_util.h"

typedef struct ucc_mc_cuda_config {
    ucc_mc_config_t super;
    size_t          mpool_elem_size;
    unsigned        mpool_max_elems;
    int             use_managed_detection;
} ucc_mc_cuda_config_t;

typedef struct ucc_mc_cuda {
    ucc_mc_base_t     super;
    ucc_thread_mode_t tm;
    ucc_mpool_t       mpool;      /* headers + pre-allocated device chunks */
    int               mpool_init;
    int               mpool_dev;
    cudaStream_t      stream[UCC_CUDA_MAX_DEVICES];
    ucc_spinlock_t    lock;
} ucc_mc_cuda_t;
extern ucc_mc_cuda_t ucc_mc_cuda;
#define MC_CFG ((ucc_mc_cuda_config_t *)ucc_mc_cuda.super.config)
#define mc_error(_f, ...) ucc_log_component(UCC_LOG_LEVEL_ERROR, &ucc_mc_cuda.super.config->log_component, _f, ##__VA_ARGS__)
#define mc_debug(_f, ...) ucc_log_component(UCC_LOG_LEVEL_DEBUG, &ucc_mc_cuda.super.config->log_component, _f, ##__VA_ARGS__)

static ucc_config_field_t ucc_mc_cuda_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_mc_cuda_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_mc_config_table)},
    {"MPOOL_ELEM_SIZE", "1Mb", "The size of each element in mc cuda mpool", ucc_offsetof(ucc_mc_cuda_config_t, mpool_elem_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"MPOOL_MAX_ELEMS", "8", "The max amount of elements in mc cuda mpool", ucc_offsetof(ucc_mc_cuda_config_t, mpool_max_elems), UCC_CONFIG_TYPE_UINT},
    {NULL}};

static ucc_status_t mc_cuda_init(const ucc_mc_params_t *p)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { (void)cudaGetLastError(); mc_debug("cuda devices are not found"); return UCC_ERR_NO_RESOURCE; }
    ucc_mc_cuda.tm = p->thread_mode; ucc_mc_cuda.mpool_init = 0;
    memset(ucc_mc_cuda.stream, 0, sizeof(ucc_mc_cuda.stream));
    ucc_spinlock_init(&ucc_mc_cuda.lock);
    return UCC_OK;
}
static ucc_status_t mc_cuda_get_attr(ucc_mc_attr_t *a)
{ if (a->field_mask & UCC_MC_ATTR_FIELD_FAST_ALLOC_SIZE) a->fast_alloc_size = MC_CFG->mpool_max_elems ? MC_CFG->mpool_elem_size : 0; return UCC_OK; }

static ucc_status_t get_stream(cudaStream_t *s)
{
    int dev;
    CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= UCC_CUDA_MAX_DEVICES) return UCC_ERR_NO_RESOURCE;
    if (!ucc_mc_cuda.stream[dev]) {
        ucc_spin_lock(&ucc_mc_cuda.lock);
        if (!ucc_mc_cuda.stream[dev]) {
            cudaError_t e = cudaStreamCreateWithFlags(&ucc_mc_cuda.stream[dev], cudaStreamNonBlocking);
            if (e != cudaSuccess) { ucc_spin_unlock(&ucc_mc_cuda.lock); mc_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); return UCC_ERR_NO_RESOURCE; }
        }
        ucc_spin_unlock(&ucc_mc_cuda.lock);
    }
    *s = ucc_mc_cuda.stream[dev];
    return UCC_OK;
}

static ucc_status_t alloc_plain(ucc_mc_buffer_header_t **h_ptr, size_t size, ucc_memory_type_t mt)
{
    ucc_mc_buffer_header_t *h = (ucc_mc_buffer_header_t *)malloc(sizeof(*h));
    cudaError_t e;
    if (!h) return UCC_ERR_NO_MEMORY;
    e = mt == UCC_MEMORY_TYPE_CUDA_MANAGED ? cudaMallocManaged(&h->addr, size, cudaMemAttachGlobal) : cudaMalloc(&h->addr, size);
    if (e != cudaSuccess) { (void)cudaGetLastError(); mc_error("failed to allocate %zu bytes of cuda memory: %s", size, cudaGetErrorString(e)); free(h); return UCC_ERR_NO_MEMORY; }
    h->from_pool = 0; h->mt = mt;
    *h_ptr = h;
    return UCC_OK;
}
/* pool objects are headers; the device chunk is allocated when the object is created */
static void pool_obj_init(ucc_mpool_t *mp, void *obj, void *chunk)
{
    ucc_mc_buffer_header_t *h = (ucc_mc_buffer_header_t *)obj;
    (void)mp; (void)chunk;
    if (cudaMalloc(&h->addr, MC_CFG->mpool_elem_size) != cudaSuccess) { (void)cudaGetLastError(); h->addr = NULL; }
    h->from_pool = 1; h->mt = UCC_MEMORY_TYPE_CUDA;
}
static void pool_obj_cleanup(ucc_mpool_t *mp, void *obj)
{ ucc_mc_buffer_header_t *h = (ucc_mc_buffer_header_t *)obj; (void)mp; if (h->addr) cudaFree(h->addr); }

static ucc_status_t mc_cuda_mem_alloc(ucc_mc_buffer_header_t **h_ptr, size_t size, ucc_memory_type_t mt)
{
    if (mt == UCC_MEMORY_TYPE_CUDA && MC_CFG->mpool_max_elems && size <= MC_CFG->mpool_elem_size) {
        int dev = -1;
        cudaGetDevice(&dev);
        if (!ucc_mc_cuda.mpool_init) {
            static ucc_mpool_ops_t ops = {NULL, NULL, pool_obj_init, pool_obj_cleanup};
            ucc_spin_lock(&ucc_mc_cuda.lock);
            if (!ucc_mc_cuda.mpool_init) {
                ucc_mpool_init(&ucc_mc_cuda.mpool, 0, sizeof(ucc_mc_buffer_header_t), 0, 8, 1, MC_CFG->mpool_max_elems, &ops, ucc_mc_cuda.tm, "mc cuda mpool");
                ucc_mc_cuda.mpool_init = 1; ucc_mc_cuda.mpool_dev = dev;
            }
            ucc_spin_unlock(&ucc_mc_cuda.lock);
        }
        if (dev == ucc_mc_cuda.mpool_dev) {
            ucc_mc_buffer_header_t *h = (ucc_mc_buffer_header_t *)ucc_mpool_get(&ucc_mc_cuda.mpool);
            if (h && h->addr) { *h_ptr = h; return UCC_OK; }
            if (h) ucc_mpool_put(h);
        }
    }
    return alloc_plain(h_ptr, size, mt);
}
static ucc_status_t mc_cuda_mem_free(ucc_mc_buffer_header_t *h)
{
    if (h->from_pool) { ucc_mpool_put(h); return UCC_OK; }
    CUDA_CHECK(cudaFree(h->addr));
    free(h);
    return UCC_OK;
}
static ucc_status_t mc_cuda_memcpy(void *dst, const void *src, size_t len, ucc_memory_type_t dmt, ucc_memory_type_t smt)
{
    cudaStream_t s;
    (void)dmt; (void)smt;
    UCC_CHECK_RET(get_stream(&s));
    CUDA_CHECK(cudaMemcpyAsync(dst, src, len, cudaMemcpyDefault, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    return UCC_OK;
}
static ucc_status_t mc_cuda_memset(void *dst, int v, size_t len)
{
    cudaStream_t s;
    UCC_CHECK_RET(get_stream(&s));
    CUDA_CHECK(cudaMemsetAsync(dst, v, len, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    return UCC_OK;
}
static ucc_status_t mc_cuda_mem_query(const void *ptr, ucc_mem_attr_t *a)
{
    struct cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return UCC_ERR_NOT_SUPPORTED; }
    switch (attr.type) {
    case cudaMemoryTypeDevice: a->mem_type = UCC_MEMORY_TYPE_CUDA; break;
    case cudaMemoryTypeManaged: a->mem_type = UCC_MEMORY_TYPE_CUDA_MANAGED; break;
    default: return UCC_ERR_NOT_SUPPORTED; /* host / unregistered: let mc/cpu answer */
    }
    if (a->field_mask & (UCC_MEM_ATTR_FIELD_BASE_ADDRESS | UCC_MEM_ATTR_FIELD_ALLOC_LENGTH)) {
        void *base = NULL; size_t len = 0;
        if (ucc_cuda_get_address_range(ptr, &base, &len) != UCC_OK) { base = (void *)ptr; len = 1; }
        a->base_address = base; a->alloc_length = len;
    }
    return UCC_OK;
}
static ucc_status_t mc_cuda_flush(void) { return UCC_OK; /* NVLink-only library: no GDR write flush needed */ }
static ucc_status_t mc_cuda_finalize(void)
{
    if (ucc_mc_cuda.mpool_init) { ucc_mpool_cleanup(&ucc_mc_cuda.mpool, 0); ucc_mc_cuda.mpool_init = 0; }
    for (int d = 0; d < UCC_CUDA_MAX_DEVICES; d++) if (ucc_mc_cuda.stream[d]) { cudaStreamDestroy(ucc_mc_cuda.stream[d]); ucc_mc_cuda.stream[d] = NULL; }
    return UCC_OK;
}

ucc_mc_cuda_t ucc_mc_cuda = {
    .super = {.super = {.name = "cuda", .score = 2}, .ref_cnt = 0, .ee_type = UCC_EE_CUDA_STREAM, .type = UCC_MEMORY_TYPE_CUDA,
              .config_table = {"CUDA memory component", "MC_CUDA_", ucc_mc_cuda_config_table, sizeof(ucc_mc_cuda_config_t), {NULL, NULL}},
              .init = mc_cuda_init, .get_attr = mc_cuda_get_attr, .finalize = mc_cuda_finalize,
              .ops = {mc_cuda_mem_query, mc_cuda_mem_alloc, mc_cuda_mem_free, mc_cuda_memcpy, mc_cuda_memset, mc_cuda_flush}}};
/* plugin entry symbol: ucc_<framework>_<name> */
UCC_EXPORT ucc_mc_cuda_t *ucc_mc_cuda_sym_keepalive = &ucc_mc_cuda;
static void UCC_CTOR mc_cuda_cfg_register(void) { ucc_config_table_register(&ucc_mc_cuda.super.config_table); }
