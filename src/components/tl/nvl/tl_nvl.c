#include "tl_nvl.h"
#include "core/ucc_global_opts.h"
#include "utils/ucc_sys.h"
#include <unistd.h>

static ucc_config_field_t tl_nvl_lib_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_tl_nvl_lib_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_tl_lib_config_table)}, {NULL}};

ucc_config_field_t ucc_tl_nvl_context_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_tl_nvl_context_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_tl_context_config_table)},
    {"SYMMETRIC_SIZE", "384Mb", "Size of the data region of the per-team symmetric heap; larger messages are processed in rounds inside one kernel "
     "(384 MB = three 128 MB round buffers of the pipelined NVLS allreduce, the fastest geometry measured on 8 x B200)",
     ucc_offsetof(ucc_tl_nvl_context_config_t, symmetric_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"NBLOCKS", "auto", "Thread blocks per collective kernel (auto: chosen from the message size)", ucc_offsetof(ucc_tl_nvl_context_config_t, nblocks), UCC_CONFIG_TYPE_UINT},
    {"MAX_BLOCKS", "256", "Upper bound of thread blocks per collective kernel (all blocks of all ranks sharing a GPU must be co-resident)",
     ucc_offsetof(ucc_tl_nvl_context_config_t, max_blocks), UCC_CONFIG_TYPE_UINT},
    {"NTHREADS", "512", "Threads per block", ucc_offsetof(ucc_tl_nvl_context_config_t, nthreads), UCC_CONFIG_TYPE_UINT},
    {"TIMEOUT", "10s", "Spin budget of a device-side wait before the kernel gives up and the collective fails with UCC_ERR_TIMED_OUT",
     ucc_offsetof(ucc_tl_nvl_context_config_t, timeout), UCC_CONFIG_TYPE_TIME},
    {"USE_NVLS", "try", "Use NVSwitch multicast / in-switch reduction (multimem.*) when the fabric supports it", ucc_offsetof(ucc_tl_nvl_context_config_t, use_nvls), UCC_CONFIG_TYPE_TERNARY},
    {"USE_VMM", "try", "Allocate the heap with the CUDA virtual memory management API and share it as a POSIX fd (required for NVLS); otherwise cudaMalloc + cudaIpc",
     ucc_offsetof(ucc_tl_nvl_context_config_t, use_vmm), UCC_CONFIG_TYPE_TERNARY},
    {"ALLREDUCE_ONESHOT_THRESH", "1M", "Allreduce messages up to (and including) this size use the one-shot push kernel", ucc_offsetof(ucc_tl_nvl_context_config_t, oneshot_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {"ZCOPY", "try", "Read and write the members' user buffers in place (CUDA IPC handles exchanged per collective, mappings cached) instead of staging through the heap",
     ucc_offsetof(ucc_tl_nvl_context_config_t, zcopy), UCC_CONFIG_TYPE_TERNARY},
    {"ZCOPY_THRESH", "1M", "Messages of at least this size use the zero-copy kernels", ucc_offsetof(ucc_tl_nvl_context_config_t, zcopy_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {"ALLREDUCE_NVLS_THRESH", "512M", "Allreduce messages of at least this size use NVLS when available", ucc_offsetof(ucc_tl_nvl_context_config_t, nvls_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {"USER_SIZE", "0", "Size of a symmetric USER region appended to every team heap (bound to the NVLS multicast object like the rest of the heap). "
     "Buffers placed there at the same offset on every member (ucc_tl_nvl_symm_region) are reduced in place through the switch: no staging, no copy-out",
     ucc_offsetof(ucc_tl_nvl_context_config_t, user_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"REQUIRE_NVLINK", "try", "Refuse teams in which some pair of GPUs has no NVLink path (such a team goes to the next TL): y = also when the topology "
     "could not be determined, try = only when the sysinfo graph says so, n = accept any pair with CUDA peer access", ucc_offsetof(ucc_tl_nvl_context_config_t, require_nvlink), UCC_CONFIG_TYPE_TERNARY},
    {"RING_REVERSE", "n", "Walk rings in the opposite direction (test knob for topology-ordered rings)", ucc_offsetof(ucc_tl_nvl_context_config_t, ring_reverse), UCC_CONFIG_TYPE_BOOL},
    {"BULK", "try", "Move the blocks of the zero-copy push exchange (allgather(v) / alltoall(v) algorithm `push`) with TMA bulk copies (cp.async.bulk through "
     "shared memory, one elected thread per one-warp thread block) instead of per-thread 16-byte loads and stores", ucc_offsetof(ucc_tl_nvl_context_config_t, bulk), UCC_CONFIG_TYPE_TERNARY},
    {"BULK_THRESH", "1M", "Total bytes from which the bulk-copy kernel is used", ucc_offsetof(ucc_tl_nvl_context_config_t, bulk_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {"BULK_CTAS", "16", "Thread blocks (one warp each, 192 KB of shared memory) of a bulk-copy kernel: the SM budget of the collective",
     ucc_offsetof(ucc_tl_nvl_context_config_t, bulk_ctas), UCC_CONFIG_TYPE_UINT},
    {"P2P_RNDV_THRESH", "1M", "Send / recv (two-member active-set bcast) of at least this size uses the rendezvous protocol: the receiver publishes its "
     "buffer, the sender's kernel stores straight into it over NVLink (no ring, no copy on the receiving GPU); smaller messages go through the "
     "pair's eager ring in the heap (two B200: the ring needs 8.7 us for 64 KB and 28.5 us for 1 MB, the rendezvous ~29 us for 1 MB, 48 us for 16 MB, "
     "209 us = 642 GB/s for 128 MB; NCCL 20 / 21.6 / 39.6 / 249 us)", ucc_offsetof(ucc_tl_nvl_context_config_t, p2p_rndv_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {"P2P_TIMEOUT", "0", "How long a send / recv kernel waits for the other side before it gives up and the request fails with UCC_ERR_TIMED_OUT; 0 = no "
     "deadline (a receive may legitimately be posted long before its send; NCCL waits forever, too).  TIMEOUT, the budget of the collectives, does not apply",
     ucc_offsetof(ucc_tl_nvl_context_config_t, p2p_timeout), UCC_CONFIG_TYPE_TIME},
    {"SLOTS", "1", "Independent collective lanes per team (reference tl/cuda MAX_CONCURRENT): consecutive collectives use consecutive lanes and, when "
     "posted on different streams, overlap.  Every lane has its own control block, one-shot slots and SYMMETRIC_SIZE of staging space; one "
     "kernel may use at most (2 x SMs) / SLOTS thread blocks so that all lanes stay co-resident", ucc_offsetof(ucc_tl_nvl_context_config_t, slots), UCC_CONFIG_TYPE_UINT},
    {"FD_VIA_PIDFD", "try", "Fetch peers' memory handles with pidfd_getfd before falling back to a unix socket", ucc_offsetof(ucc_tl_nvl_context_config_t, fd_via_pidfd), UCC_CONFIG_TYPE_TERNARY},
    {NULL}};

static ucc_status_t nvl_lib_init(const ucc_base_lib_params_t *p, const ucc_base_lib_config_t *config, ucc_base_lib_t **lib_p)
{
    ucc_tl_nvl_lib_t *lib = (ucc_tl_nvl_lib_t *)calloc(1, sizeof(*lib));
    int n = 0;
    (void)p;
    if (!lib) return UCC_ERR_NO_MEMORY;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { (void)cudaGetLastError(); free(lib); return UCC_ERR_NO_RESOURCE; }
    ucc_tl_lib_init_base(&lib->super, &ucc_tl_nvl, ucc_derived_of(config, ucc_tl_lib_config_t));
    *lib_p = &lib->super.super;
    return UCC_OK;
}
static void nvl_lib_finalize(ucc_base_lib_t *lib) { free(lib); }
static ucc_status_t nvl_lib_get_attr(const ucc_base_lib_t *lib, ucc_base_lib_attr_t *attr)
{ (void)lib; attr->attr.thread_mode = UCC_THREAD_MULTIPLE; attr->attr.coll_types = UCC_TL_NVL_SUPPORTED_COLLS; attr->flags = 0; attr->min_team_size = 1; attr->max_team_size = NVL_MAX_PEERS; return UCC_OK; }
static ucc_status_t nvl_lib_get_properties(ucc_base_lib_properties_t *p) { p->default_team_size = 2; p->min_team_size = 1; p->max_team_size = NVL_MAX_PEERS; return UCC_OK; }

static ucc_status_t nvl_ctx_create(const ucc_base_context_params_t *p, const ucc_base_ctx_config_t *config, ucc_base_context_t **ctx_p)
{
    static uint32_t seq = 0;
    ucc_tl_nvl_context_t *ctx;
    struct cudaDeviceProp prop;
    int dev, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); tl_debug(config->lib, "no current CUDA device"); return UCC_ERR_NO_RESOURCE; }
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { (void)cudaGetLastError(); return UCC_ERR_NO_RESOURCE; }
    if (cudaFree(0) != cudaSuccess) { (void)cudaGetLastError(); return UCC_ERR_NO_RESOURCE; }
    ctx = (ucc_tl_nvl_context_t *)calloc(1, sizeof(*ctx));
    if (!ctx) return UCC_ERR_NO_MEMORY;
    ctx->super.super.ucc_context = p->context; ctx->super.super.lib = config->lib; ctx->tm = p->thread_mode;
    ctx->n_ev_cache = 0; ucc_spinlock_init(&ctx->ev_lock);
    if (ucc_config_parser_clone_opts(config, &ctx->cfg, ucc_tl_nvl_context_config_table) != UCC_OK) { free(ctx); return UCC_ERR_NO_MEMORY; }
    ctx->dev = dev; ctx->sm_count = prop.multiProcessorCount;
    ctx->addr.host_hash = p->context->proc_info.host_hash; ctx->addr.pid = (int32_t)getpid(); ctx->addr.dev = dev;
    ctx->addr.ep_id = ((uint64_t)(uint32_t)ctx->addr.pid << 32) | ucc_atomic_fadd32(&seq, 1);
    ucc_cu_api_load();
    ctx->addr.vmm_ok = 0; ctx->addr.mc_ok = 0;
    { char bus[32]; unsigned dom = 0, b = 0, d = 0;
      ctx->addr.pci_domain = ctx->addr.pci_bus = ctx->addr.pci_device = -1;
      if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), dev) == cudaSuccess && sscanf(bus, "%x:%x:%x", &dom, &b, &d) == 3) { ctx->addr.pci_domain = (int32_t)dom; ctx->addr.pci_bus = (int32_t)b; ctx->addr.pci_device = (int32_t)d; }
      else (void)cudaGetLastError(); }
    if (ucc_cu.cuDeviceGetAttribute && ucc_cu.cuMemCreate) {
        CUdevice cudev;
        if (ucc_cu.cuDeviceGet(&cudev, dev) == CUDA_SUCCESS) {
            if (ucc_cu.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cudev) == CUDA_SUCCESS && v) ctx->addr.vmm_ok = 1;
            v = 0;
            if (ucc_cu.cuMulticastCreate && ucc_cu.cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev) == CUDA_SUCCESS && v) ctx->addr.mc_ok = 1;
        }
    }
    if (ctx->cfg.use_vmm == UCC_NO) ctx->addr.vmm_ok = 0;
    if (ctx->cfg.use_nvls == UCC_NO || !ctx->addr.vmm_ok) ctx->addr.mc_ok = 0;
    if (ctx->cfg.oneshot_thresh > NVL_LL_MAX) ctx->cfg.oneshot_thresh = NVL_LL_MAX;
    if (ctx->cfg.max_blocks > NVL_MAX_BLOCKS) ctx->cfg.max_blocks = NVL_MAX_BLOCKS;
    /* every block of a collective kernel must be resident at once (they wait for each other): 2 CTAs of 512 threads per SM */
    if (ctx->sm_count > 0 && ctx->cfg.max_blocks > 2u * (unsigned)ctx->sm_count) ctx->cfg.max_blocks = 2u * (unsigned)ctx->sm_count;
    if (ctx->cfg.slots < 1) ctx->cfg.slots = 1;
    if (ctx->cfg.slots > NVL_MAX_LANES) ctx->cfg.slots = NVL_MAX_LANES;
    /* kernels of different lanes may run at the same time and their blocks wait for the peers' blocks: everything that can be
     * in flight together has to fit the device (2 resident blocks per SM) */
    if (ctx->sm_count > 0 && ctx->cfg.max_blocks > 2u * (unsigned)ctx->sm_count / ctx->cfg.slots) ctx->cfg.max_blocks = 2u * (unsigned)ctx->sm_count / ctx->cfg.slots;
    if (ctx->cfg.max_blocks < 1) ctx->cfg.max_blocks = 1;
    if (ctx->cfg.nthreads < 64) ctx->cfg.nthreads = 64;
    if (ctx->cfg.nthreads > 1024) ctx->cfg.nthreads = 1024;
    ucc_mpool_init(&ctx->task_mp, 0, sizeof(ucc_tl_nvl_task_t), 0, 64, 16, (unsigned)-1, NULL, p->thread_mode, "tl_nvl_tasks");
    tl_debug(config->lib, "initialized tl context %p on device %d (%s, %d SMs): vmm %d multicast %d", (void *)ctx, dev, prop.name, ctx->sm_count, ctx->addr.vmm_ok, ctx->addr.mc_ok);
    *ctx_p = &ctx->super.super;
    return UCC_OK;
}
static void nvl_ctx_destroy(ucc_base_context_t *b)
{
    ucc_tl_nvl_context_t *ctx = ucc_derived_of(b, ucc_tl_nvl_context_t);
    ucc_mpool_cleanup(&ctx->task_mp, 1);
    for (unsigned i = 0; i < ctx->n_ev_cache; i++) cudaEventDestroy(ctx->ev_cache[i]);
    ucc_config_parser_release_opts(&ctx->cfg, ucc_tl_nvl_context_config_table);
    free(ctx);
}
static ucc_status_t nvl_ctx_get_attr(const ucc_base_context_t *b, ucc_base_ctx_attr_t *attr)
{
    const ucc_tl_nvl_context_t *ctx = ucc_derived_of(b, ucc_tl_nvl_context_t);
    if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN) attr->attr.ctx_addr_len = sizeof(ucc_tl_nvl_addr_t);
    if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR) memcpy(attr->attr.ctx_addr, &ctx->addr, sizeof(ctx->addr));
    if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_WORK_BUFFER_SIZE) attr->attr.global_work_buffer_size = 0;
    attr->topo_required = 1;
    return UCC_OK;
}

ucc_tl_iface_t ucc_tl_nvl = {
    .super = {.name = "nvl", .score = UCC_TL_NVL_DEFAULT_SCORE},
    .tl_lib_config = {"TL_NVL lib", "TL_NVL_", tl_nvl_lib_config_table, sizeof(ucc_tl_nvl_lib_config_t), {NULL, NULL}},
    .tl_context_config = {"TL_NVL context", "TL_NVL_", ucc_tl_nvl_context_config_table, sizeof(ucc_tl_nvl_context_config_t), {NULL, NULL}},
    .lib = {nvl_lib_init, nvl_lib_finalize, nvl_lib_get_attr, nvl_lib_get_properties},
    .context = {nvl_ctx_create, NULL, nvl_ctx_destroy, nvl_ctx_get_attr, ucc_tl_nvl_mem_map, ucc_tl_nvl_mem_unmap, ucc_tl_nvl_memh_pack},
    .team = {ucc_tl_nvl_team_create_post, ucc_tl_nvl_team_create_test, ucc_tl_nvl_team_destroy, ucc_tl_nvl_team_get_scores},
    .coll = {ucc_tl_nvl_coll_init},
};
static void UCC_CTOR tl_nvl_register(void)
{
    ucc_config_table_register(&ucc_tl_nvl.tl_lib_config); ucc_config_table_register(&ucc_tl_nvl.tl_context_config);
    ucc_tl_nvl_register_alg_info();
}
