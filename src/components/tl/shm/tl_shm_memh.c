/* tl/shm memory registration: ucc_mem_map() support of the host transport (reference: tl/ucp/tl_ucp_context.c:506-577 registers the
 * segment with UCX and packs an rkey that peers put to / get from).
 *
 * Host memory needs no registration to be reachable on one node - a peer in the same process dereferences the pointer, a peer in
 * another process uses cross-memory attach (process_vm_readv / process_vm_writev) - so a handle is simply the description
 * {host, pid, address, length} of the segment.  What the registration buys is that a one-sided algorithm knows every member's
 * buffer at INIT time: the put-based alltoall (tl_shm_coll.c a2a_onesided, UCC_COLL_ARGS_FLAG_DST_MEMH_GLOBAL) writes its blocks
 * straight into the members' registered destinations without the address-exchange round of the get-based variant.
 * Convention (the reference's): the buffer of a collective sits at the same offset inside the registered segment on every member. */
#include "tl_shm.h"
#include "components/mc/ucc_mc.h"
#include <unistd.h>

ucc_status_t ucc_tl_shm_mem_map(const ucc_base_context_t *b_ctx, ucc_mem_map_mode_t mode, void *memh, void *tl_h_p)
{
    ucc_tl_shm_context_t *ctx = ucc_derived_of(b_ctx, ucc_tl_shm_context_t);
    shm_memh_t *h;
    if (mode == UCC_MEM_MAP_MODE_EXPORT) {
        const ucc_mem_map_t *seg = (const ucc_mem_map_t *)memh;
        ucc_mem_attr_t attr;
        if (!seg || !seg->address || !seg->len) return UCC_ERR_INVALID_PARAM;
        memset(&attr, 0, sizeof(attr)); attr.field_mask = UCC_MEM_ATTR_FIELD_MEM_TYPE;
        if (ucc_mc_get_mem_attr(seg->address, &attr) == UCC_OK && attr.mem_type != UCC_MEMORY_TYPE_HOST) return UCC_ERR_NOT_SUPPORTED; /* device memory: tl/nvl's business */
        h = (shm_memh_t *)calloc(1, sizeof(*h));
        if (!h) return UCC_ERR_NO_MEMORY;
        h->magic = UCC_TL_SHM_MEMH_MAGIC; h->host_hash = ctx->addr.host_hash; h->pid = ctx->addr.pid; h->addr = (uint64_t)(uintptr_t)seg->address; h->len = seg->len;
        *(void **)tl_h_p = h;
        return UCC_OK;
    }
    if (mode == UCC_MEM_MAP_MODE_IMPORT) {
        shm_memh_t in;
        memcpy(&in, memh, sizeof(in));
        if (in.magic != UCC_TL_SHM_MEMH_MAGIC) return UCC_ERR_INVALID_PARAM;
        h = (shm_memh_t *)calloc(1, sizeof(*h));
        if (!h) return UCC_ERR_NO_MEMORY;
        *h = in; h->imported = 1;
        h->reachable = in.host_hash == ctx->addr.host_hash;   /* same node: pointer or CMA (checked per endpoint when the collective is built) */
        *(void **)tl_h_p = h;
        return UCC_OK;
    }
    return UCC_ERR_NOT_SUPPORTED;
}

ucc_status_t ucc_tl_shm_mem_unmap(const ucc_base_context_t *b_ctx, ucc_mem_map_mode_t mode, void *tl_h)
{
    shm_memh_t *h = (shm_memh_t *)tl_h;
    (void)b_ctx; (void)mode;
    if (!h || h->magic != UCC_TL_SHM_MEMH_MAGIC) return UCC_ERR_INVALID_PARAM;
    h->magic = 0;
    free(h);
    return UCC_OK;
}

ucc_status_t ucc_tl_shm_memh_pack(const ucc_base_context_t *b_ctx, ucc_mem_map_mode_t mode, void *tl_h, void **pack_buffer)
{
    shm_memh_t *h = (shm_memh_t *)tl_h, *out;
    char *buf;
    (void)b_ctx; (void)mode;
    if (!h || h->magic != UCC_TL_SHM_MEMH_MAGIC) return UCC_ERR_INVALID_PARAM;
    buf = (char *)calloc(1, sizeof(size_t) + sizeof(shm_memh_t)); /* {size_t len; bytes} as the core expects */
    if (!buf) return UCC_ERR_NO_MEMORY;
    *(size_t *)buf = sizeof(shm_memh_t);
    out = (shm_memh_t *)(buf + sizeof(size_t));
    *out = *h; out->imported = 0; out->reachable = 0;
    *pack_buffer = buf;
    return UCC_OK;
}
