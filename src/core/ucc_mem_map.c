/* ucc_mem_map / ucc_mem_unmap: registration handles for one-sided / zero-copy collectives.
 * EXPORT asks every TL context that implements mem_map to register the segment and concatenates
 * {tl_name[8], size, blob} records into one relocatable handle; IMPORT re-creates per-TL state from
 * such a handle on the receiving side (layout contract: reference core/ucc_context.c:1212-1571). */
#include "ucc_context.h"
#include "utils/ucc_log.h"

#define UCC_MEMH_MAGIC 0x55434d454d480001ull
typedef struct ucc_memh_tl_rec { char tl_name[8]; uint64_t size; } ucc_memh_tl_rec_t;
typedef struct ucc_mem_map_memh {
    uint64_t          magic;
    uint32_t          mode;
    uint32_t          n_tls;
    void             *address;
    size_t            len;
    ucc_context_t    *ctx;       /* local owner (not meaningful after transfer) */
    void            **tl_h;      /* local per-TL handles, parallel to the records */
    uint64_t          pack_size; /* bytes of records that follow */
    char              pack[1];
} ucc_mem_map_memh_t;

UCC_EXPORT ucc_status_t ucc_mem_map(ucc_context_h ctx, ucc_mem_map_mode_t mode, const ucc_mem_map_params_t *params, size_t *memh_size, ucc_mem_map_mem_h *memh_p)
{
    ucc_mem_map_memh_t *h;
    size_t total = 0;
    void **packs; size_t *sizes; void **tlh; int n = 0;
    if (!ctx || !memh_p || mode >= UCC_MEM_MAP_MODE_LAST) return UCC_ERR_INVALID_PARAM;
    if (mode == UCC_MEM_MAP_MODE_EXPORT_OFFLOAD || mode == UCC_MEM_MAP_MODE_IMPORT_OFFLOAD) return UCC_ERR_NOT_SUPPORTED;
    packs = (void **)calloc((size_t)ctx->n_tl_ctx + 1, sizeof(void *)); sizes = (size_t *)calloc((size_t)ctx->n_tl_ctx + 1, sizeof(size_t)); tlh = (void **)calloc((size_t)ctx->n_tl_ctx + 1, sizeof(void *));
    if (mode == UCC_MEM_MAP_MODE_EXPORT) {
        if (!params || params->n_segments != 1 || !params->segments) { free(packs); free(sizes); free(tlh); ucc_error("ucc_mem_map export needs exactly one segment"); return UCC_ERR_INVALID_PARAM; }
        for (int i = 0; i < ctx->n_tl_ctx; i++) {
            ucc_tl_iface_t *iface = UCC_TL_CTX_IFACE(ctx->tl_ctx[i]);
            if (!iface->context.mem_map || !iface->context.memh_pack) continue;
            if (iface->context.mem_map(&ctx->tl_ctx[i]->super, mode, (void *)params->segments, &tlh[i]) != UCC_OK) continue;
            if (iface->context.memh_pack(&ctx->tl_ctx[i]->super, mode, tlh[i], &packs[i]) != UCC_OK) { packs[i] = NULL; continue; }
            sizes[i] = packs[i] ? *(size_t *)packs[i] : 0; /* TLs return {size_t len; bytes...} */
            total += sizeof(ucc_memh_tl_rec_t) + sizes[i]; n++;
        }
        h = (ucc_mem_map_memh_t *)calloc(1, sizeof(*h) + total);
        if (!h) { free(packs); free(sizes); free(tlh); return UCC_ERR_NO_MEMORY; }
        h->magic = UCC_MEMH_MAGIC; h->mode = mode; h->n_tls = (uint32_t)n; h->address = params->segments[0].address; h->len = params->segments[0].len;
        h->ctx = ctx; h->tl_h = tlh; h->pack_size = total;
        { char *p = h->pack;
          for (int i = 0; i < ctx->n_tl_ctx; i++) {
              ucc_memh_tl_rec_t rec;
              if (!packs[i]) continue;
              memset(&rec, 0, sizeof(rec)); strncpy(rec.tl_name, UCC_TL_CTX_IFACE(ctx->tl_ctx[i])->super.name, 7); rec.size = sizes[i];
              memcpy(p, &rec, sizeof(rec)); p += sizeof(rec);
              memcpy(p, (char *)packs[i] + sizeof(size_t), sizes[i]); p += sizes[i];
              free(packs[i]);
          } }
        free(packs); free(sizes);
        if (memh_size) *memh_size = sizeof(*h) + total;
        *memh_p = h;
        return UCC_OK;
    }
    /* IMPORT: *memh_p points at a handle received from a peer */
    free(packs); free(sizes);
    h = (ucc_mem_map_memh_t *)*memh_p;
    if (!h || h->magic != UCC_MEMH_MAGIC) { free(tlh); ucc_error("ucc_mem_map import: not a memory handle"); return UCC_ERR_INVALID_PARAM; }
    h->ctx = ctx; h->tl_h = tlh; h->mode = mode;
    { char *p = h->pack; char *end = h->pack + h->pack_size;
      while (p < end) {
          ucc_memh_tl_rec_t rec; memcpy(&rec, p, sizeof(rec)); p += sizeof(rec);
          for (int i = 0; i < ctx->n_tl_ctx; i++) {
              ucc_tl_iface_t *iface = UCC_TL_CTX_IFACE(ctx->tl_ctx[i]);
              if (strncmp(rec.tl_name, iface->super.name, 7) || !iface->context.mem_map) continue;
              iface->context.mem_map(&ctx->tl_ctx[i]->super, mode, p, &tlh[i]);
          }
          p += rec.size;
      } }
    if (memh_size) *memh_size = sizeof(*h) + h->pack_size;
    return UCC_OK;
}

/* the per-TL part of a handle (what that TL's mem_map returned), for TLs that use registered buffers in their collectives */
void *ucc_mem_map_tl_handle(ucc_mem_map_mem_h memh, const ucc_base_context_t *tl_ctx)
{
    ucc_mem_map_memh_t *h = (ucc_mem_map_memh_t *)memh;
    if (!h || h->magic != UCC_MEMH_MAGIC || !h->ctx || !h->tl_h) return NULL;
    for (int i = 0; i < h->ctx->n_tl_ctx; i++) if (&h->ctx->tl_ctx[i]->super == tl_ctx) return h->tl_h[i];
    return NULL;
}

UCC_EXPORT ucc_status_t ucc_mem_unmap(ucc_mem_map_mem_h *memh_p)
{
    ucc_mem_map_memh_t *h;
    if (!memh_p || !*memh_p) return UCC_ERR_INVALID_PARAM;
    h = (ucc_mem_map_memh_t *)*memh_p;
    if (h->magic != UCC_MEMH_MAGIC) return UCC_ERR_INVALID_PARAM;
    if (h->ctx && h->tl_h) {
        for (int i = 0; i < h->ctx->n_tl_ctx; i++) {
            ucc_tl_iface_t *iface = UCC_TL_CTX_IFACE(h->ctx->tl_ctx[i]);
            if (h->tl_h[i] && iface->context.mem_unmap) iface->context.mem_unmap(&h->ctx->tl_ctx[i]->super, (ucc_mem_map_mode_t)h->mode, h->tl_h[i]);
        }
        free(h->tl_h);
    }
    if (h->mode == UCC_MEM_MAP_MODE_EXPORT) free(h); /* imported handles live in caller memory */
    *memh_p = NULL;
    return UCC_OK;
}
