/* Bulk asynchronous copies (the TMA engine of sm_90+/sm_100: cp.async.bulk) as the data mover of the copy-shaped collectives.
 *
 * A thread-copy loop keeps an SM's LSU busy with 16-byte transactions and needs hundreds of resident threads per SM to cover
 * the NVLink round trip; a bulk copy is ONE instruction per 8..32 KB: global -> shared memory (completion counted on an
 * mbarrier) and shared -> global (tracked by bulk groups), executed by the copy engine next to the SM while its issue slots
 * stay free.  One elected thread per CTA drives a ring of NVL_BULK_STAGES shared-memory buffers, so a handful of one-warp CTAs
 * keeps megabytes in flight - the collective leaves almost the whole GPU to the application's compute kernels.
 * Works on any global address: local HBM, a peer's memory mapped over NVLink (loads and stores), and stores into the
 * multicast mapping.  Requires 16-byte aligned addresses and sizes. */
#ifndef UCC_TL_NVL_BULK_CUH_
#define UCC_TL_NVL_BULK_CUH_
#include "nvl_device.cuh"

#define NVL_BULK_STAGES 8
#define NVL_BULK_AHEAD  4                 /* loads issued ahead of the matching store */
#define NVL_BULK_STAGE_BYTES (24 * 1024)  /* 8 x 24 KB = 192 KB of the 227 KB a CTA may use */
#define NVL_BULK_SMEM (NVL_BULK_STAGES * NVL_BULK_STAGE_BYTES + 128)

#ifndef NVL_HOST_EMU
NVL_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
NVL_DEV void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
NVL_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
NVL_DEV void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
NVL_DEV void bulk_g2s(void *smem, const void *gsrc, uint32_t bytes, uint64_t *bar)
{ asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory"); }
NVL_DEV void bulk_s2g(void *gdst, const void *smem, uint32_t bytes)
{ asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem)), "r"(bytes) : "memory"); }
NVL_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> NVL_DEV void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
NVL_DEV void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
NVL_DEV void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
NVL_DEV void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

/* state of the one driving thread: ring position survives across bulk_copy_range() calls of a kernel */
struct BulkPipe {
    char     *buf;      /* NVL_BULK_STAGES x NVL_BULK_STAGE_BYTES of shared memory, 128-byte aligned */
    uint64_t *bar;      /* one mbarrier per stage */
    uint32_t  issued;   /* chunks whose load has been issued so far (ring position = issued % STAGES) */
    NVL_DEV void init(char *smem_base)
    {
        buf = smem_base; bar = reinterpret_cast<uint64_t *>(smem_base + NVL_BULK_STAGES * NVL_BULK_STAGE_BYTES); issued = 0;
        for (int s = 0; s < NVL_BULK_STAGES; s++) mbar_init(&bar[s], 1);
        fence_mbar_init();
    }
};

/* copy [off0, off1) of src to the same offsets of dst in chunks this CTA owns: chunk c (of NVL_BULK_STAGE_BYTES) belongs to
 * CTA (c + rot) % nctas.  Called by ONE thread.  Loads run NVL_BULK_AHEAD chunks ahead of the stores. */
NVL_DEV void bulk_copy_range(BulkPipe &pp, char *dst, const char *src, size_t bytes, int cta, int nctas, int rot)
{
    const size_t nchunks = (bytes + NVL_BULK_STAGE_BYTES - 1) / NVL_BULK_STAGE_BYTES;
    size_t first = (size_t)((cta - rot % nctas + nctas) % nctas);
    size_t mine = first < nchunks ? (nchunks - first + nctas - 1) / nctas : 0;   /* chunks first, first + nctas, ... */
    for (size_t i = 0; i < mine + NVL_BULK_AHEAD; i++) {
        if (i < mine) {
            const size_t c = first + i * nctas, o = c * NVL_BULK_STAGE_BYTES;
            const uint32_t n = (uint32_t)(bytes - o < NVL_BULK_STAGE_BYTES ? bytes - o : NVL_BULK_STAGE_BYTES);
            const uint32_t s = pp.issued % NVL_BULK_STAGES;
            /* the store that last read this stage was issued NVL_BULK_STAGES - NVL_BULK_AHEAD iterations ago: at most
             * STAGES - AHEAD - 1 newer stores may still be reading their (other) stages */
            bulk_wait_read<NVL_BULK_STAGES - NVL_BULK_AHEAD - 1>();
            mbar_expect_tx(&pp.bar[s], n);
            bulk_g2s(pp.buf + (size_t)s * NVL_BULK_STAGE_BYTES, src + o, n, &pp.bar[s]);
            pp.issued++;
        }
        if (i >= NVL_BULK_AHEAD) {
            const size_t k = i - NVL_BULK_AHEAD, c = first + k * nctas, o = c * NVL_BULK_STAGE_BYTES;
            const uint32_t n = (uint32_t)(bytes - o < NVL_BULK_STAGE_BYTES ? bytes - o : NVL_BULK_STAGE_BYTES);
            const uint32_t pos = pp.issued - (uint32_t)((i < mine ? i + 1 : mine) - k);  /* ring position chunk k was loaded into */
            const uint32_t s = pos % NVL_BULK_STAGES, parity = (pos / NVL_BULK_STAGES) & 1;
            mbar_wait(&pp.bar[s], parity);
            bulk_s2g(dst + o, pp.buf + (size_t)s * NVL_BULK_STAGE_BYTES, n);
            bulk_commit();
        } else bulk_commit(); /* keep one group per iteration so the wait_group arithmetic above holds from the start */
    }
}
#endif /* NVL_HOST_EMU */
#endif
