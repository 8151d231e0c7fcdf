/* Scatter-reduce / allgather (SRA) block math for radix-k recursive exchange (capability of reference
 * coll_patterns/sra_knomial.h:17-474; this is an independent, much smaller formulation).
 * The vector of `count` elements is owned in nested segments: before the round with digit weight `dist` a rank owns
 * [seg_off, seg_off+seg_cnt); the round splits it into `radix` nearly equal parts and the rank keeps part number
 * digit = (rank / dist) % radix.  Running the rounds with dist = n_full/radix ... 1 gives the reduce-scatter,
 * the reverse order the allgather. */
#ifndef UCC_SRA_KNOMIAL_H_
#define UCC_SRA_KNOMIAL_H_
#include "knomial_tree.h"
#include "utils/ucc_math.h"

typedef struct ucc_sra_seg { size_t off, cnt; } ucc_sra_seg_t;

/* part `j` (0..radix-1) of a segment split into radix nearly equal pieces */
static inline ucc_sra_seg_t ucc_sra_part(ucc_sra_seg_t s, unsigned radix, unsigned j)
{
    ucc_sra_seg_t p;
    p.off = s.off + ucc_buffer_block_offset(s.cnt, radix, j);
    p.cnt = ucc_buffer_block_count(s.cnt, radix, j);
    return p;
}
/* segment a (base/proxy) rank owns after the scatter-reduce rounds with weights n_full/radix, ..., `dist` (inclusive);
 * dist == n_full returns the whole vector */
static inline ucc_sra_seg_t ucc_sra_owned(const ucc_kn_pattern_t *p, size_t count, uint64_t dist)
{
    ucc_sra_seg_t s = {0, count};
    for (uint64_t d = p->n_full / p->radix; d >= 1 && d >= dist; d /= p->radix) {
        s = ucc_sra_part(s, p->radix, (unsigned)((p->rank / d) % p->radix));
        if (d == 1) break;
    }
    return s;
}
/* final block of `rank` after the complete reduce-scatter */
static inline ucc_sra_seg_t ucc_sra_final(const ucc_kn_pattern_t *p, size_t count) { return ucc_sra_owned(p, count, 1); }
#endif
