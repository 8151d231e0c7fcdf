/* Ring pattern helpers: neighbours, block walk, and a topology-aware ring
 * order for GPU teams (visit ranks so that consecutive ring members share
 * the most NVLinks; on NVSwitch systems every order is equivalent and the
 * identity is returned).  Role of reference coll_patterns/ring.{h,c}. */
#ifndef UCC_RING_H_
#define UCC_RING_H_
#include "utils/ucc_compiler_def.h"
static inline ucc_rank_t ucc_ring_next(ucc_rank_t r, ucc_rank_t n) { return (r + 1) % n; }
static inline ucc_rank_t ucc_ring_prev(ucc_rank_t r, ucc_rank_t n) { return (r + n - 1) % n; }
/* block sent at `step` by rank r in a reduce-scatter ring where rank r finally owns block (r+shift)%n */
static inline ucc_rank_t ucc_ring_rs_send_block(ucc_rank_t r, ucc_rank_t n, ucc_rank_t step, int shift) { return (r + 2 * n - step - (shift ? 0 : 1)) % n; }
static inline ucc_rank_t ucc_ring_rs_recv_block(ucc_rank_t r, ucc_rank_t n, ucc_rank_t step, int shift) { return (r + 2 * n - step - 1 - (shift ? 0 : 1)) % n; }
/* greedy nearest-neighbour ring over a link-count matrix (n x n, row-major); order[] receives the permutation */
static inline void ucc_ring_build_from_links(const uint8_t *links, ucc_rank_t n, ucc_rank_t *order)
{
    uint8_t used[64] = {0};
    ucc_rank_t cur = 0;
    order[0] = 0; used[0] = 1;
    for (ucc_rank_t k = 1; k < n && n <= 64; k++) {
        ucc_rank_t best = UCC_RANK_INVALID; int best_l = -1;
        for (ucc_rank_t j = 0; j < n; j++) if (!used[j] && (int)links[cur * n + j] > best_l) { best_l = links[cur * n + j]; best = j; }
        order[k] = best; used[best] = 1; cur = best;
    }
}
#endif
