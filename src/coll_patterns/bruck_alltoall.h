/* Bruck alltoall index math (capability of reference coll_patterns/bruck_alltoall.h:18-42; formulation is own).
 * After the local rotation (block i of rank r holds the data for rank (r+i)%N), step k = 0..ceil(log2 N)-1 sends every
 * block whose index has bit k set to rank (r + 2^k) % N and receives the same indices from rank (r - 2^k + N) % N. */
#ifndef UCC_BRUCK_ALLTOALL_H_
#define UCC_BRUCK_ALLTOALL_H_
#include "utils/ucc_compiler_def.h"
#include "utils/ucc_coll_utils.h"

static inline unsigned ucc_bruck_n_steps(ucc_rank_t size) { unsigned s = 0; while (((ucc_rank_t)1 << s) < size) s++; return s; }
static inline ucc_rank_t ucc_bruck_send_peer(ucc_rank_t rank, ucc_rank_t size, unsigned step) { return (ucc_rank_t)((rank + ((uint64_t)1 << step)) % size); }
static inline ucc_rank_t ucc_bruck_recv_peer(ucc_rank_t rank, ucc_rank_t size, unsigned step) { return (ucc_rank_t)((rank + size - (((uint64_t)1 << step) % size)) % size); }
/* number of blocks exchanged in `step`, and their indices (ascending) into idx[] when it is not NULL */
static inline ucc_rank_t ucc_bruck_step_blocks(ucc_rank_t size, unsigned step, ucc_rank_t *idx)
{
    ucc_rank_t n = 0;
    for (ucc_rank_t i = 1; i < size; i++) if (i & ((ucc_rank_t)1 << step)) { if (idx) idx[n] = i; n++; }
    return n;
}
/* where the block that ends up at local index i (after all steps) has to go in the user's receive buffer */
static inline ucc_rank_t ucc_bruck_final_src(ucc_rank_t rank, ucc_rank_t size, ucc_rank_t i) { return (ucc_rank_t)((rank + size - i) % size); }
#endif
