/* k-nomial tree and recursive k-ing helpers shared by host algorithms.
 * (Role of reference coll_patterns/recursive_knomial.h; formulation here is
 * digit based: a rank's parent is found by clearing its lowest non-zero
 * base-k digit, its children by setting lower digits.) */
#ifndef UCC_KNOMIAL_TREE_H_
#define UCC_KNOMIAL_TREE_H_
#include "utils/ucc_compiler_def.h"

#define UCC_KN_MAX_CHILDREN 128

typedef struct ucc_kn_tree {
    ucc_rank_t parent;                      /* UCC_RANK_INVALID for the root */
    ucc_rank_t children[UCC_KN_MAX_CHILDREN]; /* ordered: largest subtree first */
    unsigned   n_children;
} ucc_kn_tree_t;

/* tree over `size` ranks rooted at `root`; all ranks are team ranks */
static inline void ucc_kn_tree_init(ucc_kn_tree_t *t, ucc_rank_t rank, ucc_rank_t size, ucc_rank_t root, unsigned radix)
{
    ucc_rank_t vr = (rank + size - root) % size, tmp[UCC_KN_MAX_CHILDREN];
    unsigned n = 0;
    uint64_t dist;
    if (radix < 2) radix = 2;
    t->parent = UCC_RANK_INVALID;
    for (dist = 1; dist < size; dist *= radix) {
        if (vr % (dist * radix) != 0) { t->parent = (ucc_rank_t)((vr - vr % (dist * radix) + root) % size); break; }
        for (unsigned j = 1; j < radix; j++) {
            uint64_t c = vr + j * dist;
            if (c < size && n < UCC_KN_MAX_CHILDREN) tmp[n++] = (ucc_rank_t)((c + root) % size);
        }
    }
    t->n_children = n;
    for (unsigned i = 0; i < n; i++) t->children[i] = tmp[n - 1 - i];
}

/* number of ranks in the subtree of virtual rank `vr` (itself included) */
static inline ucc_rank_t ucc_kn_subtree_size(ucc_rank_t vr, ucc_rank_t size, unsigned radix)
{
    uint64_t span = 1;
    if (vr == 0) return size;
    while (vr % (span * radix) == 0) span *= radix;
    return (ucc_rank_t)((vr + span <= size) ? span : size - vr);
}

/* recursive k-ing participation classes for non power-of-radix sizes */
typedef enum { UCC_KN_NODE_BASE, UCC_KN_NODE_PROXY, UCC_KN_NODE_EXTRA } ucc_kn_node_type_t;
typedef struct ucc_kn_pattern {
    unsigned   radix;
    ucc_rank_t size, rank, n_full; /* n_full = radix^p <= size */
    ucc_kn_node_type_t type;
    ucc_rank_t partner;            /* EXTRA: its proxy; PROXY: first extra */
    unsigned   n_extras;           /* PROXY: extras served: rank + j*n_full, j = 1..n_extras (< radix) */
} ucc_kn_pattern_t;

static inline void ucc_kn_pattern_init(ucc_kn_pattern_t *p, ucc_rank_t rank, ucc_rank_t size, unsigned radix)
{
    uint64_t f = 1;
    if (radix < 2) radix = 2;
    if (radix > size && size >= 2) radix = size;
    while (f * radix <= size) f *= radix;
    p->radix = radix; p->size = size; p->rank = rank; p->n_full = (ucc_rank_t)f; p->n_extras = 0;
    /* size < radix * n_full, so every base rank serves at most radix-1 extras */
    if (rank >= p->n_full) { p->type = UCC_KN_NODE_EXTRA; p->partner = rank % p->n_full; }
    else {
        for (uint64_t e = (uint64_t)rank + p->n_full; e < size; e += p->n_full) p->n_extras++;
        if (p->n_extras) { p->type = UCC_KN_NODE_PROXY; p->partner = rank + p->n_full; }
        else { p->type = UCC_KN_NODE_BASE; p->partner = UCC_RANK_INVALID; }
    }
}
#define ucc_kn_extra(_p, _j) ((ucc_rank_t)((_p)->rank + ((_j) + 1) * (_p)->n_full))
/* peers of `rank` in the round with digit weight `dist` (excluding itself); returns count */
static inline unsigned ucc_kn_round_peers(const ucc_kn_pattern_t *p, uint64_t dist, ucc_rank_t *peers)
{
    ucc_rank_t base = (ucc_rank_t)(p->rank - ((p->rank / dist) % p->radix) * dist);
    unsigned n = 0;
    for (unsigned j = 0; j < p->radix; j++) { ucc_rank_t q = (ucc_rank_t)(base + j * dist); if (q != p->rank && q < p->n_full) peers[n++] = q; }
    return n;
}
#endif
