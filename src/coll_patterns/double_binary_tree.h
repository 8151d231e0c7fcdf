/* Double binary tree: two spanning binary trees over the same ranks such
 * that the two halves of a message can flow through different trees.
 * Tree A is a balanced in-order tree on [0,N); tree B is the same shape on
 * ranks rotated by one, so interior nodes of A tend to be leaves of B.
 * (Role of reference coll_patterns/double_binary_tree.h; construction differs.) */
#ifndef UCC_DOUBLE_BINARY_TREE_H_
#define UCC_DOUBLE_BINARY_TREE_H_
#include "utils/ucc_compiler_def.h"

typedef struct ucc_dbt { ucc_rank_t parent[2]; ucc_rank_t children[2][2]; } ucc_dbt_t;

static inline void ucc_bt_inorder_(ucc_rank_t rank, ucc_rank_t size, ucc_rank_t *parent, ucc_rank_t *c0, ucc_rank_t *c1)
{
    ucc_rank_t lo = 0, hi = size, p = UCC_RANK_INVALID;
    for (;;) {
        ucc_rank_t mid = lo + (hi - lo) / 2;
        if (rank == mid) {
            *parent = p;
            *c0 = (mid > lo) ? lo + (mid - lo) / 2 : UCC_RANK_INVALID;
            *c1 = (hi > mid + 1) ? (mid + 1) + (hi - mid - 1) / 2 : UCC_RANK_INVALID;
            return;
        }
        p = mid;
        if (rank < mid) hi = mid; else lo = mid + 1;
    }
}
static inline void ucc_dbt_init(ucc_dbt_t *t, ucc_rank_t rank, ucc_rank_t size)
{
    ucc_rank_t p, c0, c1, sr = (rank + 1) % size;
    ucc_bt_inorder_(rank, size, &t->parent[0], &t->children[0][0], &t->children[0][1]);
    ucc_bt_inorder_(sr, size, &p, &c0, &c1);
    t->parent[1]      = p  == UCC_RANK_INVALID ? p  : (p + size - 1) % size;
    t->children[1][0] = c0 == UCC_RANK_INVALID ? c0 : (c0 + size - 1) % size;
    t->children[1][1] = c1 == UCC_RANK_INVALID ? c1 : (c1 + size - 1) % size;
}
static inline void ucc_dbt_roots(ucc_rank_t size, ucc_rank_t *r1, ucc_rank_t *r2)
{ *r1 = size / 2; *r2 = (size / 2 + size - 1) % size; }
#endif
