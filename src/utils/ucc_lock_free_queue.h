/* Bounded multi-producer/multi-consumer queue of pointers used by the
 * lock-free progress queue (role of reference utils/ucc_lock_free_queue.h).
 * Design: Vyukov-style ring with per-cell sequence numbers; when the ring is
 * full, producers fall back to a spin-locked overflow list so enqueue never
 * fails. */
#ifndef UCC_LOCK_FREE_QUEUE_H_
#define UCC_LOCK_FREE_QUEUE_H_
#include "ucc_atomic.h"
#include "ucc_list.h"

#define UCC_LFQ_SIZE 1024u /* power of two */

typedef struct ucc_lf_queue_elem { ucc_list_link_t ovf; } ucc_lf_queue_elem_t;
typedef struct ucc_lf_cell { volatile uint64_t seq; void *ptr; } ucc_lf_cell_t;
typedef struct ucc_lf_queue {
    ucc_lf_cell_t     cells[UCC_LFQ_SIZE];
    volatile uint64_t head UCC_ALIGNED(64);
    volatile uint64_t tail UCC_ALIGNED(64);
    ucc_spinlock_t    ovf_lock;
    ucc_list_link_t   ovf;
    volatile int      ovf_len;
} ucc_lf_queue_t;

static inline void ucc_lf_queue_init(ucc_lf_queue_t *q)
{
    for (uint64_t i = 0; i < UCC_LFQ_SIZE; i++) { q->cells[i].seq = i; q->cells[i].ptr = NULL; }
    q->head = q->tail = 0; ucc_spinlock_init(&q->ovf_lock); ucc_list_head_init(&q->ovf); q->ovf_len = 0;
}
static inline void ucc_lf_queue_destroy(ucc_lf_queue_t *q) { (void)q; }

static inline int ucc_lf_queue_try_push_(ucc_lf_queue_t *q, void *p)
{
    uint64_t pos = ucc_load_acquire(&q->tail);
    for (;;) {
        ucc_lf_cell_t *c = &q->cells[pos & (UCC_LFQ_SIZE - 1)];
        uint64_t seq = ucc_load_acquire(&c->seq);
        int64_t  d = (int64_t)seq - (int64_t)pos;
        if (d == 0) {
            if (ucc_atomic_bool_cswap64(&q->tail, pos, pos + 1)) { c->ptr = p; ucc_store_release(&c->seq, pos + 1); return 1; }
            pos = ucc_load_acquire(&q->tail);
        } else if (d < 0) return 0; else pos = ucc_load_acquire(&q->tail);
    }
}
/* elem must embed ucc_lf_queue_elem_t at `e` */
static inline void ucc_lf_queue_enqueue(ucc_lf_queue_t *q, ucc_lf_queue_elem_t *e)
{
    if (ucc_likely(!q->ovf_len && ucc_lf_queue_try_push_(q, e))) return;
    ucc_spin_lock(&q->ovf_lock); ucc_list_add_tail(&q->ovf, &e->ovf); q->ovf_len++; ucc_spin_unlock(&q->ovf_lock);
}
static inline ucc_lf_queue_elem_t *ucc_lf_queue_dequeue(ucc_lf_queue_t *q)
{
    uint64_t pos = ucc_load_acquire(&q->head);
    for (;;) {
        ucc_lf_cell_t *c = &q->cells[pos & (UCC_LFQ_SIZE - 1)];
        uint64_t seq = ucc_load_acquire(&c->seq);
        int64_t  d = (int64_t)seq - (int64_t)(pos + 1);
        if (d == 0) {
            if (ucc_atomic_bool_cswap64(&q->head, pos, pos + 1)) {
                void *p = c->ptr; ucc_store_release(&c->seq, pos + UCC_LFQ_SIZE); return (ucc_lf_queue_elem_t *)p;
            }
            pos = ucc_load_acquire(&q->head);
        } else if (d < 0) break; else pos = ucc_load_acquire(&q->head);
    }
    if (q->ovf_len) {
        ucc_lf_queue_elem_t *e = NULL;
        ucc_spin_lock(&q->ovf_lock);
        if (!ucc_list_is_empty(&q->ovf)) { e = ucc_list_extract_head(&q->ovf, ucc_lf_queue_elem_t, ovf); q->ovf_len--; }
        ucc_spin_unlock(&q->ovf_lock);
        return e;
    }
    return NULL;
}
#endif
