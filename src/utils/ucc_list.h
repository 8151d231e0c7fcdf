/* Intrusive circular doubly-linked list + singly linked FIFO queue. */
#ifndef UCC_LIST_H_
#define UCC_LIST_H_
#include "ucc_compiler_def.h"

typedef struct ucc_list_link {
    struct ucc_list_link *prev, *next;
} ucc_list_link_t;

#define UCC_LIST_INITIALIZER(_h) { (_h), (_h) }
static inline void ucc_list_head_init(ucc_list_link_t *h) { h->prev = h->next = h; }
static inline int  ucc_list_is_empty(const ucc_list_link_t *h) { return h->next == h; }
static inline void ucc_list_insert_between_(ucc_list_link_t *p, ucc_list_link_t *n, ucc_list_link_t *e)
{ e->prev = p; e->next = n; p->next = e; n->prev = e; }
static inline void ucc_list_add_tail(ucc_list_link_t *h, ucc_list_link_t *e) { ucc_list_insert_between_(h->prev, h, e); }
static inline void ucc_list_add_head(ucc_list_link_t *h, ucc_list_link_t *e) { ucc_list_insert_between_(h, h->next, e); }
static inline void ucc_list_insert_before(ucc_list_link_t *pos, ucc_list_link_t *e) { ucc_list_insert_between_(pos->prev, pos, e); }
static inline void ucc_list_del(ucc_list_link_t *e) { e->prev->next = e->next; e->next->prev = e->prev; }
static inline unsigned long ucc_list_length(const ucc_list_link_t *h)
{ unsigned long n = 0; const ucc_list_link_t *e; for (e = h->next; e != h; e = e->next) n++; return n; }
static inline void ucc_list_splice_tail(ucc_list_link_t *h, ucc_list_link_t *other)
{
    if (ucc_list_is_empty(other)) return;
    other->next->prev = h->prev; h->prev->next = other->next;
    other->prev->next = h; h->prev = other->prev; ucc_list_head_init(other);
}
#define ucc_list_head(_h, _type, _m) ucc_container_of((_h)->next, _type, _m)
#define ucc_list_tail(_h, _type, _m) ucc_container_of((_h)->prev, _type, _m)
#define ucc_list_next(_e, _type, _m) ucc_container_of((_e)->next, _type, _m)
#define ucc_list_extract_head(_h, _type, _m) \
    ({ ucc_list_link_t *_l = (_h)->next; ucc_list_del(_l); ucc_container_of(_l, _type, _m); })
#define ucc_list_for_each(_e, _h, _m) \
    for (_e = ucc_container_of((_h)->next, __typeof__(*_e), _m); &(_e)->_m != (_h); \
         _e = ucc_container_of((_e)->_m.next, __typeof__(*_e), _m))
#define ucc_list_for_each_safe(_e, _t, _h, _m) \
    for (_e = ucc_container_of((_h)->next, __typeof__(*_e), _m), \
         _t = ucc_container_of((_e)->_m.next, __typeof__(*_e), _m); &(_e)->_m != (_h); \
         _e = _t, _t = ucc_container_of((_t)->_m.next, __typeof__(*_t), _m))

/* FIFO queue */
typedef struct ucc_queue_elem { struct ucc_queue_elem *next; } ucc_queue_elem_t;
typedef struct ucc_queue_head { ucc_queue_elem_t *head; ucc_queue_elem_t **ptail; unsigned long length; } ucc_queue_head_t;
static inline void ucc_queue_head_init(ucc_queue_head_t *q) { q->head = NULL; q->ptail = &q->head; q->length = 0; }
static inline int  ucc_queue_is_empty(const ucc_queue_head_t *q) { return q->head == NULL; }
static inline void ucc_queue_push(ucc_queue_head_t *q, ucc_queue_elem_t *e)
{ e->next = NULL; *q->ptail = e; q->ptail = &e->next; q->length++; }
static inline ucc_queue_elem_t *ucc_queue_pull(ucc_queue_head_t *q)
{
    ucc_queue_elem_t *e = q->head;
    if (!e) return NULL;
    q->head = e->next; if (!q->head) q->ptail = &q->head; q->length--; return e;
}
#endif
