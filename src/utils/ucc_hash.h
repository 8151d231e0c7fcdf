/* Small open-addressing hash map uint64 -> pointer (replaces khash use). */
#ifndef UCC_HASH_H_
#define UCC_HASH_H_
#include "ucc_compiler_def.h"
typedef struct ucc_hash_slot { uint64_t key; void *val; uint8_t state; /* 0 empty 1 used 2 tomb */ } ucc_hash_slot_t;
typedef struct ucc_hash { ucc_hash_slot_t *slots; uint32_t cap, n_used, n_tomb; } ucc_hash_t;
static inline uint64_t ucc_hash_mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33; return k; }
static inline void ucc_hash_init(ucc_hash_t *h) { h->slots = NULL; h->cap = h->n_used = h->n_tomb = 0; }
static inline void ucc_hash_destroy(ucc_hash_t *h) { free(h->slots); ucc_hash_init(h); }
static inline void *ucc_hash_get(const ucc_hash_t *h, uint64_t key)
{
    if (!h->cap) return NULL;
    for (uint32_t i = (uint32_t)ucc_hash_mix(key) & (h->cap - 1), n = 0; n < h->cap; n++, i = (i + 1) & (h->cap - 1)) {
        if (h->slots[i].state == 0) return NULL;
        if (h->slots[i].state == 1 && h->slots[i].key == key) return h->slots[i].val;
    }
    return NULL;
}
static inline int ucc_hash_put(ucc_hash_t *h, uint64_t key, void *val);
static inline int ucc_hash_grow_(ucc_hash_t *h)
{
    ucc_hash_t n; uint32_t cap = h->cap ? h->cap * 2 : 16;
    n.slots = (ucc_hash_slot_t *)calloc(cap, sizeof(ucc_hash_slot_t)); if (!n.slots) return -1;
    n.cap = cap; n.n_used = n.n_tomb = 0;
    for (uint32_t i = 0; i < h->cap; i++) if (h->slots[i].state == 1) ucc_hash_put(&n, h->slots[i].key, h->slots[i].val);
    free(h->slots); *h = n; return 0;
}
static inline int ucc_hash_put(ucc_hash_t *h, uint64_t key, void *val)
{
    int32_t tomb = -1;
    if ((h->n_used + h->n_tomb + 1) * 4 > h->cap * 3 && ucc_hash_grow_(h)) return -1;
    for (uint32_t i = (uint32_t)ucc_hash_mix(key) & (h->cap - 1);; i = (i + 1) & (h->cap - 1)) {
        if (h->slots[i].state == 1 && h->slots[i].key == key) { h->slots[i].val = val; return 0; }
        if (h->slots[i].state == 2 && tomb < 0) tomb = (int32_t)i;
        if (h->slots[i].state == 0) {
            if (tomb >= 0) { i = (uint32_t)tomb; h->n_tomb--; }
            h->slots[i].key = key; h->slots[i].val = val; h->slots[i].state = 1; h->n_used++; return 0;
        }
    }
}
static inline void *ucc_hash_del(ucc_hash_t *h, uint64_t key)
{
    if (!h->cap) return NULL;
    for (uint32_t i = (uint32_t)ucc_hash_mix(key) & (h->cap - 1), n = 0; n < h->cap; n++, i = (i + 1) & (h->cap - 1)) {
        if (h->slots[i].state == 0) return NULL;
        if (h->slots[i].state == 1 && h->slots[i].key == key) { h->slots[i].state = 2; h->n_used--; h->n_tomb++; return h->slots[i].val; }
    }
    return NULL;
}
#define ucc_hash_for_each(_h, _k, _v) \
    for (uint32_t _i = 0; _i < (_h)->cap; _i++) \
        if ((_h)->slots[_i].state == 1 && ((_k) = (_h)->slots[_i].key, (_v) = (_h)->slots[_i].val, 1))
#endif
