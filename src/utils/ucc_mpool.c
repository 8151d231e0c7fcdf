#include "ucc_mpool.h"
#include "ucc_log.h"
#include "ucc_math.h"

ucc_status_t ucc_mpool_init(ucc_mpool_t *mp, size_t priv_size, size_t elem_size, size_t align_offset, size_t alignment,
                            unsigned elems_per_chunk, unsigned max_elems, ucc_mpool_ops_t *ops,
                            ucc_thread_mode_t tm, const char *name)
{
    (void)align_offset;
    memset(mp, 0, sizeof(*mp));
    mp->elem_size = elem_size;
    mp->align     = alignment < sizeof(void *) ? sizeof(void *) : alignment;
    mp->stride    = ucc_align_up(elem_size + mp->align, mp->align); /* header occupies one alignment unit */
    mp->elems_per_chunk = elems_per_chunk ? elems_per_chunk : 8;
    mp->max_elems = max_elems;
    mp->name      = name;
    mp->priv_size = priv_size;
    mp->priv      = priv_size ? calloc(1, priv_size) : NULL;
    if (ops) mp->ops = *ops;
    ucc_tm_lock_init(&mp->lock, tm == UCC_THREAD_MULTIPLE);
    return UCC_OK;
}

static ucc_status_t mpool_grow(ucc_mpool_t *mp)
{
    unsigned n = mp->elems_per_chunk;
    size_t   size;
    void    *raw = NULL;
    ucc_mpool_chunk_t *ch;
    if (mp->max_elems != (unsigned)-1 && mp->max_elems && mp->n_elems + n > mp->max_elems) {
        if (mp->n_elems >= mp->max_elems) return UCC_ERR_NO_RESOURCE;
        n = mp->max_elems - mp->n_elems;
    }
    size = (size_t)n * mp->stride + mp->align;
    if (mp->ops.chunk_alloc) { if (mp->ops.chunk_alloc(mp, &size, &raw) != UCC_OK) return UCC_ERR_NO_MEMORY; }
    else if (posix_memalign(&raw, ucc_max(mp->align, (size_t)64), size)) return UCC_ERR_NO_MEMORY;
    ch = (ucc_mpool_chunk_t *)malloc(sizeof(*ch));
    ch->raw = raw; ch->n_elems = n; ch->next = mp->chunks; mp->chunks = ch;
    for (unsigned i = 0; i < n; i++) {
        /* layout of one slot: [pad .. header(8B)] [object aligned to mp->align] */
        char *obj = (char *)raw + (size_t)i * mp->stride + mp->align;
        ucc_mpool_elem_t *e = (ucc_mpool_elem_t *)(obj - sizeof(ucc_mpool_elem_t));
        if (mp->ops.obj_init) mp->ops.obj_init(mp, obj, raw);
        e->next = mp->freelist; mp->freelist = e;
    }
    mp->n_elems += n;
    return UCC_OK;
}

void *ucc_mpool_get(ucc_mpool_t *mp)
{
    ucc_mpool_elem_t *e;
    ucc_tm_lock(&mp->lock);
    if (ucc_unlikely(!mp->freelist) && mpool_grow(mp) != UCC_OK) { ucc_tm_unlock(&mp->lock); return NULL; }
    e = mp->freelist; mp->freelist = e->next; e->mp = mp; mp->n_in_use++;
    ucc_tm_unlock(&mp->lock);
    return (void *)(e + 1);
}

void ucc_mpool_put(void *obj)
{
    ucc_mpool_elem_t *e = (ucc_mpool_elem_t *)obj - 1;
    ucc_mpool_t *mp = e->mp;
    ucc_tm_lock(&mp->lock);
    e->next = mp->freelist; mp->freelist = e; mp->n_in_use--;
    ucc_tm_unlock(&mp->lock);
}

void ucc_mpool_cleanup(ucc_mpool_t *mp, int leak_check)
{
    ucc_mpool_chunk_t *ch, *next;
    if (leak_check && mp->n_in_use) ucc_warn("mpool %s: %u objects were not returned", mp->name ? mp->name : "?", mp->n_in_use);
    for (ch = mp->chunks; ch; ch = next) {
        next = ch->next;
        if (mp->ops.obj_cleanup)
            for (unsigned i = 0; i < ch->n_elems; i++) mp->ops.obj_cleanup(mp, (char *)ch->raw + (size_t)i * mp->stride + mp->align);
        if (mp->ops.chunk_release) mp->ops.chunk_release(mp, ch->raw); else free(ch->raw);
        free(ch);
    }
    free(mp->priv);
    memset(mp, 0, sizeof(*mp));
}
