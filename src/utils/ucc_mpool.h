/* Fixed-size object pool with chunked growth and optional custom chunk
 * allocator (role of reference utils/ucc_mpool.c, which wraps UCS mpool). */
#ifndef UCC_MPOOL_H_
#define UCC_MPOOL_H_
#include "ucc_compiler_def.h"
#include "ucc_atomic.h"
#include <ucc/api/ucc.h>

typedef struct ucc_mpool ucc_mpool_t;
typedef struct ucc_mpool_ops {
    ucc_status_t (*chunk_alloc)(ucc_mpool_t *mp, size_t *size_p, void **chunk_p);
    void         (*chunk_release)(ucc_mpool_t *mp, void *chunk);
    void         (*obj_init)(ucc_mpool_t *mp, void *obj, void *chunk);
    void         (*obj_cleanup)(ucc_mpool_t *mp, void *obj);
} ucc_mpool_ops_t;

typedef struct ucc_mpool_elem { union { struct ucc_mpool_elem *next; ucc_mpool_t *mp; }; } ucc_mpool_elem_t;
typedef struct ucc_mpool_chunk { struct ucc_mpool_chunk *next; void *raw; unsigned n_elems; } ucc_mpool_chunk_t;

struct ucc_mpool {
    ucc_mpool_elem_t  *freelist;
    ucc_mpool_chunk_t *chunks;
    size_t             elem_size, align, stride, priv_size;
    unsigned           elems_per_chunk, max_elems, n_elems, n_in_use;
    ucc_mpool_ops_t    ops;
    ucc_tm_lock_t      lock;
    const char        *name;
    void              *priv;
};

ucc_status_t ucc_mpool_init(ucc_mpool_t *mp, size_t priv_size, size_t elem_size, size_t align_offset, size_t alignment,
                            unsigned elems_per_chunk, unsigned max_elems, ucc_mpool_ops_t *ops,
                            ucc_thread_mode_t tm, const char *name);
void  ucc_mpool_cleanup(ucc_mpool_t *mp, int leak_check);
void *ucc_mpool_get(ucc_mpool_t *mp);
void  ucc_mpool_put(void *obj);
static inline void *ucc_mpool_priv(ucc_mpool_t *mp) { return mp->priv; }
#endif
