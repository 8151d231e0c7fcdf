#include "ucc_progress_queue.h"
#include "utils/ucc_atomic.h"
#include "utils/ucc_log.h"

/* ---------------- single threaded ---------------- */
typedef struct ucc_pq_st { ucc_progress_queue_t super; ucc_list_link_t list; } ucc_pq_st_t;

static void st_enqueue(ucc_progress_queue_t *pq, ucc_coll_task_t *task) { ucc_list_add_tail(&((ucc_pq_st_t *)pq)->list, &task->list_elem); }
static void st_dequeue(ucc_progress_queue_t *pq, ucc_coll_task_t **task)
{
    ucc_pq_st_t *q = (ucc_pq_st_t *)pq;
    *task = ucc_list_is_empty(&q->list) ? NULL : ucc_list_extract_head(&q->list, ucc_coll_task_t, list_elem);
}
static int st_progress(ucc_progress_queue_t *pq)
{
    ucc_pq_st_t *q = (ucc_pq_st_t *)pq;
    ucc_coll_task_t *task, *tmp;
    int n = 0;
    ucc_list_for_each_safe(task, tmp, &q->list, list_elem) {
        if (ucc_pq_progress_one(task)) {
            /* completion may enqueue new tasks at the tail; they are visited later, `tmp` stays valid */
            ucc_list_del(&task->list_elem);
            n++;
            if (ucc_task_complete(task) < 0) return (int)task->status;
        }
    }
    return n;
}
static int  st_is_empty(ucc_progress_queue_t *pq) { return ucc_list_is_empty(&((ucc_pq_st_t *)pq)->list); }
static void st_finalize(ucc_progress_queue_t *pq) { free(pq); }

ucc_status_t ucc_pq_st_init(ucc_progress_queue_t **pq)
{
    ucc_pq_st_t *q = (ucc_pq_st_t *)calloc(1, sizeof(*q));
    if (!q) return UCC_ERR_NO_MEMORY;
    ucc_list_head_init(&q->list);
    q->super.enqueue = st_enqueue; q->super.dequeue = st_dequeue; q->super.progress = st_progress;
    q->super.is_empty = st_is_empty; q->super.finalize = st_finalize;
    *pq = &q->super;
    return UCC_OK;
}

/* ---------------- multi threaded ---------------- */
typedef struct ucc_pq_mt {
    ucc_progress_queue_t super;
    int                  lock_free;
    ucc_spinlock_t       lock;
    ucc_list_link_t      list;
    ucc_lf_queue_t      *lfq;
} ucc_pq_mt_t;

static void mt_enqueue_locked(ucc_progress_queue_t *pq, ucc_coll_task_t *task)
{ ucc_pq_mt_t *q = (ucc_pq_mt_t *)pq; ucc_spin_lock(&q->lock); ucc_list_add_tail(&q->list, &task->list_elem); ucc_spin_unlock(&q->lock); }
static void mt_dequeue_locked(ucc_progress_queue_t *pq, ucc_coll_task_t **task)
{
    ucc_pq_mt_t *q = (ucc_pq_mt_t *)pq;
    ucc_spin_lock(&q->lock);
    *task = ucc_list_is_empty(&q->list) ? NULL : ucc_list_extract_head(&q->list, ucc_coll_task_t, list_elem);
    ucc_spin_unlock(&q->lock);
}
static void mt_enqueue_lf(ucc_progress_queue_t *pq, ucc_coll_task_t *task) { ucc_lf_queue_enqueue(((ucc_pq_mt_t *)pq)->lfq, &task->lf_elem); }
static void mt_dequeue_lf(ucc_progress_queue_t *pq, ucc_coll_task_t **task)
{
    ucc_lf_queue_elem_t *e = ucc_lf_queue_dequeue(((ucc_pq_mt_t *)pq)->lfq);
    *task = e ? ucc_container_of(e, ucc_coll_task_t, lf_elem) : NULL;
}
/* take one task, progress it, put it back if still running: threads never touch the same task */
static int mt_progress(ucc_progress_queue_t *pq)
{
    ucc_coll_task_t *task;
    pq->dequeue(pq, &task);
    if (!task) return 0;
    if (ucc_pq_progress_one(task)) { ucc_status_t st = ucc_task_complete(task); return st < 0 ? (int)st : 1; } /* the task may already be finalized by its owner */
    pq->enqueue(pq, task);
    return 0;
}
static int mt_is_empty(ucc_progress_queue_t *pq)
{ ucc_pq_mt_t *q = (ucc_pq_mt_t *)pq; int e; if (q->lock_free) return 0; /* cannot tell cheaply */ ucc_spin_lock(&q->lock); e = ucc_list_is_empty(&q->list); ucc_spin_unlock(&q->lock); return e; }
static void mt_finalize(ucc_progress_queue_t *pq)
{ ucc_pq_mt_t *q = (ucc_pq_mt_t *)pq; if (q->lfq) { ucc_lf_queue_destroy(q->lfq); free(q->lfq); } free(q); }

ucc_status_t ucc_pq_mt_init(ucc_progress_queue_t **pq, uint32_t lock_free)
{
    ucc_pq_mt_t *q = (ucc_pq_mt_t *)calloc(1, sizeof(*q));
    if (!q) return UCC_ERR_NO_MEMORY;
    q->lock_free = (int)lock_free;
    ucc_spinlock_init(&q->lock); ucc_list_head_init(&q->list);
    if (lock_free) {
        if (posix_memalign((void **)&q->lfq, 64, sizeof(ucc_lf_queue_t))) { free(q); return UCC_ERR_NO_MEMORY; }
        ucc_lf_queue_init(q->lfq);
        q->super.enqueue = mt_enqueue_lf; q->super.dequeue = mt_dequeue_lf;
    } else { q->super.enqueue = mt_enqueue_locked; q->super.dequeue = mt_dequeue_locked; }
    q->super.progress = mt_progress; q->super.is_empty = mt_is_empty; q->super.finalize = mt_finalize;
    *pq = &q->super;
    return UCC_OK;
}

ucc_status_t ucc_progress_queue_init(ucc_progress_queue_t **pq, ucc_thread_mode_t tm, uint32_t lock_free_progress_q)
{ return tm == UCC_THREAD_MULTIPLE ? ucc_pq_mt_init(pq, lock_free_progress_q) : ucc_pq_st_init(pq); }
void ucc_progress_queue_finalize(ucc_progress_queue_t *pq) { if (pq) pq->finalize(pq); }
