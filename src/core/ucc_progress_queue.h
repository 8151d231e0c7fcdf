/* Progress queues: single-threaded list, spin-locked list, lock-free ring
 * (reference core/ucc_progress_queue{.h,_st.c,_mt.c}). */
#ifndef UCC_PROGRESS_QUEUE_H_
#define UCC_PROGRESS_QUEUE_H_
#include "schedule/ucc_schedule.h"

typedef struct ucc_progress_queue ucc_progress_queue_t;
struct ucc_progress_queue {
    void (*enqueue)(ucc_progress_queue_t *pq, ucc_coll_task_t *task);
    void (*dequeue)(ucc_progress_queue_t *pq, ucc_coll_task_t **task);
    int  (*progress)(ucc_progress_queue_t *pq);
    int  (*is_empty)(ucc_progress_queue_t *pq);
    void (*finalize)(ucc_progress_queue_t *pq);
};

ucc_status_t ucc_progress_queue_init(ucc_progress_queue_t **pq, ucc_thread_mode_t tm, uint32_t lock_free_progress_q);
ucc_status_t ucc_pq_st_init(ucc_progress_queue_t **pq);
ucc_status_t ucc_pq_mt_init(ucc_progress_queue_t **pq, uint32_t lock_free);
void         ucc_progress_queue_finalize(ucc_progress_queue_t *pq);

/* Enqueue = run one progress step inline; only tasks that are still in
 * flight are actually queued. */
static inline ucc_status_t ucc_progress_queue_enqueue(ucc_progress_queue_t *pq, ucc_coll_task_t *task)
{
    task->progress_queue = pq;
    task->status       = UCC_INPROGRESS;
    task->super.status = UCC_INPROGRESS;
    if (task->progress) {
        task->progress(task);
        if (task->status != UCC_INPROGRESS) return ucc_task_complete(task) < 0 ? task->status : UCC_OK;
    }
    pq->enqueue(pq, task);
    return UCC_OK;
}
static inline int ucc_progress_queue(ucc_progress_queue_t *pq) { return pq->progress(pq); }
static inline int ucc_progress_queue_is_empty(ucc_progress_queue_t *pq) { return pq->is_empty(pq); }

/* shared by both flavours: returns 1 when the task left the in-progress state */
static inline int ucc_pq_progress_one(ucc_coll_task_t *task)
{
    if (task->progress) task->progress(task);
    if (task->status == UCC_INPROGRESS && task->timeout > 0 && ucc_unlikely(ucc_get_time() - task->start_time > task->timeout))
        task->status = UCC_ERR_TIMED_OUT;
    return task->status != UCC_INPROGRESS;
}
#endif
