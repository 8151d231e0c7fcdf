#include "ucc_schedule.h"
#include "core/ucc_progress_queue.h"
#include "core/ucc_team.h"
#include "core/ucc_context.h"
#include "utils/ucc_log.h"
#include "utils/ucc_atomic.h"
#include "utils/profile/ucc_profile.h"

/* ---------------- events ---------------- */
ucc_status_t ucc_event_manager_subscribe(ucc_coll_task_t *parent, ucc_event_t event, ucc_coll_task_t *task,
                                         ucc_task_event_handler_p handler)
{
    ucc_event_manager_t *em = &parent->em;
    if (em->n == em->cap) {
        unsigned ncap = em->cap * 2;
        ucc_em_listener_t *nl = (ucc_em_listener_t *)malloc(ncap * sizeof(*nl));
        if (!nl) return UCC_ERR_NO_MEMORY;
        memcpy(nl, em->listeners, em->n * sizeof(*nl));
        if (em->listeners != em->inl) free(em->listeners);
        em->listeners = nl; em->cap = ncap;
    }
    em->listeners[em->n].task = task; em->listeners[em->n].event = event; em->listeners[em->n].handler = handler;
    em->n++;
    return UCC_OK;
}

ucc_status_t ucc_event_manager_notify(ucc_coll_task_t *parent, ucc_event_t event)
{
    ucc_event_manager_t *em = &parent->em;
    /* a COMPLETED_SCHEDULE listener may release `parent`: copy what we need first */
    unsigned n = em->n;
    for (unsigned i = 0; i < n; i++) {
        ucc_em_listener_t l = em->listeners[i];
        if (l.event != event) continue;
        ucc_status_t st = l.handler(parent, l.task);
        if (st != UCC_OK) return st;
        if (event == UCC_EVENT_COMPLETED_SCHEDULE) break; /* at most one owner */
    }
    return UCC_OK;
}

/* ---------------- task ---------------- */
ucc_status_t ucc_coll_finalize_default_(ucc_coll_task_t *task) { (void)task; return UCC_OK; }
static ucc_status_t post_default(ucc_coll_task_t *task) { (void)task; return UCC_OK; }
static ucc_status_t tp_setup_default(ucc_coll_task_t *task) { (void)task; return UCC_OK; }

ucc_status_t ucc_coll_task_init(ucc_coll_task_t *task, ucc_base_coll_args_t *bargs, ucc_base_team_t *team)
{
    task->super.status = UCC_OPERATION_INITIALIZED;
    task->status       = UCC_OPERATION_INITIALIZED;
    task->flags        = 0;
    task->team         = team;
    task->schedule     = NULL;
    task->post = post_default; task->progress = NULL; task->finalize = ucc_coll_finalize_default_;
    task->triggered_post = ucc_triggered_post; task->triggered_post_setup = tp_setup_default;
    task->cb.cb = NULL; task->cb.data = NULL;
    task->ee = NULL; task->ev = NULL; task->ee_task = NULL; task->triggered_task = NULL;
    task->em.listeners = task->em.inl; task->em.n = 0; task->em.cap = 4;
    task->executor = NULL;
    task->n_deps = task->n_deps_satisfied = task->n_deps_base = 0;
    task->start_time = 0; task->timeout = 0; task->seq_num = 0; task->progress_queue = NULL; task->generation = 0; task->init_fn = NULL;
    if (bargs) memcpy(&task->bargs, bargs, sizeof(*bargs)); else memset(&task->bargs, 0, sizeof(task->bargs));
    return UCC_OK;
}

void ucc_coll_task_destruct(ucc_coll_task_t *task)
{
    if (task->em.listeners != task->em.inl) { free(task->em.listeners); task->em.listeners = task->em.inl; }
    task->em.n = 0; task->em.cap = 4;
}

ucc_status_t ucc_coll_task_get_executor(ucc_coll_task_t *task, ucc_ee_executor_t **exec)
{
    ucc_coll_task_t *t = task;
    while (t) {
        if (t->executor) { *exec = t->executor; return UCC_OK; }
        t = t->schedule ? &t->schedule->super : NULL;
    }
    ucc_error("task %p: executor requested but none was created up the schedule chain", (void *)task);
    return UCC_ERR_NOT_FOUND;
}

/* error status spreads to every listener of the ERROR event (schedules fail as a whole) */
static ucc_status_t error_handler(ucc_coll_task_t *parent, ucc_coll_task_t *task)
{
    if (task->super.status != parent->status && task->super.status >= 0) {
        task->status = parent->status;
        ucc_task_complete(task);
    }
    return UCC_OK;
}

ucc_status_t ucc_task_complete(ucc_coll_task_t *task)
{
    ucc_status_t        status = task->status;
    ucc_coll_callback_t cb     = task->cb;
    int                 has_cb = task->flags & UCC_COLL_TASK_FLAG_CB;
    int                 in_sched = task->schedule != NULL;
    int                 auto_fin = task->flags & UCC_COLL_TASK_FLAG_AUTO_FINALIZE;

    ucc_assert(status != UCC_INPROGRESS);
    if (ucc_likely(status == UCC_OK)) {
        if (ucc_unlikely(task->bargs.asymm.scratch != NULL) && (task->flags & UCC_COLL_TASK_FLAG_TOP_LEVEL)) {
            extern ucc_status_t ucc_coll_asymm_copy_out(ucc_coll_task_t *task);
            status = ucc_coll_asymm_copy_out(task);
        }
        if (status == UCC_OK) status = ucc_event_manager_notify(task, UCC_EVENT_COMPLETED);
    }
    if (ucc_unlikely(status < 0)) {
        if (status == UCC_ERR_TIMED_OUT) {
            char buf[256];
            ucc_coll_args_str(&task->bargs.args, task->bargs.team ? ucc_team_rank_(task->bargs.team) : 0,
                              task->bargs.team ? ucc_team_size_(task->bargs.team) : 1, buf, sizeof(buf));
            if (task->timeout > 0) ucc_warn("timeout %g sec has expired on %s seq_num %u", task->timeout, buf, task->seq_num);
            else ucc_warn("%s seq_num %u timed out inside the transport (a team member did not arrive)", buf, task->seq_num);
        } else {
            ucc_error("failure in task %p, %s", (void *)task, ucc_status_string(status));
        }
        task->status = status;
        ucc_event_manager_notify(task, UCC_EVENT_ERROR);
    }
    if (task->executor) {
        if (task->flags & UCC_COLL_TASK_FLAG_EXECUTOR_STOP) {
            ucc_status_t st = ucc_ee_executor_stop(task->executor);
            if (st != UCC_OK && status >= 0) status = st;
        }
        if (task->flags & UCC_COLL_TASK_FLAG_EXECUTOR_DESTROY) {
            ucc_status_t st = ucc_ee_executor_finalize(task->executor);
            if (st != UCC_OK && status >= 0) status = st;
            task->executor = NULL;
        }
    }
    if (task->flags & UCC_COLL_TASK_FLAG_TOP_LEVEL) UCC_PROFILE_REQUEST_EVENT(task, "ucc_coll_complete", 0);
    ucc_store_release((volatile int *)&task->super.status, (int)status);
    if (has_cb) cb.cb(cb.data, status);
    if (in_sched && status == UCC_OK) ucc_event_manager_notify(task, UCC_EVENT_COMPLETED_SCHEDULE);
    if (auto_fin) task->finalize(task);
    return status;
}

/* ---------------- schedule ---------------- */
static ucc_status_t schedule_completed_handler(ucc_coll_task_t *parent, ucc_coll_task_t *task)
{
    ucc_schedule_t *s = ucc_derived_of(task, ucc_schedule_t);
    (void)parent;
    if (ucc_atomic_add32(&s->n_completed_tasks, 1) == s->n_tasks) {
        s->super.status = UCC_OK;
        ucc_task_complete(&s->super);
    }
    return UCC_OK;
}

ucc_status_t ucc_schedule_init(ucc_schedule_t *s, ucc_base_coll_args_t *bargs, ucc_base_team_t *team)
{
    ucc_status_t st = ucc_coll_task_init(&s->super, bargs, team);
    s->super.flags |= UCC_COLL_TASK_FLAG_IS_SCHEDULE;
    s->super.post = ucc_schedule_start; s->super.finalize = ucc_schedule_finalize;
    s->n_completed_tasks = 0; s->n_tasks = 0;
    s->ctx = bargs && bargs->team ? bargs->team->contexts[0] : NULL;
    return st;
}

ucc_status_t ucc_schedule_add_task(ucc_schedule_t *s, ucc_coll_task_t *task)
{
    if (s->n_tasks >= UCC_SCHEDULE_MAX_TASKS) { ucc_error("schedule %p: too many tasks", (void *)s); return UCC_ERR_NO_RESOURCE; }
    UCC_CHECK_RET(ucc_event_manager_subscribe(task, UCC_EVENT_COMPLETED_SCHEDULE, &s->super, schedule_completed_handler));
    UCC_CHECK_RET(ucc_event_manager_subscribe(task, UCC_EVENT_ERROR, &s->super, error_handler));
    task->schedule = s;
    s->tasks[s->n_tasks++] = task;
    if (task->flags & UCC_COLL_TASK_FLAG_EXECUTOR) s->super.flags |= UCC_COLL_TASK_FLAG_EXECUTOR;
    return UCC_OK;
}

ucc_status_t ucc_schedule_start(ucc_coll_task_t *task)
{
    ucc_schedule_t *s = ucc_derived_of(task, ucc_schedule_t);
    s->n_completed_tasks = 0;
    s->super.status = UCC_INPROGRESS; s->super.super.status = UCC_INPROGRESS;
    for (unsigned i = 0; i < s->n_tasks; i++) {
        s->tasks[i]->n_deps_satisfied = 0;
        s->tasks[i]->super.status = UCC_OPERATION_INITIALIZED; s->tasks[i]->status = UCC_OPERATION_INITIALIZED;
        if (task->timeout > 0) { s->tasks[i]->timeout = task->timeout; s->tasks[i]->start_time = task->start_time; }
    }
    if (s->n_tasks == 0) { s->super.status = UCC_OK; return ucc_task_complete(&s->super) >= 0 ? UCC_OK : s->super.status; }
    return ucc_event_manager_notify(&s->super, UCC_EVENT_SCHEDULE_STARTED);
}

ucc_status_t ucc_schedule_finalize(ucc_coll_task_t *task)
{
    ucc_schedule_t *s = ucc_derived_of(task, ucc_schedule_t);
    ucc_status_t st_all = UCC_OK;
    for (unsigned i = 0; i < s->n_tasks; i++) {
        ucc_status_t st = s->tasks[i]->finalize(s->tasks[i]);
        if (st != UCC_OK) st_all = st;
    }
    ucc_coll_task_destruct(task);
    return st_all;
}

ucc_status_t ucc_dependency_handler(ucc_coll_task_t *parent, ucc_coll_task_t *task)
{
    (void)parent;
    if (ucc_atomic_add32(&task->n_deps_satisfied, 1) == task->n_deps) {
        ucc_status_t st;
        task->start_time = parent->start_time;
        if (parent->timeout > 0 && task->timeout == 0) task->timeout = parent->timeout;
        uint32_t gen = task->generation;
        st = task->post(task);
        if (st < 0) { task->status = st; ucc_task_complete(task); return st; }
        /* the task may have completed inline and already been re-armed for another fragment */
        if (task->generation == gen) return ucc_event_manager_notify(task, UCC_EVENT_TASK_STARTED);
    }
    return UCC_OK;
}
