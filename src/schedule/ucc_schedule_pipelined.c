#include "ucc_schedule_pipelined.h"
#include "utils/ucc_log.h"

static ucc_status_t launch_frag(ucc_schedule_pipelined_t *sp, int slot, int global)
{
    ucc_schedule_t *f = sp->frags[slot];
    ucc_status_t    st;
    if (sp->frag_setup) { st = sp->frag_setup(sp, f, global); if (st != UCC_OK) return st; }
    sp->slot_global[slot] = global;
    f->n_completed_tasks  = 0;
    f->super.status = UCC_INPROGRESS; f->super.super.status = UCC_INPROGRESS;
    f->super.start_time = sp->super.super.start_time; f->super.timeout = sp->super.super.timeout;
    for (unsigned j = 0; j < f->n_tasks; j++) {
        ucc_coll_task_t *t = f->tasks[j];
        int xdep = (global > 0 && sp->order != UCC_PIPELINE_PARALLEL && sp->n_frags > 1);
        sp->fired[slot][j] = 0;
        t->super.status = UCC_OPERATION_INITIALIZED; t->status = UCC_OPERATION_INITIALIZED;
        t->n_deps = t->n_deps_base + (xdep ? 1 : 0);
        t->n_deps_satisfied = 0;
        t->generation++;
        if (xdep) {
            /* predecessor = task j of fragment global-1.  Fragments are launched in order but may finish out of
             * order, so look the predecessor up: not resident any more == completed, resident == check its fire mark */
            int pslot = -1;
            for (int s = 0; s < sp->n_frags; s++) if (s != slot && sp->slot_global[s] == global - 1) pslot = s;
            if (pslot < 0 || j >= sp->frags[pslot]->n_tasks || sp->fired[pslot][j]) t->n_deps_satisfied = 1;
        } else if (global > 0 && sp->order != UCC_PIPELINE_PARALLEL && sp->n_frags == 1) {
            /* single slot: strictly sequential by construction */
        }
        t->timeout = f->super.timeout; t->start_time = f->super.start_time;
    }
    if (f->n_tasks == 0) { f->super.status = UCC_OK; return ucc_task_complete(&f->super) >= 0 ? UCC_OK : f->super.status; }
    return ucc_event_manager_notify(&f->super, UCC_EVENT_SCHEDULE_STARTED);
}

/* Fragment g always runs in slot g % n_frags and fragments are launched in order: the slot decides the message tags of the
 * fragment's tasks (TL tasks take theirs at init), so every rank must use the same slot for the same fragment - also a rank
 * whose fragments complete inline (a bcast root) or out of order.  A slot that finished early therefore waits for its turn. */
static ucc_status_t pump(ucc_schedule_pipelined_t *sp, int *done)
{
    ucc_status_t st = UCC_OK;
    *done = 0;
    if (sp->done_ptr) return UCC_OK; /* a fragment completed inside its own launch: the frame that is launching goes on (no recursion) */
    sp->done_ptr = done;
    while (st == UCC_OK && !*done && sp->n_frags_started < sp->n_frags_total) {
        int g = sp->n_frags_started, s = g % sp->n_frags;
        if (!sp->slot_idle[s]) break;
        sp->slot_idle[s] = 0; sp->n_frags_started = g + 1;
        st = launch_frag(sp, s, g);
    }
    sp->done_ptr = NULL;
    return st;
}
/* the last fragment completed: the schedule is completed by the outermost frame, after it released the lock (the user may finalize the
 * request as soon as its status is final) */
static void complete_schedule(ucc_schedule_pipelined_t *sp) { sp->super.super.status = UCC_OK; ucc_task_complete(&sp->super.super); }

static void find_task(ucc_schedule_pipelined_t *sp, ucc_coll_task_t *t, int *slot, int *j)
{
    for (int s = 0; s < sp->n_frags; s++)
        for (unsigned k = 0; k < sp->frags[s]->n_tasks; k++)
            if (sp->frags[s]->tasks[k] == t) { *slot = s; *j = (int)k; return; }
    *slot = -1; *j = -1;
}

/* record that (slot ps, task pj) emitted its ordering event and wake the matching task of the
 * next fragment if that fragment is armed and waiting for exactly this predecessor */
static ucc_status_t fire(ucc_schedule_pipelined_t *sp, int ps, int pj)
{
    int ss = -1;
    if (sp->fired[ps][pj]) return UCC_OK;
    sp->fired[ps][pj] = 1;
    for (int s = 0; s < sp->n_frags; s++) if (s != ps && sp->slot_global[s] == sp->slot_global[ps] + 1) ss = s;
    if (ss >= 0 && (unsigned)pj < sp->frags[ss]->n_tasks) {
        ucc_coll_task_t *succ = sp->frags[ss]->tasks[pj];
        if (succ->super.status == UCC_OPERATION_INITIALIZED) return ucc_dependency_handler(sp->frags[ps]->tasks[pj], succ);
    }
    return UCC_OK;
}

static ucc_status_t cross_frag_handler(ucc_coll_task_t *parent, ucc_coll_task_t *task)
{
    ucc_schedule_pipelined_t *sp = ucc_derived_of(parent->schedule->super.schedule, ucc_schedule_pipelined_t);
    int ps, pj;
    ucc_status_t st = UCC_OK;
    (void)task;
    ucc_recursive_spin_lock(&sp->lock);
    find_task(sp, parent, &ps, &pj);
    /* an event of an incarnation that has been re-armed since is stale (everything it had to deliver was delivered by
     * frag_completed_handler): the current incarnation has emitted COMPLETED only if it is OK, TASK_STARTED only if posted */
    if (ps >= 0 && sp->slot_global[ps] >= 0 && sp->slot_global[ps] < sp->n_frags_total &&
        (sp->order == UCC_PIPELINE_SEQUENTIAL ? parent->super.status == UCC_OK : parent->super.status != UCC_OPERATION_INITIALIZED))
        st = fire(sp, ps, pj);
    ucc_recursive_spin_unlock(&sp->lock);
    return st;
}

static ucc_status_t frag_completed_handler(ucc_coll_task_t *parent, ucc_coll_task_t *task)
{
    ucc_schedule_pipelined_t *sp   = ucc_derived_of(task, ucc_schedule_pipelined_t);
    ucc_schedule_t           *frag = ucc_derived_of(parent, ucc_schedule_t);
    ucc_status_t              st   = UCC_OK;
    int                       slot = -1, done = 0;
    ucc_recursive_spin_lock(&sp->lock);
    for (int s = 0; s < sp->n_frags; s++) if (sp->frags[s] == frag) slot = s;
    sp->n_frags_completed++;
    if (sp->n_frags_completed == sp->n_frags_total) {
        int nested = sp->done_ptr != NULL;
        for (int s = 0; s < sp->n_frags; s++) sp->slot_global[s] = -1;
        if (nested) *sp->done_ptr = 1;
        ucc_recursive_spin_unlock(&sp->lock);
        if (!nested) complete_schedule(sp);
        return UCC_OK;
    }
    /* deliver any ordering event of the finished incarnation that was not delivered yet
     * (e.g. TASK_STARTED of a task that completed inside its own post) */
    if (sp->order != UCC_PIPELINE_PARALLEL && sp->n_frags > 1)
        for (unsigned j = 0; j < frag->n_tasks && st == UCC_OK; j++) st = fire(sp, slot, (int)j);
    if (st == UCC_OK) {
        for (unsigned j = 0; j < frag->n_tasks; j++) sp->fired[slot][j] = 1;
        sp->slot_global[slot] = sp->n_frags_total + slot; /* idle: not a predecessor of anything */
        sp->slot_idle[slot] = 1;
        st = pump(sp, &done);
    }
    ucc_recursive_spin_unlock(&sp->lock);
    if (done) complete_schedule(sp);
    return st;
}

static ucc_status_t frag_error_handler(ucc_coll_task_t *parent, ucc_coll_task_t *task)
{
    if (task->super.status >= 0) { task->status = parent->status; ucc_task_complete(task); }
    return UCC_OK;
}

ucc_status_t ucc_schedule_pipelined_post(ucc_coll_task_t *task)
{
    ucc_schedule_pipelined_t *sp = ucc_derived_of(task, ucc_schedule_pipelined_t);
    ucc_status_t st = UCC_OK;
    int done = 0;
    sp->super.super.status = UCC_INPROGRESS; sp->super.super.super.status = UCC_INPROGRESS;
    sp->n_frags_completed = 0; sp->n_frags_started = 0;
    ucc_recursive_spin_lock(&sp->lock);
    for (int s = 0; s < sp->n_frags; s++) { sp->slot_global[s] = -1; sp->slot_idle[s] = 1; memset(sp->fired[s], 0, sizeof(sp->fired[s])); }
    if (sp->n_frags_total == 0) {
        ucc_recursive_spin_unlock(&sp->lock);
        sp->super.super.status = UCC_OK; ucc_task_complete(&sp->super.super); return UCC_OK;
    }
    st = pump(sp, &done);
    ucc_recursive_spin_unlock(&sp->lock);
    if (done) complete_schedule(sp);
    return st;
}

ucc_status_t ucc_schedule_pipelined_finalize(ucc_coll_task_t *task)
{
    ucc_schedule_pipelined_t *sp = ucc_derived_of(task, ucc_schedule_pipelined_t);
    ucc_status_t st_all = UCC_OK;
    for (int s = 0; s < sp->n_frags; s++) {
        ucc_status_t st = sp->frags[s]->super.finalize(&sp->frags[s]->super);
        if (st != UCC_OK) st_all = st;
    }
    ucc_coll_task_destruct(task);
    return st_all;
}

ucc_status_t ucc_schedule_pipelined_init(ucc_base_coll_args_t *coll_args, ucc_base_team_t *team,
                                         ucc_schedule_frag_init_fn_t frag_init, ucc_schedule_frag_setup_fn_t frag_setup,
                                         int n_frags, int n_frags_total, ucc_pipeline_order_t order,
                                         ucc_schedule_pipelined_t *sp)
{
    ucc_status_t st;
    if (n_frags < 1 || n_frags > UCC_SCHEDULE_PIPELINED_MAX_FRAGS) return UCC_ERR_INVALID_PARAM;
    UCC_CHECK_RET(ucc_schedule_init(&sp->super, coll_args, team));
    sp->super.super.flags |= UCC_COLL_TASK_FLAG_IS_PIPELINED_SCHEDULE;
    sp->super.super.post = ucc_schedule_pipelined_post;
    sp->super.super.finalize = ucc_schedule_pipelined_finalize;
    sp->n_frags = n_frags; sp->n_frags_total = n_frags_total; sp->order = order; sp->frag_setup = frag_setup;
    sp->n_frags_started = sp->n_frags_completed = 0; sp->done_ptr = NULL;
    ucc_recursive_spinlock_init(&sp->lock);
    for (int s = 0; s < n_frags; s++) {
        st = frag_init(coll_args, sp, team, &sp->frags[s]);
        if (st != UCC_OK) {
            for (int k = 0; k < s; k++) sp->frags[k]->super.finalize(&sp->frags[k]->super);
            return st;
        }
        sp->frags[s]->super.schedule = &sp->super;
        sp->frags[s]->super.flags |= UCC_COLL_TASK_FLAG_INTERNAL;
        if (sp->frags[s]->super.flags & UCC_COLL_TASK_FLAG_EXECUTOR) sp->super.super.flags |= UCC_COLL_TASK_FLAG_EXECUTOR;
        for (unsigned j = 0; j < sp->frags[s]->n_tasks; j++) sp->frags[s]->tasks[j]->n_deps_base = sp->frags[s]->tasks[j]->n_deps;
        UCC_CHECK_RET(ucc_event_manager_subscribe(&sp->frags[s]->super, UCC_EVENT_COMPLETED_SCHEDULE, &sp->super.super, frag_completed_handler));
        UCC_CHECK_RET(ucc_event_manager_subscribe(&sp->frags[s]->super, UCC_EVENT_ERROR, &sp->super.super, frag_error_handler));
    }
    if (order != UCC_PIPELINE_PARALLEL && n_frags > 1) {
        ucc_event_t ev = order == UCC_PIPELINE_ORDERED ? UCC_EVENT_TASK_STARTED : UCC_EVENT_COMPLETED;
        for (int s = 0; s < n_frags; s++) {
            ucc_schedule_t *a = sp->frags[s], *b = sp->frags[(s + 1) % n_frags];
            for (unsigned j = 0; j < a->n_tasks && j < b->n_tasks; j++)
                UCC_CHECK_RET(ucc_event_manager_subscribe(a->tasks[j], ev, b->tasks[j], cross_frag_handler));
        }
    }
    return UCC_OK;
}
