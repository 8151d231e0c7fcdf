/* Task / schedule engine.
 *
 * A `ucc_coll_task_t` is what a user request points to.  Tasks expose
 * post/progress/finalize entry points; an event manager lets tasks listen
 * to each other's lifecycle events, which is how schedules (DAGs of tasks)
 * and pipelined schedules are built.  Concepts follow reference
 * schedule/ucc_schedule.h:22-279; the implementation is independent
 * (growable listener arrays, no 4-entry blocks).
 */
#ifndef UCC_SCHEDULE_H_
#define UCC_SCHEDULE_H_
#include <ucc/api/ucc.h>
#include "utils/ucc_compiler_def.h"
#include "utils/ucc_list.h"
#include "utils/ucc_lock_free_queue.h"
#include "utils/ucc_coll_utils.h"
#include "utils/ucc_time.h"
#include "components/ec/ucc_ec.h"
#include "components/mc/ucc_mc.h"

#define UCC_SCHEDULE_MAX_TASKS 8

typedef struct ucc_coll_task     ucc_coll_task_t;
typedef struct ucc_schedule      ucc_schedule_t;
typedef struct ucc_base_team     ucc_base_team_t;
typedef struct ucc_team          ucc_team_t;

typedef enum {
    UCC_EVENT_COMPLETED = 0,
    UCC_EVENT_SCHEDULE_STARTED,
    UCC_EVENT_TASK_STARTED,
    UCC_EVENT_COMPLETED_SCHEDULE, /* fired last; the task may be released by the listener */
    UCC_EVENT_ERROR,
    UCC_EVENT_LAST
} ucc_event_t;

typedef ucc_status_t (*ucc_task_event_handler_p)(ucc_coll_task_t *parent, ucc_coll_task_t *task);
typedef ucc_status_t (*ucc_coll_post_fn_t)(ucc_coll_task_t *task);
typedef void         (*ucc_coll_progress_fn_t)(ucc_coll_task_t *task);
typedef ucc_status_t (*ucc_coll_finalize_fn_t)(ucc_coll_task_t *task);
typedef ucc_status_t (*ucc_coll_triggered_post_fn_t)(ucc_ee_h ee, ucc_ev_t *ev, ucc_coll_task_t *task);
typedef ucc_status_t (*ucc_coll_triggered_post_setup_fn_t)(ucc_coll_task_t *task);

typedef struct ucc_em_listener { ucc_coll_task_t *task; ucc_event_t event; ucc_task_event_handler_p handler; } ucc_em_listener_t;
typedef struct ucc_event_manager { ucc_em_listener_t *listeners; unsigned n, cap; ucc_em_listener_t inl[4]; } ucc_event_manager_t;

enum {
    UCC_BASE_CARGS_MAX_FRAG_COUNT = UCC_BIT(0)
};

/* arguments as seen by CL/TL init functions */
typedef struct ucc_base_coll_args {
    uint64_t        mask;
    ucc_coll_args_t args;
    ucc_team_t     *team;
    size_t          max_frag_count;
    /* rooted collectives whose root has different src/dst memory types are
     * staged through a scratch buffer of the "network" memory type */
    struct {
        ucc_mc_buffer_header_t *scratch;
        union { ucc_coll_buffer_info_t info; ucc_coll_buffer_info_v_t info_v; } saved; /* user's original buffer */
        int is_src; /* which side was replaced */
    } asymm;
} ucc_base_coll_args_t;

typedef ucc_status_t (*ucc_base_coll_init_fn_t)(ucc_base_coll_args_t *coll_args, ucc_base_team_t *team, ucc_coll_task_t **task);

enum {
    UCC_COLL_TASK_FLAG_CB                    = UCC_BIT(0),
    UCC_COLL_TASK_FLAG_EXECUTOR              = UCC_BIT(1), /* task needs an ee executor */
    UCC_COLL_TASK_FLAG_TOP_LEVEL             = UCC_BIT(2), /* request handed to the user */
    UCC_COLL_TASK_FLAG_EXECUTOR_STOP         = UCC_BIT(3),
    UCC_COLL_TASK_FLAG_EXECUTOR_DESTROY      = UCC_BIT(4),
    UCC_COLL_TASK_FLAG_IS_SCHEDULE           = UCC_BIT(5),
    UCC_COLL_TASK_FLAG_IS_PIPELINED_SCHEDULE = UCC_BIT(6),
    UCC_COLL_TASK_FLAG_INTERNAL              = UCC_BIT(7), /* service task: no coll-trace */
    UCC_COLL_TASK_FLAG_AUTO_FINALIZE         = UCC_BIT(8), /* finalize() is called at the end of completion */
    UCC_COLL_TASK_FLAG_ARGS_UPDATED          = UCC_BIT(9), /* a pipelined parent changed bargs.args since the last post */
};

struct ucc_coll_task {
    ucc_coll_req_t                     super;   /* user-visible status; MUST be first */
    ucc_status_t                       status;  /* internal status driven by progress */
    uint32_t                           flags;
    ucc_base_coll_args_t               bargs;
    ucc_base_team_t                   *team;    /* CL/TL team that owns the task */
    ucc_schedule_t                    *schedule;/* parent schedule, if any */
    ucc_coll_post_fn_t                 post;
    ucc_coll_triggered_post_setup_fn_t triggered_post_setup;
    ucc_coll_triggered_post_fn_t       triggered_post;
    ucc_coll_progress_fn_t             progress;
    ucc_coll_finalize_fn_t             finalize;
    ucc_coll_callback_t                cb;
    ucc_ee_h                           ee;
    ucc_ev_t                          *ev;
    void                              *ee_task;
    ucc_coll_task_t                   *triggered_task;
    ucc_event_manager_t                em;
    ucc_ee_executor_t                 *executor;
    union { ucc_list_link_t list_elem; ucc_lf_queue_elem_t lf_elem; };
    volatile uint32_t                  n_deps, n_deps_satisfied;
    uint32_t                           n_deps_base;
    double                             start_time; /* set at post when a timeout is requested */
    double                             timeout;
    uint32_t                           seq_num;
    uint32_t                           generation; /* bumped each time a pipelined fragment re-arms the task */
    void                              *progress_queue; /* ucc_progress_queue_t the task is enqueued to */
    ucc_base_coll_init_fn_t            init_fn; /* algorithm entry that built the task (set by ucc_coll_init_as): a TL that captures its arguments at
                                                 * init rebuilds the task through it when a pipelined parent set UCC_COLL_TASK_FLAG_ARGS_UPDATED */
};

struct ucc_schedule {
    ucc_coll_task_t  super;
    volatile uint32_t n_completed_tasks;
    uint32_t         n_tasks;
    struct ucc_context *ctx;
    ucc_coll_task_t *tasks[UCC_SCHEDULE_MAX_TASKS];
};

/* ---- task ---- */
ucc_status_t ucc_coll_task_init(ucc_coll_task_t *task, ucc_base_coll_args_t *args, ucc_base_team_t *team);
void         ucc_coll_task_destruct(ucc_coll_task_t *task);
ucc_status_t ucc_coll_task_get_executor(ucc_coll_task_t *task, ucc_ee_executor_t **exec);
ucc_status_t ucc_task_complete(ucc_coll_task_t *task); /* honours task->status */
ucc_status_t ucc_triggered_post(ucc_ee_h ee, ucc_ev_t *ev, ucc_coll_task_t *task); /* default impl, in core/ucc_coll.c */
ucc_status_t ucc_coll_finalize_default_(ucc_coll_task_t *task);

/* ---- events ---- */
ucc_status_t ucc_event_manager_subscribe(ucc_coll_task_t *parent, ucc_event_t event, ucc_coll_task_t *task,
                                         ucc_task_event_handler_p handler);
ucc_status_t ucc_event_manager_notify(ucc_coll_task_t *parent, ucc_event_t event);

/* ---- schedule ---- */
ucc_status_t ucc_schedule_init(ucc_schedule_t *schedule, ucc_base_coll_args_t *bargs, ucc_base_team_t *team);
ucc_status_t ucc_schedule_add_task(ucc_schedule_t *schedule, ucc_coll_task_t *task);
ucc_status_t ucc_schedule_start(ucc_coll_task_t *task);   /* post fn of a schedule */
ucc_status_t ucc_schedule_finalize(ucc_coll_task_t *task);/* finalizes all sub tasks */
/* post `task` once all its dependencies fired (handler for subscribe) */
ucc_status_t ucc_dependency_handler(ucc_coll_task_t *parent, ucc_coll_task_t *task);
/* convenience: task starts when `parent` emits `event` */
static inline ucc_status_t ucc_task_subscribe_dep(ucc_coll_task_t *parent, ucc_coll_task_t *task, ucc_event_t event)
{
    ucc_status_t st = ucc_event_manager_subscribe(parent, event, task, ucc_dependency_handler);
    task->n_deps++;
    return st;
}
#define UCC_TASK_CORE_CTX(_task) ((_task)->bargs.team->contexts[0])
#endif
