/* Pipelined schedule: a large collective is cut into n_frags_total
 * fragments executed by up to `n_frags` (<=4) identical fragment schedules
 * that are re-armed round-robin.  Fragment orderings:
 *   PARALLEL   - fragments are independent
 *   ORDERED    - task j of fragment g starts after task j of fragment g-1 STARTED
 *   SEQUENTIAL - task j of fragment g starts after task j of fragment g-1 COMPLETED
 * (concept: reference schedule/ucc_schedule_pipelined.h:12-106; the
 * cross-fragment dependency bookkeeping here is per (slot, task) fire
 * tracking rather than n_deps_base arithmetic). */
#ifndef UCC_SCHEDULE_PIPELINED_H_
#define UCC_SCHEDULE_PIPELINED_H_
#include "ucc_schedule.h"
#include "utils/ucc_parser.h"

#define UCC_SCHEDULE_PIPELINED_MAX_FRAGS 4

typedef struct ucc_schedule_pipelined ucc_schedule_pipelined_t;
typedef ucc_status_t (*ucc_schedule_frag_init_fn_t)(ucc_base_coll_args_t *coll_args, ucc_schedule_pipelined_t *sp,
                                                    ucc_base_team_t *team, ucc_schedule_t **frag);
typedef ucc_status_t (*ucc_schedule_frag_setup_fn_t)(ucc_schedule_pipelined_t *sp, ucc_schedule_t *frag, int frag_num);

struct ucc_schedule_pipelined {
    ucc_schedule_t               super;
    ucc_schedule_t              *frags[UCC_SCHEDULE_PIPELINED_MAX_FRAGS];
    int                          n_frags;          /* fragment schedules (pipeline depth) */
    int                          n_frags_total;    /* fragments to execute per post */
    int                          n_frags_started;
    int                          n_frags_completed;
    ucc_pipeline_order_t         order;
    ucc_schedule_frag_setup_fn_t frag_setup;
    int                          slot_global[UCC_SCHEDULE_PIPELINED_MAX_FRAGS]; /* global frag index running in slot, -1 idle */
    uint8_t                      fired[UCC_SCHEDULE_PIPELINED_MAX_FRAGS][UCC_SCHEDULE_MAX_TASKS];
    uint8_t                      slot_idle[UCC_SCHEDULE_PIPELINED_MAX_FRAGS];   /* the slot's fragment completed (or none was launched yet) */
    int                         *done_ptr;  /* set while a frame is launching fragments: where a nested completion of the LAST fragment is reported */
    ucc_recursive_spinlock_t     lock;
};

ucc_status_t ucc_schedule_pipelined_init(ucc_base_coll_args_t *coll_args, ucc_base_team_t *team,
                                         ucc_schedule_frag_init_fn_t frag_init, ucc_schedule_frag_setup_fn_t frag_setup,
                                         int n_frags, int n_frags_total, ucc_pipeline_order_t order,
                                         ucc_schedule_pipelined_t *schedule);
ucc_status_t ucc_schedule_pipelined_post(ucc_coll_task_t *task);
ucc_status_t ucc_schedule_pipelined_finalize(ucc_coll_task_t *task);
#endif
