/* Execution engine handle: a user stream/thread plus in/out event queues
 * (contract: reference core/ucc_ee.{h,c}). */
#ifndef UCC_EE_H_
#define UCC_EE_H_
#include <ucc/api/ucc.h>
#include "utils/ucc_atomic.h"
#include "utils/ucc_list.h"

typedef struct ucc_event_desc { ucc_queue_elem_t queue; ucc_ev_t ev; } ucc_event_desc_t;
typedef struct ucc_ee {
    ucc_team_h       team;
    ucc_ee_type_t    ee_type;
    size_t           ee_context_size;
    void            *ee_context;
    ucc_spinlock_t   lock;
    ucc_queue_head_t event_in_queue;
    ucc_queue_head_t event_out_queue;
} ucc_ee_t;
ucc_status_t ucc_ee_get_event_internal(ucc_ee_h ee, ucc_ev_t **ev, ucc_queue_head_t *queue);
ucc_status_t ucc_ee_set_event_internal(ucc_ee_h ee, ucc_ev_t *ev, ucc_queue_head_t *queue);
#endif
