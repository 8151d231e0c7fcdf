/* Internal (service) collectives over the service TL team, with subsets,
 * plus an OOB adaptor built on them (contract: reference core/ucc_service_coll.{h,c}). */
#ifndef UCC_SERVICE_COLL_H_
#define UCC_SERVICE_COLL_H_
#include "ucc_team.h"
#include "schedule/ucc_schedule.h"

struct ucc_service_coll_req {
    ucc_coll_task_t *task;
    ucc_team_t      *team;
    ucc_subset_t     subset;   /* subset in service-team rank space */
    int              subset_owned;
    void            *data;
};
ucc_status_t ucc_service_allreduce(ucc_team_t *team, void *sbuf, void *rbuf, ucc_datatype_t dt, size_t count,
                                   ucc_reduction_op_t op, ucc_subset_t subset, ucc_service_coll_req_t **req);
ucc_status_t ucc_service_allgather(ucc_team_t *team, void *sbuf, void *rbuf, size_t msgsize, ucc_subset_t subset,
                                   ucc_service_coll_req_t **req);
ucc_status_t ucc_service_bcast(ucc_team_t *team, void *buf, size_t msgsize, ucc_rank_t root, ucc_subset_t subset,
                               ucc_service_coll_req_t **req);
ucc_status_t ucc_service_coll_test(ucc_service_coll_req_t *req);
ucc_status_t ucc_service_coll_finalize(ucc_service_coll_req_t *req);
/* oob whose allgather runs over `subset` (team ranks) of `team` */
ucc_status_t ucc_internal_oob_init(ucc_team_t *team, ucc_subset_t subset, ucc_team_oob_coll_t *oob);
void         ucc_internal_oob_finalize(ucc_team_oob_coll_t *oob);
/* consistency check of (datatype, mem type) across ranks for v-collectives; wraps `task` in a schedule */
ucc_status_t ucc_service_dt_check(ucc_team_t *team, int64_t *values, ucc_subset_t subset, ucc_coll_task_t **check_task);
#endif
