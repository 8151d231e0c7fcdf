#ifndef UCC_TL_SHM_COLL_H_
#define UCC_TL_SHM_COLL_H_
#include "tl_shm_prog.h"
#include "utils/ucc_string.h"
typedef ucc_status_t (*ucc_tl_shm_alg_fn_t)(ucc_tl_shm_task_t *task);
#define SHM_ALG(_n) ucc_status_t ucc_tl_shm_##_n(ucc_tl_shm_task_t *t)
SHM_ALG(barrier_knomial); SHM_ALG(fanin_knomial); SHM_ALG(fanout_knomial);
SHM_ALG(bcast_knomial); SHM_ALG(bcast_sag); SHM_ALG(bcast_dbt);
SHM_ALG(allreduce_knomial); SHM_ALG(allreduce_sra); SHM_ALG(allreduce_ring); SHM_ALG(allreduce_dbt); SHM_ALG(allreduce_sliding_window);
SHM_ALG(reduce_knomial); SHM_ALG(reduce_srg); SHM_ALG(reduce_dbt);
SHM_ALG(reduce_scatter_ring); SHM_ALG(reduce_scatter_knomial); SHM_ALG(reduce_scatterv_ring);
SHM_ALG(allgather_knomial); SHM_ALG(allgather_ring); SHM_ALG(allgather_neighbor); SHM_ALG(allgather_bruck);
SHM_ALG(allgather_sparbit); SHM_ALG(allgather_linear); SHM_ALG(allgather_batched);
SHM_ALG(allgatherv_ring); SHM_ALG(allgatherv_knomial); SHM_ALG(allgatherv_linear);
SHM_ALG(alltoall_pairwise); SHM_ALG(alltoall_bruck); SHM_ALG(alltoallv_pairwise); SHM_ALG(alltoallv_hybrid); SHM_ALG(alltoall_onesided); SHM_ALG(alltoallv_onesided);
SHM_ALG(gather_knomial); SHM_ALG(gather_linear); SHM_ALG(gatherv_linear);
SHM_ALG(scatter_knomial); SHM_ALG(scatter_linear); SHM_ALG(scatterv_linear);
ucc_status_t ucc_tl_shm_bcast_knomial_prog(ucc_tl_shm_task_t *t, void *buf, size_t len, ucc_memory_type_t mt, ucc_rank_t root, unsigned radix);
ucc_status_t ucc_tl_shm_service_allreduce_prog(ucc_tl_shm_task_t *t, void *sbuf, void *rbuf, size_t count);
ucc_status_t ucc_tl_shm_service_allgather_prog(ucc_tl_shm_task_t *t, void *sbuf, void *rbuf, size_t msgsize);
#endif
