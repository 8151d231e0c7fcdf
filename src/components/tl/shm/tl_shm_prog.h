/* Collective "programs": an algorithm emits a flat list of steps at init
 * time; one engine executes them.  Steps between two WAITs run concurrently
 * (sends/recvs are posted asynchronously); REDUCE and COPY go through the
 * EC executor of the collective's memory type, so the same algorithm works
 * for host and (as a fallback path) device buffers. */
#ifndef UCC_TL_SHM_PROG_H_
#define UCC_TL_SHM_PROG_H_
#include "tl_shm.h"

typedef enum { SHM_OP_SEND, SHM_OP_RECV, SHM_OP_WAIT, SHM_OP_REDUCE, SHM_OP_COPY,
               SHM_OP_GET /* one-sided read of `len` bytes from address *(uint64_t *)src1 in `peer` (same process: pointer, other process: CMA) into dst */,
               SHM_OP_PUT /* one-sided write of `len` bytes at dst (local) to the address src2 in `peer`'s registered memory */ } shm_op_type_t;

typedef struct shm_op {
    uint8_t           type;
    uint8_t           with_alpha; /* REDUCE: multiply the result by task->alpha (AVG) */
    uint16_t          step;       /* SEND/RECV: message id both sides agree on */
    ucc_rank_t        peer;       /* SEND/RECV: rank inside the (sub)team */
    void             *dst;        /* RECV/REDUCE/COPY destination; SEND source */
    void             *src1, *src2;/* REDUCE operands (src2 may be NULL: scale only); COPY source in src1 */
    size_t            len;        /* SEND/RECV/COPY bytes; REDUCE element count */
    ucc_memory_type_t mt, mt_src; /* memory type of dst (and of src1 for COPY) */
} shm_op_t;

typedef struct ucc_tl_shm_task {
    ucc_coll_task_t    super;
    ucc_tl_shm_team_t *team;
    /* virtual team the algorithm runs on (whole team, a subset, or an active set) */
    ucc_rank_t         vrank, vsize;
    ucc_ep_map_t       vmap;          /* vrank -> TL team rank */
    uint32_t           coll_seq;
    /* program */
    shm_op_t          *ops;
    unsigned           n_ops, cap_ops, pc;
    shm_req_t        **reqs;          /* outstanding p2p requests since the last WAIT */
    unsigned           n_reqs, cap_reqs;
    ucc_ee_executor_task_t *etask;    /* in-flight executor task */
    ucc_ee_executor_t *own_exec;      /* executor created by the task itself (service colls) */
    /* reduction parameters */
    ucc_datatype_t     dt;
    ucc_reduction_op_t op;
    double             alpha;
    /* scratch owned by the task */
    ucc_mc_buffer_header_t *scratch[4];
    unsigned           n_scratch;
    void              *host_copy;     /* malloc'ed aux (counts tables etc.) */
    /* program builder: kept so that a pipelined parent can change bargs.args between posts (UCC_COLL_TASK_FLAG_ARGS_UPDATED) */
    ucc_status_t     (*build)(struct ucc_tl_shm_task *t);
    ucc_status_t     (*setup)(struct ucc_tl_shm_task *t);
} ucc_tl_shm_task_t;

/* builder */
ucc_status_t shm_prog_send(ucc_tl_shm_task_t *t, ucc_rank_t peer, const void *buf, size_t len, ucc_memory_type_t mt, unsigned step);
ucc_status_t shm_prog_recv(ucc_tl_shm_task_t *t, ucc_rank_t peer, void *buf, size_t len, ucc_memory_type_t mt, unsigned step);
ucc_status_t shm_prog_wait(ucc_tl_shm_task_t *t);
/* dst <- len bytes at the remote address that will be stored in *remote_addr_slot by the time the step runs */
ucc_status_t shm_prog_get(ucc_tl_shm_task_t *t, ucc_rank_t peer, void *dst, const uint64_t *remote_addr_slot, size_t len, ucc_memory_type_t mt);
ucc_status_t shm_prog_get_off(ucc_tl_shm_task_t *t, ucc_rank_t peer, void *dst, const uint64_t *remote_addr_slot, size_t off, size_t len, ucc_memory_type_t mt);
/* len bytes at `src` -> address `remote_addr` inside peer's host memory (known at build time: registered buffers) */
ucc_status_t shm_prog_put(ucc_tl_shm_task_t *t, ucc_rank_t peer, const void *src, uint64_t remote_addr, size_t len, ucc_memory_type_t smt);
ucc_status_t shm_prog_reduce(ucc_tl_shm_task_t *t, void *dst, const void *src1, const void *src2, size_t count, ucc_memory_type_t mt, int with_alpha);
ucc_status_t shm_prog_copy(ucc_tl_shm_task_t *t, void *dst, const void *src, size_t len, ucc_memory_type_t dmt, ucc_memory_type_t smt);
ucc_status_t shm_task_scratch(ucc_tl_shm_task_t *t, size_t len, ucc_memory_type_t mt, void **ptr);

/* life cycle */
ucc_status_t ucc_tl_shm_task_alloc(ucc_base_coll_args_t *bargs, ucc_base_team_t *team, ucc_tl_shm_task_t **task);
ucc_status_t ucc_tl_shm_task_post(ucc_coll_task_t *task);
void         ucc_tl_shm_task_progress(ucc_coll_task_t *task);
ucc_status_t ucc_tl_shm_task_finalize(ucc_coll_task_t *task);
#endif
