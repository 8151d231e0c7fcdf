/* tl/shm: host transport over POSIX shared memory.
 *
 * Plays the role tl_ucp has in the reference (host + fallback TL, the only
 * TL able to run the core's service collectives) without UCX: every context
 * owns one multi-producer receive ring in a shm segment; peers (other
 * processes or other contexts of the same process) push fixed-size cells.
 * Tag matching, unexpected-message stash, zero-copy rendezvous inside one
 * process.  Collective algorithms are expressed as small "programs" of
 * SEND/RECV/REDUCE/COPY/WAIT steps executed by one generic engine
 * (tl_shm_prog.c), so an algorithm is a plain loop that emits steps.
 */
#ifndef UCC_TL_SHM_H_
#define UCC_TL_SHM_H_
#include "components/tl/ucc_tl.h"
#include "components/mc/ucc_mc.h"
#include "components/ec/ucc_ec.h"
#include "core/ucc_context.h"
#include "core/ucc_team.h"
#include "utils/ucc_mpool.h"
#include "utils/ucc_hash.h"
#include "utils/ucc_atomic.h"

#define UCC_TL_SHM_DEFAULT_SCORE 10
#define UCC_TL_SHM_SUPPORTED_COLLS UCC_COLL_TYPE_ALL
#define UCC_TL_SHM_NAME_MAX 64

/* ---- wire format ---- */
typedef enum { SHM_CELL_DATA = 1, SHM_CELL_RTS = 2, SHM_CELL_ACK = 3 } shm_cell_type_t;
typedef struct shm_cell_hdr {
    volatile uint64_t seq;      /* Vyukov sequence: == pos when free, pos+1 when filled */
    uint64_t          tag;
    uint64_t          src_ep;   /* unique id of the sending context */
    uint64_t          total_len;
    uint64_t          offset;   /* DATA: byte offset; RTS: source pointer; ACK: sender request pointer */
    uint64_t          cookie;   /* RTS: sender request pointer */
    uint32_t          len;      /* payload bytes in this cell */
    uint16_t          type;
    uint16_t          src_mt;   /* RTS: memory type of the source buffer */
} shm_cell_hdr_t;

typedef struct shm_ring_hdr {
    uint64_t          magic;
    uint32_t          n_cells, cell_size; /* cell_size includes the header */
    volatile uint64_t head UCC_ALIGNED(64); /* consumer position */
    volatile uint64_t tail UCC_ALIGNED(64); /* producer position */
    char              pad[48];
} shm_ring_hdr_t;
#define SHM_RING_MAGIC 0x5543435f53484d31ull
#define SHM_CELL(_r, _i) ((shm_cell_hdr_t *)((char *)(_r) + sizeof(shm_ring_hdr_t) + (size_t)((_i) & ((_r)->n_cells - 1)) * (_r)->cell_size))
#define SHM_CELL_PAYLOAD(_c) ((void *)((shm_cell_hdr_t *)(_c) + 1))

/* address published through the core address exchange */
typedef struct ucc_tl_shm_addr { uint64_t ep_id; uint64_t host_hash; int32_t pid; uint32_t n_cells, cell_size; char name[UCC_TL_SHM_NAME_MAX];
                                 uint64_t probe_addr, probe_val; /* a word in the owner's address space: peers verify cross-memory-attach with it */ } ucc_tl_shm_addr_t;

typedef struct ucc_tl_shm_ep { ucc_tl_shm_addr_t addr; shm_ring_hdr_t *ring; size_t map_len; int same_process;
                               int cma_ok; /* process_vm_readv() into the peer works: large host messages are read in one copy */ } ucc_tl_shm_ep_t;

/* ---- config ---- */
typedef struct ucc_tl_shm_lib_config { ucc_tl_lib_config_t super; } ucc_tl_shm_lib_config_t;
typedef struct ucc_tl_shm_context_config {
    ucc_tl_context_config_t super;
    unsigned  n_cells;           /* ring depth (power of two) */
    size_t    cell_payload;      /* eager payload per cell */
    size_t    rndv_thresh;       /* rendezvous (single copy) threshold: same process, or other processes through CMA */
    int       cma;               /* ternary: cross-memory-attach rendezvous between processes */
    ucc_mrange_uint_t allreduce_kn_radix, allreduce_sra_kn_radix, reduce_srg_kn_radix, reduce_scatter_kn_radix, bcast_kn_radix, bcast_sag_kn_radix, reduce_kn_radix, barrier_kn_radix, allgather_kn_radix, gather_kn_radix;
    unsigned  kn_radix;          /* > 0: default of every k-nomial radix that is left on auto */
    unsigned  fanin_kn_radix, fanout_kn_radix, scatter_kn_radix;   /* 0 / auto: follow BARRIER_KN_RADIX (fanin, fanout) / GATHER_KN_RADIX (scatter) */
    unsigned  alltoall_pairwise_num_posts, alltoallv_pairwise_num_posts, allgather_batched_num_posts, gatherv_linear_num_posts, scatterv_linear_num_posts;
    int       reduce_scatter_ring_bidirectional, reduce_scatterv_ring_bidirectional;
    int       use_topo;          /* ternary: may the transport look at the team topology */
    int       ranks_reordering;  /* ring algorithms walk the members host by host (needs the topology) */
    size_t    allreduce_sliding_win_buf_size; /* sliding_window allreduce: bytes fetched from a peer and reduced per step */
    size_t    alltoallv_hybrid_thresh;   /* alltoallv `hybrid`: messages up to this size ride the Bruck rounds */
    int       reduce_avg_pre_op;
    ucc_pipeline_params_t allreduce_sra_kn_pipeline, reduce_srg_kn_pipeline;
} ucc_tl_shm_context_config_t;

typedef struct ucc_tl_shm_lib { ucc_tl_lib_t super; } ucc_tl_shm_lib_t;

/* ---- p2p requests ---- */
typedef struct shm_req {
    ucc_list_link_t    list;
    int                is_send, done;
    uint64_t           tag, peer_ep;
    void              *buf;
    size_t             len, progressed; /* send: bytes pushed; recv: bytes received */
    ucc_memory_type_t  mt;
    ucc_tl_shm_ep_t   *ep;             /* send: destination */
    int                rndv;           /* send: waiting for ACK */
} shm_req_t;
/* `done` is set under the context lock by whoever drains the ring and polled without it by the owning task (THREAD_MULTIPLE) */
static inline void shm_req_complete(shm_req_t *r) { __atomic_store_n(&r->done, 1, __ATOMIC_RELEASE); }
static inline int  shm_req_is_done(const shm_req_t *r) { return __atomic_load_n(&r->done, __ATOMIC_ACQUIRE); }

typedef struct shm_unexp { ucc_list_link_t list; uint64_t tag, src_ep, total_len, received; int is_rts; void *rts_ptr; uint64_t rts_cookie; uint16_t rts_mt; void *data; } shm_unexp_t;

typedef struct ucc_tl_shm_context {
    ucc_tl_context_t            super;
    ucc_tl_shm_context_config_t cfg;
    ucc_tl_shm_addr_t           addr;
    shm_ring_hdr_t             *ring;     /* my receive ring */
    size_t                      ring_len;
    ucc_hash_t                  eps;      /* ep_id -> ucc_tl_shm_ep_t* */
    ucc_list_link_t             posted_recvs, pending_sends, rndv_sends, unexpected, pending_acks;
    ucc_mpool_t                 req_mp, task_mp;
    ucc_thread_mode_t           tm;
    ucc_recursive_spinlock_t    lock;     /* THREAD_MULTIPLE */
    uint64_t                    n_progress;
} ucc_tl_shm_context_t;

typedef struct ucc_tl_shm_team {
    ucc_tl_team_t     super;
    ucc_tl_shm_ep_t **eps;       /* team rank -> endpoint (lazy) */
    uint32_t          seq_num;   /* collective sequence on this team */
    uint64_t          tag_base;  /* team identity bits of the tag */
    struct shm_svc_cnt { uint64_t hash; uint32_t cnt; } *svc; /* per-subset service collective counters */
    unsigned          n_svc;
    ucc_rank_t       *ring_order; /* RANKS_REORDERING: ring position -> team rank, members of one host (then socket) adjacent; NULL: rank order */
    ucc_rank_t       *ring_pos;   /* team rank -> ring position */
} ucc_tl_shm_team_t;

#define SHM_CTX(_team) ucc_derived_of((_team)->super.super.context, ucc_tl_shm_context_t)

/* tag layout: [team id 16][scope 4][scope_id 4][membership hash 8][coll seq 16][step 16] */
static inline uint64_t shm_make_tag(const ucc_tl_shm_team_t *t, uint32_t coll_seq, uint32_t step)
{ return t->tag_base | ((uint64_t)(coll_seq & 0xffff) << 16) | (uint64_t)(step & 0xffff); }

/* p2p (tl_shm_p2p.c) */
ucc_status_t ucc_tl_shm_ctx_p2p_init(ucc_tl_shm_context_t *ctx);
void         ucc_tl_shm_ctx_p2p_cleanup(ucc_tl_shm_context_t *ctx);
ucc_status_t ucc_tl_shm_get_ep(ucc_tl_shm_team_t *team, ucc_rank_t team_rank, ucc_tl_shm_ep_t **ep);
ucc_status_t ucc_tl_shm_send_nb(ucc_tl_shm_team_t *team, ucc_rank_t dst, uint64_t tag, void *buf, size_t len, ucc_memory_type_t mt, shm_req_t **req);
ucc_status_t ucc_tl_shm_recv_nb(ucc_tl_shm_team_t *team, ucc_rank_t src, uint64_t tag, void *buf, size_t len, ucc_memory_type_t mt, shm_req_t **req);
void         ucc_tl_shm_req_free(ucc_tl_shm_context_t *ctx, shm_req_t *req);
int          ucc_tl_shm_can_get(ucc_tl_shm_team_t *team, ucc_rank_t rank); /* direct (one-sided) reads from that rank's memory possible? */
ucc_status_t ucc_tl_shm_get(ucc_tl_shm_team_t *team, ucc_rank_t rank, void *dst, ucc_memory_type_t dmt, uint64_t remote_addr, size_t len);
int          ucc_tl_shm_can_put(ucc_tl_shm_team_t *team, ucc_rank_t rank); /* direct writes into that rank's host memory possible? */
ucc_status_t ucc_tl_shm_put(ucc_tl_shm_team_t *team, ucc_rank_t rank, const void *src, ucc_memory_type_t smt, uint64_t remote_addr, size_t len);

/* tl/shm part of a ucc_mem_map() handle (tl_shm_memh.c) */
#define UCC_TL_SHM_MEMH_MAGIC 0x53484d4d454d4831ull /* "SHMMEMH1" */
typedef struct shm_memh { uint64_t magic, host_hash, addr, len; int32_t pid, imported, reachable, pad; } shm_memh_t;
ucc_status_t ucc_tl_shm_mem_map(const ucc_base_context_t *ctx, ucc_mem_map_mode_t mode, void *memh, void *tl_h);
ucc_status_t ucc_tl_shm_mem_unmap(const ucc_base_context_t *ctx, ucc_mem_map_mode_t mode, void *tl_h);
ucc_status_t ucc_tl_shm_memh_pack(const ucc_base_context_t *ctx, ucc_mem_map_mode_t mode, void *tl_h, void **pack_buffer);
ucc_status_t ucc_tl_shm_progress(void *ctx); /* registered with the core context */

extern ucc_tl_iface_t ucc_tl_shm;
#endif
