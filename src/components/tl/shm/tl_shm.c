/* tl/shm component glue: lib / context / team objects, algorithm tables,
 * score publication, service collectives. */
#include "tl_shm_coll.h"
#include "core/ucc_global_opts.h"
#include "core/ucc_team.h"
#include "schedule/ucc_schedule_pipelined.h"
#include "components/topo/ucc_topo.h"
#include <strings.h>

/* ------------------------------------------------------------------ */
/* config                                                              */
/* ------------------------------------------------------------------ */
static ucc_config_field_t tl_shm_lib_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_tl_shm_lib_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_tl_lib_config_table)}, {NULL}};

static ucc_config_field_t tl_shm_context_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_tl_shm_context_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_tl_context_config_table)},
    {"N_CELLS", "256", "Depth of the per-context receive ring (rounded up to a power of two)", ucc_offsetof(ucc_tl_shm_context_config_t, n_cells), UCC_CONFIG_TYPE_UINT},
    {"CELL_SIZE", "8K", "Eager payload carried by one ring cell", ucc_offsetof(ucc_tl_shm_context_config_t, cell_payload), UCC_CONFIG_TYPE_MEMUNITS},
    {"CMA", "try", "Cross-memory attach (process_vm_readv) rendezvous between processes of one host: large host messages are copied once, straight from the sender's buffer",
     ucc_offsetof(ucc_tl_shm_context_config_t, cma), UCC_CONFIG_TYPE_TERNARY},
    {"RNDV_THRESH", "16K", "Messages of at least this size are copied directly from the source buffer (same process: always; other processes: through CMA)",
     ucc_offsetof(ucc_tl_shm_context_config_t, rndv_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {"ALLREDUCE_KN_RADIX", "auto", "Radix of the recursive k-nomial allreduce (per msg range)", ucc_offsetof(ucc_tl_shm_context_config_t, allreduce_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"ALLREDUCE_SRA_KN_RADIX", "auto", "Radix of the scatter-reduce-allgather (SRA) k-nomial allreduce (per msg range)", ucc_offsetof(ucc_tl_shm_context_config_t, allreduce_sra_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"REDUCE_SRG_KN_RADIX", "auto", "Radix of the scatter-reduce-gather (SRG) k-nomial reduce (per msg range)", ucc_offsetof(ucc_tl_shm_context_config_t, reduce_srg_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"REDUCE_SCATTER_KN_RADIX", "auto", "Radix of the k-nomial reduce_scatter (teams of radix^m ranks; other power-of-two teams use recursive halving)", ucc_offsetof(ucc_tl_shm_context_config_t, reduce_scatter_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"ALLREDUCE_SLIDING_WIN_BUF_SIZE", "512K", "Window of the sliding_window allreduce: bytes of the own slice fetched from each peer and reduced per step",
     ucc_offsetof(ucc_tl_shm_context_config_t, allreduce_sliding_win_buf_size), UCC_CONFIG_TYPE_MEMUNITS},
    {"BCAST_KN_RADIX", "auto", "Radix of the k-nomial tree bcast", ucc_offsetof(ucc_tl_shm_context_config_t, bcast_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"REDUCE_KN_RADIX", "auto", "Radix of the k-nomial tree reduce", ucc_offsetof(ucc_tl_shm_context_config_t, reduce_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"BARRIER_KN_RADIX", "auto", "Radix of the k-nomial barrier / fanin / fanout", ucc_offsetof(ucc_tl_shm_context_config_t, barrier_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"ALLGATHER_KN_RADIX", "auto", "Radix of the k-nomial allgather(v)", ucc_offsetof(ucc_tl_shm_context_config_t, allgather_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"GATHER_KN_RADIX", "auto", "Radix of the k-nomial gather / scatter", ucc_offsetof(ucc_tl_shm_context_config_t, gather_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"KN_RADIX", "0", "Radix of every k-nomial algorithm whose own *_KN_RADIX is left on auto (0: built-in defaults)", ucc_offsetof(ucc_tl_shm_context_config_t, kn_radix), UCC_CONFIG_TYPE_UINT},
    {"FANIN_KN_RADIX", "auto", "Radix of the k-nomial fanin (auto: BARRIER_KN_RADIX)", ucc_offsetof(ucc_tl_shm_context_config_t, fanin_kn_radix), UCC_CONFIG_TYPE_UINT},
    {"FANOUT_KN_RADIX", "auto", "Radix of the k-nomial fanout (auto: BARRIER_KN_RADIX)", ucc_offsetof(ucc_tl_shm_context_config_t, fanout_kn_radix), UCC_CONFIG_TYPE_UINT},
    {"SCATTER_KN_RADIX", "auto", "Radix of the k-nomial scatter (auto: GATHER_KN_RADIX)", ucc_offsetof(ucc_tl_shm_context_config_t, scatter_kn_radix), UCC_CONFIG_TYPE_UINT},
    {"BCAST_SAG_KN_RADIX", "auto", "Radix of the scatter tree of the scatter-allgather bcast", ucc_offsetof(ucc_tl_shm_context_config_t, bcast_sag_kn_radix), UCC_CONFIG_TYPE_UINT_RANGED},
    {"ALLTOALLV_PAIRWISE_NUM_POSTS", "auto", "Maximum number of outstanding send/recv pairs in pairwise alltoallv (auto/0: unlimited)",
     ucc_offsetof(ucc_tl_shm_context_config_t, alltoallv_pairwise_num_posts), UCC_CONFIG_TYPE_UINT},
    {"ALLGATHER_BATCHED_NUM_POSTS", "auto", "Maximum number of outstanding send/recv pairs in the batched allgather (auto: 4, 0: unlimited)",
     ucc_offsetof(ucc_tl_shm_context_config_t, allgather_batched_num_posts), UCC_CONFIG_TYPE_UINT},
    {"GATHERV_LINEAR_NUM_POSTS", "0", "Maximum number of receives the root of a linear gatherv keeps outstanding (0: all)",
     ucc_offsetof(ucc_tl_shm_context_config_t, gatherv_linear_num_posts), UCC_CONFIG_TYPE_UINT},
    {"SCATTERV_LINEAR_NUM_POSTS", "0", "Maximum number of sends the root of a linear scatterv keeps outstanding (0: all)",
     ucc_offsetof(ucc_tl_shm_context_config_t, scatterv_linear_num_posts), UCC_CONFIG_TYPE_UINT},
    {"REDUCE_SCATTER_RING_BIDIRECTIONAL", "n", "Ring reduce_scatter: the two halves of every block travel around two inverted rings concurrently",
     ucc_offsetof(ucc_tl_shm_context_config_t, reduce_scatter_ring_bidirectional), UCC_CONFIG_TYPE_BOOL},
    {"REDUCE_SCATTERV_RING_BIDIRECTIONAL", "n", "Ring reduce_scatterv: the two halves of every block travel around two inverted rings concurrently",
     ucc_offsetof(ucc_tl_shm_context_config_t, reduce_scatterv_ring_bidirectional), UCC_CONFIG_TYPE_BOOL},
    {"ALLTOALL_PAIRWISE_NUM_POSTS", "auto", "Maximum number of outstanding send/recv pairs in pairwise alltoall(v) (auto/0: unlimited)",
     ucc_offsetof(ucc_tl_shm_context_config_t, alltoall_pairwise_num_posts), UCC_CONFIG_TYPE_UINT},
    {"ALLTOALLV_HYBRID_THRESH", "256", "alltoallv algorithm `hybrid`: messages of at most this many bytes are aggregated into log2(N) Bruck rounds, larger ones go pairwise",
     ucc_offsetof(ucc_tl_shm_context_config_t, alltoallv_hybrid_thresh), UCC_CONFIG_TYPE_MEMUNITS},
    {"USE_TOPO", "try", "Allow the transport to use the team topology (y: a team without topology is refused)", ucc_offsetof(ucc_tl_shm_context_config_t, use_topo), UCC_CONFIG_TYPE_TERNARY},
    {"RANKS_REORDERING", "y", "Ring algorithms visit the members host by host (and socket by socket) instead of in rank order; needs USE_TOPO",
     ucc_offsetof(ucc_tl_shm_context_config_t, ranks_reordering), UCC_CONFIG_TYPE_BOOL},
    {"REDUCE_AVG_PRE_OP", "n", "Apply the 1/N scaling of AVG before (y) or after (n) the reduction", ucc_offsetof(ucc_tl_shm_context_config_t, reduce_avg_pre_op), UCC_CONFIG_TYPE_BOOL},
    {"REDUCE_SRG_KN_PIPELINE", "n", "Pipelining of the SRG reduce: thresh=<size>:fragsize=<size>:nfrags=<n>:pdepth=<n>:<parallel|ordered|sequential>", ucc_offsetof(ucc_tl_shm_context_config_t, reduce_srg_kn_pipeline), UCC_CONFIG_TYPE_PIPELINE_PARAMS},
    {"ALLREDUCE_SRA_KN_PIPELINE", "n", "Pipelining of the SRA allreduce", ucc_offsetof(ucc_tl_shm_context_config_t, allreduce_sra_kn_pipeline), UCC_CONFIG_TYPE_PIPELINE_PARAMS},
    {NULL}};

/* ------------------------------------------------------------------ */
/* algorithm tables                                                    */
/* ------------------------------------------------------------------ */
typedef struct shm_alg { const char *name; const char *desc; ucc_tl_shm_alg_fn_t fn; } shm_alg_t;
#define A(_n, _d, _f) {_n, _d, ucc_tl_shm_##_f}
static const shm_alg_t algs_allreduce[] = {A("knomial", "recursive k-nomial exchange (latency)", allreduce_knomial), A("sra_knomial", "scatter-reduce by recursive vector splitting + allgather, radix k (bandwidth)", allreduce_sra),
                                           A("dbt", "double binary tree reduce + bcast", allreduce_dbt), A("ring", "ring reduce-scatter + ring allgather", allreduce_ring),
                                           A("sliding_window", "one-sided: windows of the own slice are read from the peers' buffers (pointer / CMA) and reduced", allreduce_sliding_window), {NULL}};
static const shm_alg_t algs_allgather[] = {A("knomial", "recursive k-ing of block sets (any team size: extra ranks through proxies)", allgather_knomial), A("ring", "ring", allgather_ring), A("neighbor", "neighbor exchange (even team size)", allgather_neighbor),
                                           A("bruck", "O(log N) Bruck allgather", allgather_bruck), A("sparbit", "O(log N) data-locality aware allgather", allgather_sparbit),
                                           A("linear", "everyone sends to everyone", allgather_linear), A("batched", "linear with bounded outstanding messages", allgather_batched), {NULL}};
static const shm_alg_t algs_allgatherv[] = {A("ring", "ring", allgatherv_ring), A("knomial", "recursive k-ing of block sets (extra ranks through proxies)", allgatherv_knomial), A("linear", "everyone sends to everyone", allgatherv_linear), {NULL}};
static const shm_alg_t algs_alltoall[] = {A("pairwise", "pairwise exchange", alltoall_pairwise), A("bruck", "O(log N) Bruck alltoall", alltoall_bruck),
                                          A("onesided", "every rank reads its blocks directly from the peers' source buffers (pointer / CMA)", alltoall_onesided), {NULL}};
static const shm_alg_t algs_alltoallv[] = {A("pairwise", "pairwise exchange", alltoallv_pairwise), A("hybrid", "small messages in fixed slots through Bruck's log2(N) rounds, big ones pairwise", alltoallv_hybrid),
                                           A("onesided", "every rank reads its blocks directly from the peers' source buffers (pointer / CMA)", alltoallv_onesided), {NULL}};
static const shm_alg_t algs_barrier[] = {A("knomial", "k-nomial fanin + fanout", barrier_knomial), {NULL}};
static const shm_alg_t algs_bcast[] = {A("knomial", "k-nomial tree", bcast_knomial), A("sag_knomial", "scatter + ring allgather", bcast_sag), A("dbt", "double binary tree", bcast_dbt), {NULL}};
static const shm_alg_t algs_fanin[] = {A("knomial", "k-nomial tree", fanin_knomial), {NULL}};
static const shm_alg_t algs_fanout[] = {A("knomial", "k-nomial tree", fanout_knomial), {NULL}};
static const shm_alg_t algs_gather[] = {A("knomial", "k-nomial tree", gather_knomial), A("linear", "root receives from everyone", gather_linear), {NULL}};
static const shm_alg_t algs_gatherv[] = {A("linear", "root receives from everyone", gatherv_linear), {NULL}};
static const shm_alg_t algs_reduce[] = {A("knomial", "k-nomial tree", reduce_knomial), A("dbt", "double binary tree", reduce_dbt), A("srg", "k-nomial scatter-reduce + gather", reduce_srg), {NULL}};
static const shm_alg_t algs_reduce_scatter[] = {A("ring", "ring", reduce_scatter_ring), A("knomial", "k-nomial scatter-reduce (teams of radix^m ranks) / recursive halving (power-of-two teams)", reduce_scatter_knomial), {NULL}};
static const shm_alg_t algs_reduce_scatterv[] = {A("ring", "ring", reduce_scatterv_ring), {NULL}};
static const shm_alg_t algs_scatter[] = {A("knomial", "k-nomial tree", scatter_knomial), A("linear", "root sends to everyone", scatter_linear), {NULL}};
static const shm_alg_t algs_scatterv[] = {A("linear", "root sends to everyone", scatterv_linear), {NULL}};
static const shm_alg_t *const shm_algs[UCC_COLL_TYPE_NUM] = {
    algs_allgather, algs_allgatherv, algs_allreduce, algs_alltoall, algs_alltoallv, algs_barrier, algs_bcast, algs_fanin, algs_fanout,
    algs_gather, algs_gatherv, algs_reduce, algs_reduce_scatter, algs_reduce_scatterv, algs_scatter, algs_scatterv};
static ucc_base_coll_alg_info_t shm_alg_info[UCC_COLL_TYPE_NUM][8];

/* per-size defaults (same cut-overs the reference's host TL uses: latency algs below 4K, bandwidth algs above) */
#define SHM_DEFAULT_SELECT_STR \
    "allreduce:0-4k:@knomial#allreduce:4k-inf:@sra_knomial#allgather:0-4k:@bruck#allgather:4k-inf:@ring#" \
    "allgatherv:0-4k:@knomial#allgatherv:4k-inf:@ring#alltoall:0-512:@bruck#bcast:0-32k:@knomial#bcast:32k-inf:@sag_knomial#" \
    "reduce:0-32k:@knomial#reduce:32k-inf:@srg#reduce_scatter:@ring#gather:@knomial#scatter:@knomial"

/* ------------------------------------------------------------------ */
/* collective init                                                     */
/* ------------------------------------------------------------------ */
static ucc_status_t shm_task_setup(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = &t->super.bargs.args;
    ucc_rank_t rank = UCC_TL_TEAM_RANK(t->team);
    int root = (ucc_rank_t)a->root == rank;
    if (ucc_coll_has_reduction(a->coll_type)) {
        t->op = a->op;
        switch (a->coll_type) {
        case UCC_COLL_TYPE_REDUCE: t->dt = root ? a->dst.info.datatype : a->src.info.datatype; break;
        case UCC_COLL_TYPE_REDUCE_SCATTERV: t->dt = a->dst.info_v.datatype; break;
        default: t->dt = a->dst.info.datatype; break;
        }
        t->alpha = 1.0 / (double)t->vsize;
        if (a->op == UCC_OP_MAXLOC || a->op == UCC_OP_MINLOC) return UCC_ERR_NOT_SUPPORTED;
        if (UCC_DT_IS_GENERIC(t->dt) && !UCC_DT_HAS_REDUCE(t->dt)) return UCC_ERR_NOT_SUPPORTED;
        t->super.flags |= UCC_COLL_TASK_FLAG_EXECUTOR;
    }
    if (UCC_COLL_ARGS_ACTIVE_SET(a)) {
        /* active set: the algorithm runs on a strided subset; team-wide sequence numbers must not advance */
        memset(&t->vmap, 0, sizeof(t->vmap));
        t->vmap.type = UCC_EP_MAP_STRIDED; t->vmap.ep_num = a->active_set.size; t->vmap.strided.start = a->active_set.start; t->vmap.strided.stride = a->active_set.stride;
        t->vsize = (ucc_rank_t)a->active_set.size; t->vrank = ucc_ep_map_local_rank(t->vmap, rank);
        if (t->vrank == UCC_RANK_INVALID) return UCC_ERR_INVALID_PARAM;
        t->coll_seq = 0xC000u | ((a->mask & UCC_COLL_ARGS_FIELD_TAG) ? (a->tag & 0x3fff) : 0);   /* (the caller takes the team sequence back) */
    }
    return UCC_OK;
}

static ucc_status_t shm_coll_init_plain(ucc_base_coll_args_t *bargs, ucc_base_team_t *team, ucc_coll_task_t **task_p, ucc_tl_shm_alg_fn_t fn)
{
    ucc_tl_shm_task_t *t;
    ucc_status_t st = ucc_tl_shm_task_alloc(bargs, team, &t);
    if (st != UCC_OK) return st;
    st = shm_task_setup(t);
    t->build = fn;
    if (st == UCC_OK) st = fn(t);
    /* task_alloc advanced the team sequence.  A task that is not created gives it back - and so does an active-set task, which only
     * some members create (the others must not fall behind): exactly ONE decrement in either case.  (It used to be taken back in
     * shm_task_setup AND here when the algorithm then declined the active set - sag_knomial, the default for bcasts >= 32 KB, does -
     * so the members of such a bcast ran one sequence number behind the rest of the team afterwards.) */
    if (st != UCC_OK) { t->team->seq_num--; ucc_tl_shm_task_finalize(&t->super); return st; }
    if (UCC_COLL_ARGS_ACTIVE_SET(&bargs->args)) t->team->seq_num--;
    *task_p = &t->super;
    return UCC_OK;
}

/* ALLREDUCE_SRA_KN_PIPELINE / REDUCE_SRG_KN_PIPELINE (reference allreduce_sra_knomial.c:175-243, reduce_srg_knomial.c): the vector is cut into
 * fragments, `pdepth` SRA / SRG tasks are in flight and re-armed round robin by the pipelined schedule - the scatter-reduce of fragment i+1
 * overlaps the allgather of fragment i, and the scratch of a task is a fragment, not the vector.  A re-armed task rebuilds its step program
 * from the new arguments at post (UCC_COLL_TASK_FLAG_ARGS_UPDATED). */
typedef struct shm_pipe { ucc_schedule_pipelined_t super; ucc_tl_shm_alg_fn_t fn; size_t total, dts, fcount; } shm_pipe_t;
static ucc_status_t shm_frag_post(ucc_coll_task_t *t) { return ucc_schedule_start(t); }
static ucc_status_t shm_frag_finalize(ucc_coll_task_t *t) { ucc_status_t st = ucc_schedule_finalize(t); free(t); return st; }
static void shm_frag_args(const ucc_coll_args_t *a, size_t off_bytes, size_t cnt, ucc_coll_args_t *f)
{
    *f = *a;
    if (f->src.info.buffer) f->src.info.buffer = PTR_OFFSET(f->src.info.buffer, off_bytes);
    if (f->dst.info.buffer) f->dst.info.buffer = PTR_OFFSET(f->dst.info.buffer, off_bytes);
    f->src.info.count = cnt; f->dst.info.count = cnt;
}
static ucc_status_t shm_pipe_frag_init(ucc_base_coll_args_t *b, ucc_schedule_pipelined_t *sp, ucc_base_team_t *team, ucc_schedule_t **frag)
{
    shm_pipe_t *pp = (shm_pipe_t *)sp;
    ucc_base_coll_args_t fb = *b; ucc_coll_task_t *task = NULL; ucc_status_t st;
    ucc_schedule_t *f = (ucc_schedule_t *)calloc(1, sizeof(*f));
    if (!f) return UCC_ERR_NO_MEMORY;
    shm_frag_args(&b->args, 0, pp->fcount, &fb.args); /* the largest fragment */
    ucc_schedule_init(f, &fb, team);
    f->super.post = shm_frag_post; f->super.finalize = shm_frag_finalize;
    st = shm_coll_init_plain(&fb, team, &task, pp->fn);
    if (st == UCC_OK) st = ucc_schedule_add_task(f, task);
    if (st == UCC_OK) st = ucc_task_subscribe_dep(&f->super, task, UCC_EVENT_SCHEDULE_STARTED);
    if (st != UCC_OK) { if (task && !f->n_tasks) task->finalize(task); shm_frag_finalize(&f->super); return st; }
    *frag = f;
    return UCC_OK;
}
static ucc_status_t shm_pipe_frag_setup(ucc_schedule_pipelined_t *sp, ucc_schedule_t *frag, int frag_num)
{
    shm_pipe_t *pp = (shm_pipe_t *)sp;
    size_t off = (size_t)frag_num * pp->fcount, cnt = ucc_min(pp->fcount, pp->total - off);
    ucc_coll_args_t fa, *ta = &frag->tasks[0]->bargs.args;
    shm_frag_args(&sp->super.super.bargs.args, off * pp->dts, cnt, &fa);
    ta->src = fa.src; ta->dst = fa.dst;
    frag->tasks[0]->flags |= UCC_COLL_TASK_FLAG_ARGS_UPDATED;
    return UCC_OK;
}
static ucc_status_t shm_pipe_finalize(ucc_coll_task_t *t) { ucc_status_t st = ucc_schedule_pipelined_finalize(t); free(t); return st; }
/* UCC_ERR_NOT_FOUND: pipelining does not apply, build the plain task */
static ucc_status_t shm_coll_init_pipelined(ucc_base_coll_args_t *b, ucc_base_team_t *team, ucc_coll_task_t **task_p, ucc_tl_shm_alg_fn_t fn)
{
    ucc_tl_shm_context_t *ctx = ucc_derived_of(team->context, ucc_tl_shm_context_t);
    const ucc_pipeline_params_t *pp = fn == ucc_tl_shm_allreduce_sra ? &ctx->cfg.allreduce_sra_kn_pipeline : &ctx->cfg.reduce_srg_kn_pipeline;
    const ucc_coll_args_t *a = &b->args;
    ucc_rank_t N = team->params.size;
    const ucc_coll_buffer_info_t *info = (a->coll_type == UCC_COLL_TYPE_REDUCE && !((ucc_rank_t)a->root == team->params.rank)) ? &a->src.info : &a->dst.info;
    size_t count = info->count, dts, fcount; int n_total, depth; shm_pipe_t *sp; ucc_status_t st;
    if (!pp->n_frags || !pp->pdepth || pp->threshold == UCC_MEMUNITS_INF || pp->threshold == UCC_MEMUNITS_AUTO || UCC_DT_IS_GENERIC(info->datatype)) return UCC_ERR_NOT_FOUND;
    dts = ucc_dt_size(info->datatype);
    if (!dts || count * dts < pp->threshold) return UCC_ERR_NOT_FOUND;
    fcount = (pp->frag_size == UCC_MEMUNITS_INF || pp->frag_size == UCC_MEMUNITS_AUTO) ? ucc_div_round_up(count, pp->n_frags) : ucc_max(1, pp->frag_size / dts);
    if (fcount < N) fcount = N; /* every rank of a fragment's exchange gets a part */
    n_total = (int)ucc_div_round_up(count, fcount);
    if (n_total < 2 || (count % fcount && count % fcount < N)) return UCC_ERR_NOT_FOUND;
    depth = (int)ucc_min(ucc_min(pp->pdepth, (unsigned)n_total), UCC_SCHEDULE_PIPELINED_MAX_FRAGS);
    sp = (shm_pipe_t *)calloc(1, sizeof(*sp));
    if (!sp) return UCC_ERR_NO_MEMORY;
    sp->fn = fn; sp->total = count; sp->dts = dts; sp->fcount = fcount;
    st = ucc_schedule_pipelined_init(b, team, shm_pipe_frag_init, shm_pipe_frag_setup, depth, n_total, pp->order, &sp->super);
    if (st != UCC_OK) { free(sp); return st; }
    sp->super.super.super.finalize = shm_pipe_finalize;
    tl_debug(team->context->lib, "%s pipelined: %d fragments of %zu elements, %d in flight", ucc_coll_type_str(a->coll_type), n_total, fcount, depth);
    *task_p = &sp->super.super.super;
    return UCC_OK;
}
static ucc_status_t shm_coll_init_alg(ucc_base_coll_args_t *bargs, ucc_base_team_t *team, ucc_coll_task_t **task_p, ucc_tl_shm_alg_fn_t fn)
{
    if (fn == ucc_tl_shm_allreduce_sra || fn == ucc_tl_shm_reduce_srg) {
        ucc_status_t st = shm_coll_init_pipelined(bargs, team, task_p, fn);
        if (st != UCC_ERR_NOT_FOUND) return st;
    }
    return shm_coll_init_plain(bargs, team, task_p, fn);
}

/* one tiny init wrapper per (coll, alg) so that the score map can point at distinct functions */
#define MAX_ALGS 8
#define DEFINE_INIT(_c, _i) \
    static ucc_status_t shm_init_##_c##_##_i(ucc_base_coll_args_t *b, ucc_base_team_t *tm, ucc_coll_task_t **tk) \
    { return shm_coll_init_alg(b, tm, tk, shm_algs[_c][_i].fn); }
#define DEFINE_INITS(_c) DEFINE_INIT(_c, 0) DEFINE_INIT(_c, 1) DEFINE_INIT(_c, 2) DEFINE_INIT(_c, 3) DEFINE_INIT(_c, 4) DEFINE_INIT(_c, 5) DEFINE_INIT(_c, 6)
DEFINE_INITS(0) DEFINE_INITS(1) DEFINE_INITS(2) DEFINE_INITS(3) DEFINE_INITS(4) DEFINE_INITS(5) DEFINE_INITS(6) DEFINE_INITS(7)
DEFINE_INITS(8) DEFINE_INITS(9) DEFINE_INITS(10) DEFINE_INITS(11) DEFINE_INITS(12) DEFINE_INITS(13) DEFINE_INITS(14) DEFINE_INITS(15)
#define INIT_ROW(_c) {shm_init_##_c##_0, shm_init_##_c##_1, shm_init_##_c##_2, shm_init_##_c##_3, shm_init_##_c##_4, shm_init_##_c##_5, shm_init_##_c##_6}
static const ucc_base_coll_init_fn_t shm_init_fns[UCC_COLL_TYPE_NUM][7] = {
    INIT_ROW(0), INIT_ROW(1), INIT_ROW(2), INIT_ROW(3), INIT_ROW(4), INIT_ROW(5), INIT_ROW(6), INIT_ROW(7),
    INIT_ROW(8), INIT_ROW(9), INIT_ROW(10), INIT_ROW(11), INIT_ROW(12), INIT_ROW(13), INIT_ROW(14), INIT_ROW(15)};

static unsigned n_algs(int c) { unsigned n = 0; while (shm_algs[c][n].name) n++; return n; }

static ucc_status_t shm_alg_id_to_init(int alg_id, const char *alg_id_str, ucc_coll_type_t coll_type, ucc_memory_type_t mem_type, ucc_base_coll_init_fn_t *init)
{
    int c = ucc_coll_type_index(coll_type);
    (void)mem_type;
    if (alg_id_str) { alg_id = -1; for (unsigned i = 0; i < n_algs(c); i++) if (!strcasecmp(alg_id_str, shm_algs[c][i].name)) alg_id = (int)i; }
    if (alg_id < 0 || (unsigned)alg_id >= n_algs(c)) return alg_id_str ? UCC_ERR_NOT_SUPPORTED : UCC_ERR_INVALID_PARAM;
    *init = shm_init_fns[c][alg_id];
    return UCC_OK;
}

/* default entry: algorithm 0 of the collective (score strings refine it) */
static ucc_status_t shm_coll_init(ucc_base_coll_args_t *b, ucc_base_team_t *team, ucc_coll_task_t **task)
{ return shm_init_fns[ucc_coll_type_index(b->args.coll_type)][0](b, team, task); }

static const char *shm_init_name(ucc_base_coll_init_fn_t init, ucc_base_team_t *team)
{
    (void)team;
    for (int c = 0; c < UCC_COLL_TYPE_NUM; c++) for (unsigned i = 0; i < n_algs(c); i++) if (shm_init_fns[c][i] == init) return shm_algs[c][i].name;
    return NULL;
}

/* ------------------------------------------------------------------ */
/* lib / context / team                                                */
/* ------------------------------------------------------------------ */
static ucc_status_t shm_lib_init(const ucc_base_lib_params_t *p, const ucc_base_lib_config_t *config, ucc_base_lib_t **lib_p)
{
    ucc_tl_shm_lib_t *lib = (ucc_tl_shm_lib_t *)calloc(1, sizeof(*lib));
    (void)p;
    if (!lib) return UCC_ERR_NO_MEMORY;
    ucc_tl_lib_init_base(&lib->super, &ucc_tl_shm, ucc_derived_of(config, ucc_tl_lib_config_t));
    /* algorithm plugins (reference tl/ucc_tl.h:64-69, tl/ucp/coll_plugins/example): modules libucc_tlcp_shm_<name>.so
     * exporting `ucc_tlcp_shm_<name>`; each contributes a score table that is merged into the TL's */
    if (!ucc_tl_shm.coll_plugins.framework_name) {
        if (ucc_components_load("tlcp_shm", &ucc_tl_shm.coll_plugins) != UCC_OK) ucc_tl_shm.coll_plugins.n_components = 0;
        for (int i = 0; i < ucc_tl_shm.coll_plugins.n_components; i++) /* so that ucc_info -c and the unused-variable check know them */
            ucc_config_table_register(&ucc_derived_of(ucc_tl_shm.coll_plugins.components[i], ucc_tl_coll_plugin_iface_t)->config);
    }
    *lib_p = &lib->super.super;
    return UCC_OK;
}
static void shm_lib_finalize(ucc_base_lib_t *lib) { free(lib); }
static ucc_status_t shm_lib_get_attr(const ucc_base_lib_t *lib, ucc_base_lib_attr_t *attr)
{ (void)lib; attr->attr.thread_mode = UCC_THREAD_MULTIPLE; attr->attr.coll_types = UCC_TL_SHM_SUPPORTED_COLLS; attr->flags = 0; attr->min_team_size = 2; attr->max_team_size = UCC_RANK_MAX; return UCC_OK; }
static ucc_status_t shm_lib_get_properties(ucc_base_lib_properties_t *p) { p->default_team_size = 2; p->min_team_size = 2; p->max_team_size = UCC_RANK_MAX; return UCC_OK; }

static ucc_status_t shm_ctx_create(const ucc_base_context_params_t *p, const ucc_base_ctx_config_t *config, ucc_base_context_t **ctx_p)
{
    ucc_tl_shm_context_t *ctx = (ucc_tl_shm_context_t *)calloc(1, sizeof(*ctx));
    ucc_status_t st;
    if (!ctx) return UCC_ERR_NO_MEMORY;
    ctx->super.super.ucc_context = p->context; ctx->super.super.lib = config->lib; ctx->tm = p->thread_mode;
    st = ucc_config_parser_clone_opts(config, &ctx->cfg, tl_shm_context_config_table);
    if (st != UCC_OK) { free(ctx); return st; }
    st = ucc_tl_shm_ctx_p2p_init(ctx);
    if (st != UCC_OK) { ucc_config_parser_release_opts(&ctx->cfg, tl_shm_context_config_table); free(ctx); return st; }
    ucc_mpool_init(&ctx->task_mp, 0, sizeof(ucc_tl_shm_task_t), 0, 64, 16, (unsigned)-1, NULL, p->thread_mode, "tl_shm_tasks");
    ucc_context_progress_register(p->context, ucc_tl_shm_progress, ctx);
    ucc_coll_score_set_name_resolver(shm_init_name);
    tl_debug(config->lib, "initialized tl context: %p, segment %s", (void *)ctx, ctx->addr.name);
    *ctx_p = &ctx->super.super;
    return UCC_OK;
}
static void shm_ctx_destroy(ucc_base_context_t *b)
{
    ucc_tl_shm_context_t *ctx = ucc_derived_of(b, ucc_tl_shm_context_t);
    ucc_context_progress_deregister(b->ucc_context, ucc_tl_shm_progress, ctx);
    ucc_tl_shm_ctx_p2p_cleanup(ctx);
    ucc_mpool_cleanup(&ctx->task_mp, 1);
    ucc_config_parser_release_opts(&ctx->cfg, tl_shm_context_config_table);
    free(ctx);
}
static ucc_status_t shm_ctx_get_attr(const ucc_base_context_t *b, ucc_base_ctx_attr_t *attr)
{
    const ucc_tl_shm_context_t *ctx = ucc_derived_of(b, ucc_tl_shm_context_t);
    if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR_LEN) attr->attr.ctx_addr_len = sizeof(ucc_tl_shm_addr_t);
    if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_CTX_ADDR) memcpy(attr->attr.ctx_addr, &ctx->addr, sizeof(ctx->addr));
    if (attr->attr.mask & UCC_CONTEXT_ATTR_FIELD_WORK_BUFFER_SIZE) attr->attr.global_work_buffer_size = 0;
    attr->topo_required = 0;
    return UCC_OK;
}

static ucc_status_t shm_team_create_post(ucc_base_context_t *b_ctx, const ucc_base_team_params_t *params, ucc_base_team_t **team_p)
{
    ucc_tl_shm_context_t *ctx = ucc_derived_of(b_ctx, ucc_tl_shm_context_t);
    ucc_tl_shm_team_t *team;
    uint64_t memb = 0;
    if (params->size < 2) return UCC_ERR_NOT_SUPPORTED;
    team = (ucc_tl_shm_team_t *)calloc(1, sizeof(*team));
    if (!team) return UCC_ERR_NO_MEMORY;
    team->super.super.context = b_ctx; team->super.super.params = *params;
    team->eps = (ucc_tl_shm_ep_t **)calloc(params->size, sizeof(void *));
    if (!team->eps) { free(team); return UCC_ERR_NO_MEMORY; }
    /* every member must be on this host and run tl/shm */
    for (ucc_rank_t r = 0; r < params->size; r++) {
        ucc_tl_shm_ep_t *ep;
        ucc_status_t st = ucc_tl_shm_get_ep(team, r, &ep);
        if (st != UCC_OK || ep->addr.host_hash != ctx->addr.host_hash) {
            tl_debug(b_ctx->lib, "rank %u is not reachable through shared memory", r);
            free(team->eps); free(team);
            return UCC_ERR_NOT_SUPPORTED;
        }
        memb = ucc_hash_mix(memb ^ ep->addr.ep_id);
    }
    /* RANKS_REORDERING (reference tl_ucp allgather_ring.c:120-124, reduce_scatter_ring.c:378-381: UCC_SBGP_FULL_HOST_ORDERED): when the team
     * IS the core team (not a sub-group of a hierarchy) and its members are interleaved over hosts / sockets, rings follow the host order */
    if (ctx->cfg.use_topo != UCC_NO && params->team && params->team->topo && params->size == params->team->size && ucc_ep_map_is_identity(&params->map)) {
        ucc_sbgp_t *ho = ctx->cfg.ranks_reordering ? ucc_topo_get_sbgp(params->team->topo, UCC_SBGP_FULL_HOST_ORDERED) : NULL;
        int reordered = 0;
        if (ho && ho->status == UCC_SBGP_ENABLED && ho->rank_map && ho->group_size == params->size)
            for (ucc_rank_t r = 0; r < params->size; r++) if (ho->rank_map[r] != r) reordered = 1;
        if (reordered) {
            team->ring_order = (ucc_rank_t *)malloc(2 * sizeof(ucc_rank_t) * params->size);
            if (team->ring_order) {
                char buf[256]; size_t o = 0;
                team->ring_pos = team->ring_order + params->size;
                for (ucc_rank_t r = 0; r < params->size; r++) { team->ring_order[r] = ho->rank_map[r]; team->ring_pos[ho->rank_map[r]] = r; }
                for (ucc_rank_t r = 0; r < params->size && o + 12 < sizeof(buf); r++) o += (size_t)snprintf(buf + o, sizeof(buf) - o, " %u", team->ring_order[r]);
                tl_debug(b_ctx->lib, "team %u: ring order by host:%s", (unsigned)params->id, buf);
            }
        }
    } else if (ctx->cfg.use_topo == UCC_YES && !(params->team && params->team->topo)) {
        tl_debug(b_ctx->lib, "USE_TOPO=y and the team has no topology");
        free(team->eps); free(team);
        return UCC_ERR_NOT_SUPPORTED;
    }
    team->tag_base = ((uint64_t)params->id << 48) | ((uint64_t)(params->scope & 0xf) << 44) | ((uint64_t)(params->scope_id & 0xf) << 40) | ((memb & 0xff) << 32);
    *team_p = &team->super.super;
    return UCC_OK;
}
static ucc_status_t shm_team_create_test(ucc_base_team_t *t) { (void)t; return UCC_OK; }
static ucc_status_t shm_team_destroy(ucc_base_team_t *b) { ucc_tl_shm_team_t *team = ucc_derived_of(b, ucc_tl_shm_team_t); free(team->eps); free(team->svc); free(team->ring_order); free(team); return UCC_OK; }

static ucc_status_t shm_team_get_scores(ucc_base_team_t *b_team, ucc_coll_score_t **score_p)
{
    ucc_tl_shm_team_t *team = ucc_derived_of(b_team, ucc_tl_shm_team_t);
    ucc_tl_shm_context_t *ctx = SHM_CTX(team);
    ucc_memory_type_t mt[3] = {UCC_MEMORY_TYPE_HOST, UCC_MEMORY_TYPE_CUDA, UCC_MEMORY_TYPE_CUDA_MANAGED};
    int n_mt = 1 + (ucc_mc_available(UCC_MEMORY_TYPE_CUDA) == UCC_OK ? 2 : 0);
    ucc_coll_score_team_info_t info = {UCC_TL_SHM_DEFAULT_SCORE, UCC_TL_TEAM_SIZE(team), UCC_TL_SHM_SUPPORTED_COLLS, mt, n_mt, shm_coll_init, shm_alg_id_to_init};
    ucc_coll_score_t *score;
    ucc_status_t st = ucc_coll_score_build_default(b_team, UCC_TL_SHM_DEFAULT_SCORE, shm_coll_init, UCC_TL_SHM_SUPPORTED_COLLS, mt, n_mt, &score);
    if (st != UCC_OK) return st;
    st = ucc_tl_apply_tune(&team->super, score, &info, SHM_DEFAULT_SELECT_STR, ctx->cfg.super.super.score_str);
    if (st != UCC_OK) { ucc_coll_score_free(score); return st; }
    for (int i = 0; i < ucc_tl_shm.coll_plugins.n_components; i++) {
        ucc_tl_coll_plugin_iface_t *p = ucc_derived_of(ucc_tl_shm.coll_plugins.components[i], ucc_tl_coll_plugin_iface_t);
        ucc_coll_score_t *ps = NULL;
        if (!p->get_scores || p->get_scores(b_team, &ps) != UCC_OK || !ps) continue;
        st = ucc_coll_score_merge_in(&score, ps); /* higher score wins per range, the TL's own entry stays as the fallback */
        if (st != UCC_OK) return st;
    }
    *score_p = score;
    return UCC_OK;
}

/* ------------------------------------------------------------------ */
/* service collectives                                                 */
/* ------------------------------------------------------------------ */
static ucc_status_t svc_task(ucc_base_team_t *b_team, ucc_subset_t subset, ucc_coll_type_t ct, ucc_tl_shm_task_t **tp)
{
    ucc_base_coll_args_t ba;
    ucc_tl_shm_task_t *t;
    ucc_status_t st;
    memset(&ba, 0, sizeof(ba));
    ba.args.coll_type = ct; ba.team = b_team->params.team;
    st = ucc_tl_shm_task_alloc(&ba, b_team, &t);
    if (st != UCC_OK) return st;
    t->vmap = subset.map; t->vsize = (ucc_rank_t)subset.map.ep_num; t->vrank = subset.myrank;
    /* service collectives on sub-sets must not disturb the team-wide sequence: tag them apart */
    t->team->seq_num--;
    {
        /* identity of the subset + how many service collectives ran on exactly this subset */
        uint64_t h = 0x9e3779b97f4a7c15ull; unsigned i;
        ucc_tl_shm_team_t *tm = t->team;
        for (ucc_rank_t r = 0; r < t->vsize; r++) h = ucc_hash_mix(h ^ ucc_ep_map_eval(subset.map, r));
        for (i = 0; i < tm->n_svc; i++) if (tm->svc[i].hash == h) break;
        if (i == tm->n_svc) {
            tm->svc = (struct shm_svc_cnt *)realloc(tm->svc, (tm->n_svc + 1) * sizeof(*tm->svc));
            tm->svc[i].hash = h; tm->svc[i].cnt = 0; tm->n_svc++;
        }
        t->coll_seq = 0x8000u | (uint32_t)((h & 0xff) << 6) | (tm->svc[i].cnt++ & 0x3f);
    }
    t->super.flags |= UCC_COLL_TASK_FLAG_INTERNAL;
    *tp = t;
    return UCC_OK;
}
static ucc_status_t shm_service_allreduce(ucc_base_team_t *team, void *sbuf, void *rbuf, ucc_datatype_t dt, size_t count,
                                          ucc_reduction_op_t op, ucc_subset_t subset, ucc_coll_task_t **task_p)
{
    ucc_tl_shm_task_t *t; ucc_status_t st = svc_task(team, subset, UCC_COLL_TYPE_ALLREDUCE, &t);
    if (st != UCC_OK) return st;
    t->dt = dt; t->op = op; t->alpha = 1.0 / (double)t->vsize;
    st = ucc_tl_shm_service_allreduce_prog(t, sbuf, rbuf, count);
    if (st != UCC_OK) { ucc_tl_shm_task_finalize(&t->super); return st; }
    *task_p = &t->super;
    return UCC_OK;
}
static ucc_status_t shm_service_allgather(ucc_base_team_t *team, void *sbuf, void *rbuf, size_t msgsize, ucc_subset_t subset, ucc_coll_task_t **task_p)
{
    ucc_tl_shm_task_t *t; ucc_status_t st = svc_task(team, subset, UCC_COLL_TYPE_ALLGATHER, &t);
    if (st != UCC_OK) return st;
    st = ucc_tl_shm_service_allgather_prog(t, sbuf, rbuf, msgsize);
    if (st != UCC_OK) { ucc_tl_shm_task_finalize(&t->super); return st; }
    *task_p = &t->super;
    return UCC_OK;
}
static ucc_status_t shm_service_bcast(ucc_base_team_t *team, void *buf, size_t msgsize, ucc_rank_t root, ucc_subset_t subset, ucc_coll_task_t **task_p)
{
    ucc_tl_shm_task_t *t; ucc_status_t st = svc_task(team, subset, UCC_COLL_TYPE_BCAST, &t);
    if (st != UCC_OK) return st;
    st = ucc_tl_shm_bcast_knomial_prog(t, buf, msgsize, UCC_MEMORY_TYPE_HOST, root, 4);
    if (st != UCC_OK) { ucc_tl_shm_task_finalize(&t->super); return st; }
    *task_p = &t->super;
    return UCC_OK;
}
static void shm_service_update_id(ucc_base_team_t *b, uint16_t id)
{
    ucc_tl_shm_team_t *team = ucc_derived_of(b, ucc_tl_shm_team_t);
    team->tag_base = (team->tag_base & ~(0xffffull << 48)) | ((uint64_t)id << 48);
    b->params.id = id;
}

ucc_tl_iface_t ucc_tl_shm = {
    .super = {.name = "shm", .score = UCC_TL_SHM_DEFAULT_SCORE},
    .tl_lib_config = {"TL_SHM lib", "TL_SHM_", tl_shm_lib_config_table, sizeof(ucc_tl_shm_lib_config_t), {NULL, NULL}},
    .tl_context_config = {"TL_SHM context", "TL_SHM_", tl_shm_context_config_table, sizeof(ucc_tl_shm_context_config_t), {NULL, NULL}},
    .lib = {shm_lib_init, shm_lib_finalize, shm_lib_get_attr, shm_lib_get_properties},
    .context = {shm_ctx_create, NULL, shm_ctx_destroy, shm_ctx_get_attr, ucc_tl_shm_mem_map, ucc_tl_shm_mem_unmap, ucc_tl_shm_memh_pack},
    .team = {shm_team_create_post, shm_team_create_test, shm_team_destroy, shm_team_get_scores},
    .coll = {shm_coll_init},
    .scoll = {shm_service_allreduce, shm_service_allgather, shm_service_bcast, shm_service_update_id},
};
UCC_COMPONENT_REGISTER_STATIC(tl, ucc_tl_shm)
static void UCC_CTOR tl_shm_register(void)
{
    ucc_config_table_register(&ucc_tl_shm.tl_lib_config); ucc_config_table_register(&ucc_tl_shm.tl_context_config);
    for (int c = 0; c < UCC_COLL_TYPE_NUM; c++) {
        unsigned i;
        for (i = 0; i < n_algs(c) && i < 7; i++) { shm_alg_info[c][i].id = i; shm_alg_info[c][i].name = shm_algs[c][i].name; shm_alg_info[c][i].desc = shm_algs[c][i].desc; }
        shm_alg_info[c][i].name = NULL;
        ucc_tl_shm.alg_info[c] = shm_alg_info[c];
    }
}
