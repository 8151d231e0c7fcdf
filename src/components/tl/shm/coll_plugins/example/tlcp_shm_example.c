/* Example algorithm plugin for tl/shm (role of reference tl/ucp/coll_plugins/example): a "linear" allreduce - everybody
 * sends to rank 0, rank 0 reduces and sends the result back - built with the TL's public step-program API.
 * Disabled unless UCC_TLCP_SHM_EXAMPLE_SCORE is set to a non-zero score; with a score above the TL's own (10) the
 * plugin's init function wins the allreduce ranges and the built-in algorithms stay behind it as fallbacks. */
#include "components/tl/shm/tl_shm.h"
#include "components/tl/shm/tl_shm_prog.h"
#include "coll_score/ucc_coll_score.h"

typedef struct tlcp_example_config { unsigned score; } tlcp_example_config_t;
static ucc_config_field_t tlcp_example_table[] = {
    {"SCORE", "0", "Score of the example allreduce (0: plugin disabled)", ucc_offsetof(tlcp_example_config_t, score), UCC_CONFIG_TYPE_UINT},
    {NULL}};
int ucc_tlcp_shm_example_calls = 0; /* how many collectives the plugin instantiated (read by the tests) */

static ucc_status_t example_build(ucc_tl_shm_task_t *t)
{
    ucc_coll_args_t *a = &t->super.bargs.args;
    ucc_rank_t N = t->vsize, r = t->vrank;
    size_t count = a->dst.info.count, len = count * ucc_dt_size(a->dst.info.datatype);
    ucc_memory_type_t mt = a->dst.info.mem_type;
    void *src = UCC_IS_INPLACE(*a) ? a->dst.info.buffer : a->src.info.buffer, *dst = a->dst.info.buffer, *scratch;
    ucc_status_t st;
    if (r == 0) {
        if ((st = shm_task_scratch(t, len, mt, &scratch)) != UCC_OK) return st;
        if ((st = shm_prog_copy(t, dst, src, len, mt, mt)) != UCC_OK) return st;
        for (ucc_rank_t p = 1; p < N; p++) {
            if ((st = shm_prog_recv(t, p, scratch, len, mt, 1)) != UCC_OK || (st = shm_prog_wait(t)) != UCC_OK) return st;
            if ((st = shm_prog_reduce(t, dst, dst, scratch, count, mt, 0)) != UCC_OK) return st;
        }
        for (ucc_rank_t p = 1; p < N; p++) if ((st = shm_prog_send(t, p, dst, len, mt, 2)) != UCC_OK) return st;
    } else {
        if ((st = shm_prog_send(t, 0, src, len, mt, 1)) != UCC_OK || (st = shm_prog_recv(t, 0, dst, len, mt, 2)) != UCC_OK) return st;
    }
    return shm_prog_wait(t);
}

static ucc_status_t example_allreduce_init(ucc_base_coll_args_t *b, ucc_base_team_t *team, ucc_coll_task_t **task_p)
{
    ucc_tl_shm_task_t *t;
    ucc_status_t st;
    if (b->args.op == UCC_OP_AVG || UCC_COLL_ARGS_ACTIVE_SET(&b->args)) return UCC_ERR_NOT_SUPPORTED; /* the built-in algorithms take it */
    st = ucc_tl_shm_task_alloc(b, team, &t);
    if (st != UCC_OK) return st;
    t->dt = b->args.dst.info.datatype; t->op = b->args.op; t->build = example_build;
    st = example_build(t);
    if (st != UCC_OK) { t->team->seq_num--; ucc_tl_shm_task_finalize(&t->super); return st; }
    ucc_tlcp_shm_example_calls++;
    *task_p = &t->super;
    return UCC_OK;
}

static ucc_status_t example_get_scores(ucc_base_team_t *team, ucc_coll_score_t **score_p);
ucc_tl_coll_plugin_iface_t ucc_tlcp_shm_example = {
    .super = {.name = "example", .score = 0},
    .default_score = 0,
    .config = {"TLCP_SHM_EXAMPLE", "TLCP_SHM_EXAMPLE_", tlcp_example_table, sizeof(tlcp_example_config_t), {NULL, NULL}},
    .get_scores = example_get_scores,
};

static ucc_status_t example_get_scores(ucc_base_team_t *team, ucc_coll_score_t **score_p)
{
    tlcp_example_config_t cfg;
    ucc_coll_score_t *score;
    ucc_status_t st = ucc_config_parser_fill_opts(&cfg, &ucc_tlcp_shm_example.config, "UCC_", 1);
    if (st != UCC_OK || cfg.score == 0) return UCC_ERR_NOT_SUPPORTED;
    if ((st = ucc_coll_score_alloc(&score)) != UCC_OK) return st;
    st = ucc_coll_score_add_range(score, UCC_COLL_TYPE_ALLREDUCE, UCC_MEMORY_TYPE_HOST, 0, UCC_MSG_MAX, cfg.score, example_allreduce_init, team);
    if (st != UCC_OK) { ucc_coll_score_free(score); return st; }
    *score_p = score;
    return UCC_OK;
}
