/* ec/cpu: synchronous host executor — a posted task is executed inline and
 * completes immediately (reference ec/cpu/ec_cpu.c:68-200 has the same model). */
#include "ec_cpu.h"
#include "core/ucc_dt.h"

static ucc_config_field_t ucc_ec_cpu_config_table[] = {
    {"", "", NULL, ucc_offsetof(ucc_ec_cpu_config_t, super), UCC_CONFIG_TYPE_TABLE(ucc_ec_config_table)},
    {"EXEC_MAX_TASKS", "inf", "Maximum number of outstanding tasks per executor", ucc_offsetof(ucc_ec_cpu_config_t, exec_max_tasks), UCC_CONFIG_TYPE_UINT},
    {NULL}};

static ucc_status_t ec_cpu_init(const ucc_ec_params_t *p)
{
    ucc_ec_cpu.thread_mode = p->thread_mode;
    ucc_mpool_init(&ucc_ec_cpu.executors, 0, sizeof(ucc_ee_executor_t), 0, 64, 16, (unsigned)-1, NULL, p->thread_mode, "ec cpu executors");
    ucc_mpool_init(&ucc_ec_cpu.executor_tasks, 0, sizeof(ucc_ee_executor_task_t), 0, 64, 16, (unsigned)-1, NULL, p->thread_mode, "ec cpu tasks");
    return UCC_OK;
}
static ucc_status_t ec_cpu_get_attr(ucc_ec_attr_t *a) { (void)a; return UCC_OK; }
static ucc_status_t ec_cpu_finalize(void) { ucc_mpool_cleanup(&ucc_ec_cpu.executors, 1); ucc_mpool_cleanup(&ucc_ec_cpu.executor_tasks, 1); return UCC_OK; }
static ucc_status_t ev_create(void **e) { *e = NULL; return UCC_OK; }
static ucc_status_t ev_destroy(void *e) { (void)e; return UCC_OK; }
static ucc_status_t ev_post(void *c, void *e) { (void)c; (void)e; return UCC_OK; }
static ucc_status_t ev_test(void *e) { (void)e; return UCC_OK; }

static ucc_status_t exec_init(const ucc_ee_executor_params_t *p, ucc_ee_executor_t **x)
{
    ucc_ee_executor_t *e = (ucc_ee_executor_t *)ucc_mpool_get(&ucc_ec_cpu.executors);
    (void)p;
    if (!e) return UCC_ERR_NO_MEMORY;
    e->ee_type = UCC_EE_CPU_THREAD; e->ee_context = NULL; *x = e;
    return UCC_OK;
}
static ucc_status_t exec_status(const ucc_ee_executor_t *x) { (void)x; return UCC_OK; }
static ucc_status_t exec_start(ucc_ee_executor_t *x, void *c) { x->ee_context = c; return UCC_OK; }
static ucc_status_t exec_stop(ucc_ee_executor_t *x) { (void)x; return UCC_OK; }
static ucc_status_t exec_finalize(ucc_ee_executor_t *x) { ucc_mpool_put(x); return UCC_OK; }

static ucc_status_t exec_task_post(ucc_ee_executor_t *x, const ucc_ee_executor_task_args_t *a, ucc_ee_executor_task_t **tp)
{
    ucc_ee_executor_task_t *t = (ucc_ee_executor_task_t *)ucc_mpool_get(&ucc_ec_cpu.executor_tasks);
    ucc_status_t st = UCC_OK;
    int with_alpha = !!(a->flags & UCC_EEE_TASK_FLAG_REDUCE_WITH_ALPHA);
    if (!t) return UCC_ERR_NO_MEMORY;
    t->eee = x; t->args = *a; t->handle = NULL;
    switch (a->task_type) {
    case UCC_EE_EXECUTOR_TASK_REDUCE: {
        void *const *srcs = (a->flags & UCC_EEE_TASK_FLAG_REDUCE_SRCS_EXT) ? (void *const *)a->reduce.srcs_ext : (void *const *)a->reduce.srcs;
        st = ucc_ec_cpu_reduce(a->reduce.dst, srcs, a->reduce.n_srcs, a->reduce.count, a->reduce.dt, a->reduce.op, with_alpha, a->reduce.alpha);
        break;
    }
    case UCC_EE_EXECUTOR_TASK_REDUCE_STRIDED: {
        unsigned n = (unsigned)a->reduce_strided.n_src2 + 1;
        void *small[16], **srcs = n <= 16 ? small : (void **)malloc(n * sizeof(void *));
        srcs[0] = a->reduce_strided.src1;
        for (unsigned k = 1; k < n; k++) srcs[k] = (char *)a->reduce_strided.src2 + (size_t)(k - 1) * a->reduce_strided.stride;
        st = ucc_ec_cpu_reduce(a->reduce_strided.dst, srcs, n, a->reduce_strided.count, a->reduce_strided.dt, a->reduce_strided.op,
                               with_alpha, a->reduce_strided.alpha);
        if (srcs != small) free(srcs);
        break;
    }
    case UCC_EE_EXECUTOR_TASK_REDUCE_MULTI_DST:
        for (unsigned j = 0; j < a->reduce_multi_dst.n_bufs && st == UCC_OK; j++) {
            void *srcs[2] = {a->reduce_multi_dst.src1[j], a->reduce_multi_dst.src2[j]};
            st = ucc_ec_cpu_reduce(a->reduce_multi_dst.dst[j], srcs, 2, a->reduce_multi_dst.counts[j], a->reduce_multi_dst.dt,
                                   a->reduce_multi_dst.op, 0, 0.0);
        }
        break;
    case UCC_EE_EXECUTOR_TASK_COPY: memcpy(a->copy.dst, a->copy.src, a->copy.len); break;
    case UCC_EE_EXECUTOR_TASK_COPY_MULTI:
        for (size_t j = 0; j < a->copy_multi.num_vectors; j++) memcpy(a->copy_multi.dst[j], a->copy_multi.src[j], a->copy_multi.counts[j]);
        break;
    default: st = UCC_ERR_INVALID_PARAM; break;
    }
    if (st != UCC_OK) { ucc_mpool_put(t); return st; }
    t->status = UCC_OK;
    *tp = t;
    return UCC_OK;
}
static ucc_status_t exec_task_test(const ucc_ee_executor_task_t *t) { return t->status; }
static ucc_status_t exec_task_finalize(ucc_ee_executor_task_t *t) { ucc_mpool_put(t); return UCC_OK; }

ucc_ec_cpu_t ucc_ec_cpu = {
    .super = {.super = {.name = "cpu", .score = 1}, .ref_cnt = 0, .type = UCC_EE_CPU_THREAD,
              .config_table = {"CPU execution component", "EC_CPU_", ucc_ec_cpu_config_table, sizeof(ucc_ec_cpu_config_t), {NULL, NULL}},
              .init = ec_cpu_init, .get_attr = ec_cpu_get_attr, .finalize = ec_cpu_finalize,
              .ops = {ev_create, ev_destroy, ev_post, ev_test},
              .executor_ops = {exec_init, exec_status, exec_start, exec_stop, exec_finalize, exec_task_post, exec_task_test, exec_task_finalize}}};
UCC_COMPONENT_REGISTER_STATIC(ec, ucc_ec_cpu)
static void UCC_CTOR ec_cpu_cfg_register(void) { ucc_config_table_register(&ucc_ec_cpu.super.config_table); }
