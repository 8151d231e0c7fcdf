#ifndef UCC_EC_CPU_H_
#define UCC_EC_CPU_H_
#include "components/ec/base/ucc_ec_base.h"
#include "utils/ucc_mpool.h"
typedef struct ucc_ec_cpu_config { ucc_ec_config_t super; unsigned exec_max_tasks; } ucc_ec_cpu_config_t;
typedef struct ucc_ec_cpu { ucc_ec_base_t super; ucc_thread_mode_t thread_mode; ucc_mpool_t executors, executor_tasks; } ucc_ec_cpu_t;
extern ucc_ec_cpu_t ucc_ec_cpu;
ucc_status_t ucc_ec_cpu_reduce(void *dst, void *const *srcs, unsigned n_srcs, size_t count, ucc_datatype_t dt,
                               ucc_reduction_op_t op, int with_alpha, double alpha);
#endif
