#include "ucc_ec.h"
#include "core/ucc_global_opts.h"

static ucc_ec_base_t *ec_ops[UCC_EE_LAST];

ucc_config_field_t ucc_ec_config_table[] = {
    {"LOG_LEVEL", "warn", "UCC logging level of the execution component.",
     ucc_offsetof(ucc_ec_config_t, log_component.log_level), UCC_CONFIG_TYPE_ENUM(ucc_log_level_cfg_names)},
    {NULL}};

#define CHECK_EE(_t) do { if ((_t) >= UCC_EE_LAST || !ec_ops[_t]) return UCC_ERR_NOT_SUPPORTED; } while (0)

ucc_status_t ucc_ec_init(const ucc_ec_params_t *ec_params)
{
    ucc_component_framework_t *fw = &ucc_global_config.ec_framework;
    for (int i = 0; i < fw->n_components; i++) {
        ucc_ec_base_t *ec = ucc_derived_of(fw->components[i], ucc_ec_base_t);
        if (ec->ref_cnt == 0) {
            ucc_status_t st;
            ec->config = (ucc_ec_config_t *)calloc(1, ec->config_table.size);
            if (!ec->config) return UCC_ERR_NO_MEMORY;
            st = ucc_config_parser_fill_opts(ec->config, &ec->config_table, "UCC_", 1);
            if (st != UCC_OK) { free(ec->config); ec->config = NULL; continue; }
            snprintf(ec->config->log_component.name, sizeof(ec->config->log_component.name), "EC_%s", ec->super.name);
            st = ec->init(ec_params);
            if (st != UCC_OK) {
                ucc_debug("ec %s is not available: %s", ec->super.name, ucc_status_string(st));
                ucc_config_parser_release_opts(ec->config, ec->config_table.table);
                free(ec->config); ec->config = NULL;
                continue;
            }
        }
        ec->ref_cnt++;
        ec_ops[ec->type] = ec;
    }
    return UCC_OK;
}
ucc_status_t ucc_ec_available(ucc_ee_type_t t) { CHECK_EE(t); return UCC_OK; }
ucc_status_t ucc_ec_get_attr(ucc_ec_attr_t *a, ucc_ee_type_t t) { CHECK_EE(t); return ec_ops[t]->get_attr(a); }
ucc_status_t ucc_ec_finalize(void)
{
    for (int t = 0; t < UCC_EE_LAST; t++) {
        ucc_ec_base_t *ec = ec_ops[t];
        if (!ec) continue;
        if (--ec->ref_cnt == 0) {
            ec->finalize();
            ucc_config_parser_release_opts(ec->config, ec->config_table.table);
            free(ec->config); ec->config = NULL; ec_ops[t] = NULL;
        }
    }
    return UCC_OK;
}
ucc_status_t ucc_ec_create_event(void **e, ucc_ee_type_t t) { CHECK_EE(t); return ec_ops[t]->ops.create_event(e); }
ucc_status_t ucc_ec_destroy_event(void *e, ucc_ee_type_t t) { CHECK_EE(t); return ec_ops[t]->ops.destroy_event(e); }
ucc_status_t ucc_ec_event_post(void *c, void *e, ucc_ee_type_t t) { CHECK_EE(t); return ec_ops[t]->ops.event_post(c, e); }
ucc_status_t ucc_ec_event_test(void *e, ucc_ee_type_t t) { CHECK_EE(t); return ec_ops[t]->ops.event_test(e); }
ucc_status_t ucc_ee_executor_init(const ucc_ee_executor_params_t *p, ucc_ee_executor_t **x) { CHECK_EE(p->ee_type); return ec_ops[p->ee_type]->executor_ops.init(p, x); }
ucc_status_t ucc_ee_executor_status(const ucc_ee_executor_t *x) { CHECK_EE(x->ee_type); return ec_ops[x->ee_type]->executor_ops.status(x); }
ucc_status_t ucc_ee_executor_start(ucc_ee_executor_t *x, void *c) { CHECK_EE(x->ee_type); return ec_ops[x->ee_type]->executor_ops.start(x, c); }
ucc_status_t ucc_ee_executor_stop(ucc_ee_executor_t *x) { CHECK_EE(x->ee_type); return ec_ops[x->ee_type]->executor_ops.stop(x); }
ucc_status_t ucc_ee_executor_finalize(ucc_ee_executor_t *x) { CHECK_EE(x->ee_type); return ec_ops[x->ee_type]->executor_ops.finalize(x); }
ucc_status_t ucc_ee_executor_task_post(ucc_ee_executor_t *x, const ucc_ee_executor_task_args_t *a, ucc_ee_executor_task_t **t)
{ CHECK_EE(x->ee_type); return ec_ops[x->ee_type]->executor_ops.task_post(x, a, t); }
ucc_status_t ucc_ee_executor_task_test(const ucc_ee_executor_task_t *t) { CHECK_EE(t->eee->ee_type); return ec_ops[t->eee->ee_type]->executor_ops.task_test(t); }
ucc_status_t ucc_ee_executor_task_finalize(ucc_ee_executor_task_t *t) { CHECK_EE(t->eee->ee_type); return ec_ops[t->eee->ee_type]->executor_ops.task_finalize(t); }
