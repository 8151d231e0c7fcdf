#include "ucc_mc.h"
#include "core/ucc_global_opts.h"

static ucc_mc_base_t *mc_ops[UCC_MEMORY_TYPE_LAST];

ucc_config_field_t ucc_mc_config_table[] = {
    {"LOG_LEVEL", "warn", "UCC logging level of the memory component.",
     ucc_offsetof(ucc_mc_config_t, log_component.log_level), UCC_CONFIG_TYPE_ENUM(ucc_log_level_cfg_names)},
    {NULL}};

#define CHECK_MT(_mt)                                                                          \
    do { if ((_mt) >= UCC_MEMORY_TYPE_LAST || !mc_ops[_mt]) {                                  \
        ucc_debug("no memory component for memory type %d", (int)(_mt)); return UCC_ERR_NOT_SUPPORTED; } } while (0)

ucc_status_t ucc_mc_init(const ucc_mc_params_t *mc_params)
{
    ucc_component_framework_t *fw = &ucc_global_config.mc_framework;
    for (int i = 0; i < fw->n_components; i++) {
        ucc_mc_base_t *mc = ucc_derived_of(fw->components[i], ucc_mc_base_t);
        if (mc->ref_cnt == 0) {
            ucc_status_t st;
            mc->config = (ucc_mc_config_t *)calloc(1, mc->config_table.size);
            if (!mc->config) return UCC_ERR_NO_MEMORY;
            st = ucc_config_parser_fill_opts(mc->config, &mc->config_table, "UCC_", 1);
            if (st != UCC_OK) { free(mc->config); mc->config = NULL; continue; }
            snprintf(mc->config->log_component.name, sizeof(mc->config->log_component.name), "MC_%s", mc->super.name);
            st = mc->init(mc_params);
            if (st != UCC_OK) {
                ucc_debug("mc %s is not available: %s", mc->super.name, ucc_status_string(st));
                ucc_config_parser_release_opts(mc->config, mc->config_table.table);
                free(mc->config); mc->config = NULL;
                continue;
            }
        }
        mc->ref_cnt++;
        mc_ops[mc->type] = mc;
        if (mc->type == UCC_MEMORY_TYPE_CUDA && !mc_ops[UCC_MEMORY_TYPE_CUDA_MANAGED]) mc_ops[UCC_MEMORY_TYPE_CUDA_MANAGED] = mc;
    }
    return UCC_OK;
}
ucc_status_t ucc_mc_available(ucc_memory_type_t mt) { CHECK_MT(mt); return UCC_OK; }
ucc_status_t ucc_mc_get_attr(ucc_mc_attr_t *attr, ucc_memory_type_t mt) { CHECK_MT(mt); return mc_ops[mt]->get_attr(attr); }

ucc_status_t ucc_mc_get_mem_attr(const void *ptr, ucc_mem_attr_t *mem_attr)
{
    /* device components first: the cpu component accepts everything */
    for (int mt = UCC_MEMORY_TYPE_HOST + 1; mt < UCC_MEMORY_TYPE_LAST; mt++) {
        if (!mc_ops[mt] || (mt == UCC_MEMORY_TYPE_CUDA_MANAGED && mc_ops[mt] == mc_ops[UCC_MEMORY_TYPE_CUDA])) continue;
        if (mc_ops[mt]->ops.mem_query(ptr, mem_attr) == UCC_OK) return UCC_OK;
    }
    if (mc_ops[UCC_MEMORY_TYPE_HOST]) return mc_ops[UCC_MEMORY_TYPE_HOST]->ops.mem_query(ptr, mem_attr);
    mem_attr->mem_type = UCC_MEMORY_TYPE_HOST; mem_attr->base_address = (void *)ptr; mem_attr->alloc_length = 1;
    return UCC_OK;
}
ucc_status_t ucc_mc_alloc(ucc_mc_buffer_header_t **h, size_t len, ucc_memory_type_t mt) { CHECK_MT(mt); return mc_ops[mt]->ops.mem_alloc(h, len, mt); }
ucc_status_t ucc_mc_free(ucc_mc_buffer_header_t *h) { ucc_memory_type_t mt = h->mt; CHECK_MT(mt); return mc_ops[mt]->ops.mem_free(h); }
ucc_status_t ucc_mc_flush(ucc_memory_type_t mt) { CHECK_MT(mt); return mc_ops[mt]->ops.flush ? mc_ops[mt]->ops.flush() : UCC_OK; }
ucc_status_t ucc_mc_memcpy(void *dst, const void *src, size_t len, ucc_memory_type_t dst_mem, ucc_memory_type_t src_mem)
{
    ucc_memory_type_t mt;
    if (dst_mem == UCC_MEMORY_TYPE_UNKNOWN || src_mem == UCC_MEMORY_TYPE_UNKNOWN) return UCC_ERR_INVALID_PARAM;
    if (src_mem == UCC_MEMORY_TYPE_HOST && dst_mem == UCC_MEMORY_TYPE_HOST) { CHECK_MT(UCC_MEMORY_TYPE_HOST); return mc_ops[UCC_MEMORY_TYPE_HOST]->ops.memcpy(dst, src, len, dst_mem, src_mem); }
    mt = dst_mem == UCC_MEMORY_TYPE_HOST ? src_mem : dst_mem;
    CHECK_MT(mt);
    return mc_ops[mt]->ops.memcpy(dst, src, len, dst_mem, src_mem);
}
ucc_status_t ucc_mc_memset(void *ptr, int value, size_t size, ucc_memory_type_t mt) { CHECK_MT(mt); return mc_ops[mt]->ops.memset(ptr, value, size); }
ucc_status_t ucc_mc_finalize(void)
{
    /* CUDA_MANAGED is served by the cuda component (an alias entry, not a reference of its own): it must neither be counted
     * twice nor dropped while another lib instance still holds the component.  (Counting it after the CUDA entry had been
     * cleared drove ref_cnt below zero, so the NEXT ucc_init in the process skipped the component's init and used it without a
     * config - found by tools/ucc_test_dist.py -M cuda, which creates one lib per team kind.) */
    ucc_mc_base_t *cuda = mc_ops[UCC_MEMORY_TYPE_CUDA];
    int managed_alias = cuda && mc_ops[UCC_MEMORY_TYPE_CUDA_MANAGED] == cuda;
    for (int mt = 0; mt < UCC_MEMORY_TYPE_LAST; mt++) {
        ucc_mc_base_t *mc = mc_ops[mt];
        if (!mc || (mt == UCC_MEMORY_TYPE_CUDA_MANAGED && managed_alias)) continue;
        if (--mc->ref_cnt == 0) {
            mc->finalize();
            ucc_config_parser_release_opts(mc->config, mc->config_table.table);
            free(mc->config); mc->config = NULL; mc_ops[mt] = NULL;
        }
    }
    if (managed_alias && !mc_ops[UCC_MEMORY_TYPE_CUDA]) mc_ops[UCC_MEMORY_TYPE_CUDA_MANAGED] = NULL;
    return UCC_OK;
}
