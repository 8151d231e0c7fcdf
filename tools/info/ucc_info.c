/* ucc_info: version, build configuration, configuration dump, default scores, algorithm lists.
 * Flags follow the reference tool (tools/info/ucc_info.c:21-170): -v -b -c -a -f -s -A. */
#include <ucc/api/ucc.h>
#include "core/ucc_global_opts.h"
#include "components/cl/ucc_cl.h"
#include "components/tl/ucc_tl.h"
#include "utils/ucc_coll_utils.h"
#include <getopt.h>
#include <stdio.h>

ucc_status_t ucc_constructor(void);

static void usage(void)
{
    printf("Usage: ucc_info [options]\n"
           "  -v   Show version information\n  -b   Show build configuration\n  -c   Show UCC configuration\n"
           "  -a   Show also hidden configuration\n  -f   Display fully decorated output (documentation)\n"
           "  -s   Show default components scores\n  -A   Show collective algorithms available for selection with TUNE\n  -h   Show this help\n");
}
static void print_algs(const char *kind, const char *name, const ucc_base_coll_alg_info_t *const *alg_info)
{
    int any = 0;
    for (int c = 0; c < UCC_COLL_TYPE_NUM; c++) if (alg_info[c] && alg_info[c][0].name) any = 1;
    if (!any) return;
    printf("%s/%s algorithms:\n", kind, name);
    for (int c = 0; c < UCC_COLL_TYPE_NUM; c++) {
        if (!alg_info[c] || !alg_info[c][0].name) continue;
        printf("  %s\n", ucc_coll_type_str((ucc_coll_type_t)UCC_BIT(c)));
        for (int i = 0; alg_info[c][i].name; i++) printf("    %u : %16s : %s\n", alg_info[c][i].id, alg_info[c][i].name, alg_info[c][i].desc ? alg_info[c][i].desc : "");
    }
    printf("\n");
}
int main(int argc, char **argv)
{
    int c, flags = 0, show_v = 0, show_b = 0, show_c = 0, show_s = 0, show_A = 0;
    ucc_config_print_flags_t pf = (ucc_config_print_flags_t)0;
    while ((c = getopt(argc, argv, "vbcafsAh")) != -1) {
        switch (c) {
        case 'v': show_v = 1; break; case 'b': show_b = 1; break; case 'c': show_c = 1; pf |= UCC_CONFIG_PRINT_CONFIG; break;
        case 'a': pf |= UCC_CONFIG_PRINT_HIDDEN; break; case 'f': pf |= UCC_CONFIG_PRINT_CONFIG | UCC_CONFIG_PRINT_HEADER | UCC_CONFIG_PRINT_DOC; break;
        case 's': show_s = 1; break; case 'A': show_A = 1; break; default: usage(); return c == 'h' ? 0 : 1;
        }
        flags = 1;
    }
    if (!flags) { usage(); return 0; }
    if (ucc_constructor() != UCC_OK) { fprintf(stderr, "ucc_info: library initialisation failed\n"); return 1; }
    if (show_v) printf("# UCC version=%s revision %s\n", ucc_get_version_string(), UCC_GIT_REVISION);
    if (show_b) printf("# Built for: NVIDIA B200 (sm_100a), CUDA plugins: mc/cuda ec/cuda tl/nvl tl/nccl sysinfo/cuda; core: C11, no UCX dependency\n# module dir: %s\n",
                       ucc_global_config.component_path ? ucc_global_config.component_path : "?");
    if (show_c || (pf & UCC_CONFIG_PRINT_DOC)) ucc_config_parser_print_all_opts(stdout, "UCC_", pf | UCC_CONFIG_PRINT_CONFIG);
    if (show_s) {
        printf("Default CLs scores:");
        for (int i = 0; i < ucc_global_config.cl_framework.n_components; i++) printf(" %s=%u", ucc_global_config.cl_framework.components[i]->name, ucc_global_config.cl_framework.components[i]->score);
        printf("\nDefault TLs scores:");
        for (int i = 0; i < ucc_global_config.tl_framework.n_components; i++) printf(" %s=%u", ucc_global_config.tl_framework.components[i]->name, ucc_global_config.tl_framework.components[i]->score);
        printf("\n");
    }
    if (show_A) {
        for (int i = 0; i < ucc_global_config.cl_framework.n_components; i++) { ucc_cl_iface_t *cl = ucc_derived_of(ucc_global_config.cl_framework.components[i], ucc_cl_iface_t); print_algs("cl", cl->super.name, cl->alg_info); }
        for (int i = 0; i < ucc_global_config.tl_framework.n_components; i++) { ucc_tl_iface_t *tl = ucc_derived_of(ucc_global_config.tl_framework.components[i], ucc_tl_iface_t); print_algs("tl", tl->super.name, tl->alg_info); }
    }
    return 0;
}
