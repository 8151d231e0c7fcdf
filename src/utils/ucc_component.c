#include "ucc_component.h"
#include "ucc_string.h"
#include "core/ucc_global_opts.h"
#include <dlfcn.h>
#include <glob.h>
#include <libgen.h>

#define MAX_STATIC 64
static struct { const char *fw; ucc_component_iface_t *iface; } static_comps[MAX_STATIC];
static int n_static = 0;

void ucc_component_register_static(const char *framework, ucc_component_iface_t *iface)
{
    if (n_static < MAX_STATIC) { static_comps[n_static].fw = framework; static_comps[n_static].iface = iface; n_static++; }
}

static void framework_add(ucc_component_framework_t *fw, ucc_component_iface_t *iface)
{
    for (int i = 0; i < fw->n_components; i++) if (!strcmp(fw->components[i]->name, iface->name)) return;
    fw->components = (ucc_component_iface_t **)realloc(fw->components, (fw->n_components + 1) * sizeof(void *));
    fw->components[fw->n_components++] = iface;
    iface->id = ucc_str_hash_djb2(iface->name);
    fw->names.names = (char **)realloc(fw->names.names, (fw->names.count + 1) * sizeof(char *));
    fw->names.names[fw->names.count++] = strdup(iface->name);
}

/* module file name: libucc_<fw>_<name>.so -> iface symbol ucc_<fw>_<name> */
static ucc_status_t load_one(const char *path, const char *fw_name, ucc_component_framework_t *fw)
{
    char  sym[2 * UCC_MAX_COMPONENT_NAME_LEN + 16], base[512], *b, *dot;
    void *h;
    ucc_component_iface_t *iface;
    snprintf(base, sizeof(base), "%s", path);
    b   = basename(base);
    dot = strstr(b, ".so");
    if (!dot || strncmp(b, "libucc_", 7)) return UCC_ERR_INVALID_PARAM;
    *dot = 0;
    snprintf(sym, sizeof(sym), "%s", b + 3); /* ucc_<fw>_<name> */
    h = dlopen(path, RTLD_LAZY | RTLD_GLOBAL);
    if (!h) { ucc_debug("component %s: dlopen failed: %s", path, dlerror()); return UCC_ERR_NO_RESOURCE; }
    iface = (ucc_component_iface_t *)dlsym(h, sym);
    if (!iface) { ucc_debug("component %s: no symbol %s", path, sym); dlclose(h); return UCC_ERR_NOT_FOUND; }
    (void)fw_name;
    iface->handle = h;
    framework_add(fw, iface);
    return UCC_OK;
}

ucc_status_t ucc_components_load(const char *framework_name, ucc_component_framework_t *fw)
{
    char   pattern[1024];
    glob_t g;
    const char *dir = (ucc_global_config.module_dir && ucc_global_config.module_dir[0]) ? ucc_global_config.module_dir
                                                                                      : ucc_global_config.component_path;
    memset(fw, 0, sizeof(*fw));
    fw->framework_name = strdup(framework_name);
    for (int i = 0; i < n_static; i++) if (!strcmp(static_comps[i].fw, framework_name)) framework_add(fw, static_comps[i].iface);
    if (dir && dir[0]) {
        snprintf(pattern, sizeof(pattern), "%s/libucc_%s_*.so", dir, framework_name);
        if (glob(pattern, 0, NULL, &g) == 0) {
            for (size_t i = 0; i < g.gl_pathc; i++) load_one(g.gl_pathv[i], framework_name, fw);
            globfree(&g);
        }
    }
    return fw->n_components ? UCC_OK : UCC_ERR_NOT_FOUND;
}

ucc_component_iface_t *ucc_get_component(ucc_component_framework_t *fw, const char *name)
{
    for (int i = 0; i < fw->n_components; i++) if (!strcmp(fw->components[i]->name, name)) return fw->components[i];
    return NULL;
}

ucc_status_t ucc_component_check_scores_uniq(ucc_component_framework_t *fw)
{
    for (int i = 0; i < fw->n_components; i++)
        for (int j = i + 1; j < fw->n_components; j++)
            if (fw->components[i]->score == fw->components[j]->score) {
                ucc_error("components %s and %s of framework %s have the same default score %u", fw->components[i]->name,
                          fw->components[j]->name, fw->framework_name, fw->components[i]->score);
                return UCC_ERR_INVALID_PARAM;
            }
    return UCC_OK;
}

void ucc_components_unload(ucc_component_framework_t *fw)
{
    /* dlopen handles are intentionally kept: component code (e.g. CUDA
     * runtime registrations) must outlive finalize */
    free(fw->components); fw->components = NULL; fw->n_components = 0;
    ucc_config_names_array_free(&fw->names);
    free(fw->framework_name); fw->framework_name = NULL;
}
