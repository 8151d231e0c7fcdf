/* Component framework: static registry + dlopen of libucc_<fw>_<name>.so
 * modules exporting a `ucc_<fw>_<name>` iface symbol
 * (convention of reference utils/ucc_component.c:25-97). */
#ifndef UCC_COMPONENT_H_
#define UCC_COMPONENT_H_
#include "ucc_compiler_def.h"
#include "ucc_parser.h"
#define UCC_MAX_FRAMEWORK_NAME_LEN 64
#define UCC_MAX_COMPONENT_NAME_LEN 64

typedef struct ucc_component_iface {
    const char   *name;
    unsigned long id;      /* djb2(name) */
    void         *handle;  /* dlopen handle or NULL for built-ins */
    ucc_score_t   score;   /* default score; must be unique within a framework */
} ucc_component_iface_t;

typedef struct ucc_component_framework {
    char                     *framework_name;
    int                       n_components;
    ucc_component_iface_t   **components;
    ucc_config_names_array_t  names;
} ucc_component_framework_t;

/* built-in (statically linked) components register themselves from a constructor */
void ucc_component_register_static(const char *framework, ucc_component_iface_t *iface);
ucc_status_t ucc_components_load(const char *framework_name, ucc_component_framework_t *framework);
ucc_component_iface_t *ucc_get_component(ucc_component_framework_t *framework, const char *component_name);
ucc_status_t ucc_component_check_scores_uniq(ucc_component_framework_t *framework);
void ucc_components_unload(ucc_component_framework_t *framework);

#define UCC_COMPONENT_REGISTER_STATIC(_fw, _iface)                                        \
    static void UCC_CTOR ucc_static_register_##_fw##_##_iface(void)                       \
    { ucc_component_register_static(#_fw, (ucc_component_iface_t *)&(_iface)); }
#endif
