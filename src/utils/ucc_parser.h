/* Configuration system (UCS-free).
 *
 * Declarative tables {name, default, doc, offset, type} -> typed option
 * structs.  Value sources, lowest to highest priority: table default,
 * ucc.conf file, environment `[<PREFIX>_]UCC_[<TABLE_PREFIX>]<NAME>`.
 * A table with prefix "CL_BASIC_" also honours the un-prefixed variable
 * (UCC_TLS for UCC_CL_BASIC_TLS) like the reference's parser does
 * (behaviour documented in SURVEY.md §5.6; reference utils/ucc_parser.h).
 */
#ifndef UCC_PARSER_H_
#define UCC_PARSER_H_
#include "ucc_compiler_def.h"
#include "ucc_list.h"
#include "ucc_log.h"
#include <ucc/api/ucc.h>
#include <stdio.h>

typedef struct ucc_config_parser {
    int  (*read)(const char *buf, void *dest, const void *arg);            /* 1 ok, 0 error */
    int  (*write)(char *buf, size_t max, const void *src, const void *arg);
    ucc_status_t (*clone)(const void *src, void *dest, const void *arg);
    void (*release)(void *ptr, const void *arg);
    void (*help)(char *buf, size_t max, const void *arg);
    const void *arg;
} ucc_config_parser_t;

typedef struct ucc_config_field {
    const char         *name;
    const char         *dfl_value;
    const char         *doc;
    size_t              offset;
    ucc_config_parser_t parser;
} ucc_config_field_t;

typedef struct ucc_config_global_list_entry {
    const char         *name;   /* human readable */
    const char         *prefix; /* table prefix e.g. "TL_NVL_" */
    ucc_config_field_t *table;
    size_t              size;
    ucc_list_link_t     list;
} ucc_config_global_list_entry_t;

extern ucc_list_link_t ucc_config_global_list;

#define UCC_CONFIG_REGISTER_TABLE(_table, _name, _prefix, _type)                                   \
    ucc_config_global_list_entry_t _table##_config_entry = {_name, _prefix, _table, sizeof(_type), \
                                                             {NULL, NULL}};                         \
    static void UCC_CTOR _table##_config_register(void)                                            \
    { ucc_config_table_register(&_table##_config_entry); }
void ucc_config_table_register(ucc_config_global_list_entry_t *e);
void ucc_config_parser_warn_unused_env_vars_once(void); /* UCC_* variables that no registered table consumes */

typedef struct ucc_config_names_array { char **names; unsigned count; unsigned pad; } ucc_config_names_array_t;
typedef enum { UCC_CONFIG_ALLOW_LIST_ALLOW_ALL, UCC_CONFIG_ALLOW_LIST_ALLOW, UCC_CONFIG_ALLOW_LIST_NEGATE } ucc_config_allow_list_mode_t;
typedef struct ucc_config_allow_list { ucc_config_names_array_t array; ucc_config_allow_list_mode_t mode; } ucc_config_allow_list_t;
typedef struct ucc_config_names_list { ucc_config_names_array_t array; int requested; /* explicitly set by user */ } ucc_config_names_list_t;
typedef enum { UCC_NO = 0, UCC_YES = 1, UCC_TRY = 2, UCC_AUTO = 3 } ucc_ternary_auto_value_t;

/* per message-range / memory-type unsigned value: "0-4k:host:8,4k-inf:4,auto" */
typedef struct ucc_mrange { ucc_list_link_t list; size_t start, end; uint32_t mtypes; unsigned value; } ucc_mrange_t;
typedef struct ucc_mrange_uint { ucc_list_link_t ranges; unsigned default_value; } ucc_mrange_uint_t;
unsigned ucc_mrange_uint_get(const ucc_mrange_uint_t *r, size_t msgsize, ucc_memory_type_t mt);

typedef enum { UCC_PIPELINE_PARALLEL, UCC_PIPELINE_ORDERED, UCC_PIPELINE_SEQUENTIAL, UCC_PIPELINE_LAST } ucc_pipeline_order_t;
typedef struct ucc_pipeline_params {
    size_t threshold; size_t frag_size; unsigned n_frags; unsigned pdepth; ucc_pipeline_order_t order;
} ucc_pipeline_params_t;

/* type parsers */
int  ucc_config_sscanf_string(const char *, void *, const void *);
int  ucc_config_sprintf_string(char *, size_t, const void *, const void *);
ucc_status_t ucc_config_clone_string(const void *, void *, const void *);
void ucc_config_release_string(void *, const void *);
int  ucc_config_sscanf_int(const char *, void *, const void *);
int  ucc_config_sprintf_int(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_uint(const char *, void *, const void *);
int  ucc_config_sprintf_uint(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_ulong(const char *, void *, const void *);
int  ucc_config_sprintf_ulong(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_double(const char *, void *, const void *);
int  ucc_config_sprintf_double(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_bool(const char *, void *, const void *);
int  ucc_config_sprintf_bool(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_ternary(const char *, void *, const void *);
int  ucc_config_sprintf_ternary(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_memunits(const char *, void *, const void *);
int  ucc_config_sprintf_memunits(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_ulunits(const char *, void *, const void *);
int  ucc_config_sprintf_ulunits(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_enum(const char *, void *, const void *);
int  ucc_config_sprintf_enum(char *, size_t, const void *, const void *);
void ucc_config_help_enum(char *, size_t, const void *);
int  ucc_config_sscanf_time(const char *, void *, const void *);
int  ucc_config_sprintf_time(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_array(const char *, void *, const void *);
int  ucc_config_sprintf_array(char *, size_t, const void *, const void *);
ucc_status_t ucc_config_clone_array(const void *, void *, const void *);
void ucc_config_release_array(void *, const void *);
int  ucc_config_sscanf_allow_list(const char *, void *, const void *);
int  ucc_config_sprintf_allow_list(char *, size_t, const void *, const void *);
ucc_status_t ucc_config_clone_allow_list(const void *, void *, const void *);
void ucc_config_release_allow_list(void *, const void *);
int  ucc_config_sscanf_names_list(const char *, void *, const void *);
int  ucc_config_sprintf_names_list(char *, size_t, const void *, const void *);
ucc_status_t ucc_config_clone_names_list(const void *, void *, const void *);
void ucc_config_release_names_list(void *, const void *);
int  ucc_config_sscanf_uint_ranged(const char *, void *, const void *);
int  ucc_config_sprintf_uint_ranged(char *, size_t, const void *, const void *);
ucc_status_t ucc_config_clone_uint_ranged(const void *, void *, const void *);
void ucc_config_release_uint_ranged(void *, const void *);
int  ucc_config_sscanf_pipeline_params(const char *, void *, const void *);
int  ucc_config_sprintf_pipeline_params(char *, size_t, const void *, const void *);
int  ucc_config_sscanf_table(const char *, void *, const void *);
ucc_status_t ucc_config_clone_pod(const void *, void *, const void *); /* size from help prefix "I:" "U:" "L:" "D:" "P:" */
ucc_status_t ucc_config_clone_enum(const void *, void *, const void *);
void ucc_config_release_nop(void *, const void *);
void ucc_config_help_generic(char *, size_t, const void *);

#define UCC_CFG_POD_(_sz, _r, _w, _h) {_r, _w, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, (const void *)(_h)}
#define UCC_CONFIG_TYPE_STRING   {ucc_config_sscanf_string, ucc_config_sprintf_string, ucc_config_clone_string, ucc_config_release_string, ucc_config_help_generic, "string"}
#define UCC_CONFIG_TYPE_INT      {ucc_config_sscanf_int, ucc_config_sprintf_int, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "I:integer"}
#define UCC_CONFIG_TYPE_UINT     {ucc_config_sscanf_uint, ucc_config_sprintf_uint, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "U:unsigned integer or 'auto'/'inf'"}
#define UCC_CONFIG_TYPE_ULONG    {ucc_config_sscanf_ulong, ucc_config_sprintf_ulong, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "L:unsigned long"}
#define UCC_CONFIG_TYPE_DOUBLE   {ucc_config_sscanf_double, ucc_config_sprintf_double, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "D:floating point number"}
#define UCC_CONFIG_TYPE_BOOL     {ucc_config_sscanf_bool, ucc_config_sprintf_bool, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "I:<y|n>"}
#define UCC_CONFIG_TYPE_TERNARY  {ucc_config_sscanf_ternary, ucc_config_sprintf_ternary, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "I:<yes|no|try|auto>"}
#define UCC_CONFIG_TYPE_MEMUNITS {ucc_config_sscanf_memunits, ucc_config_sprintf_memunits, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "L:memory units: <number>[b|kb|mb|gb], \"inf\", or \"auto\""}
#define UCC_CONFIG_TYPE_ULUNITS  {ucc_config_sscanf_ulunits, ucc_config_sprintf_ulunits, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "L:unsigned long: <number>, \"inf\", or \"auto\""}
#define UCC_CONFIG_TYPE_TIME     {ucc_config_sscanf_time, ucc_config_sprintf_time, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "D:time value: <number>[s|us|ms|ns|m]"}
#define UCC_CONFIG_TYPE_ENUM(_names) {ucc_config_sscanf_enum, ucc_config_sprintf_enum, ucc_config_clone_enum, ucc_config_release_nop, ucc_config_help_enum, (_names)}
#define UCC_CONFIG_TYPE_STRING_ARRAY {ucc_config_sscanf_array, ucc_config_sprintf_array, ucc_config_clone_array, ucc_config_release_array, ucc_config_help_generic, "comma-separated list of strings"}
#define UCC_CONFIG_TYPE_ALLOW_LIST   {ucc_config_sscanf_allow_list, ucc_config_sprintf_allow_list, ucc_config_clone_allow_list, ucc_config_release_allow_list, ucc_config_help_generic, "comma-separated list (use \"all\" for all, prefix with ^ for negation)"}
#define UCC_CONFIG_TYPE_NAMES_LIST   {ucc_config_sscanf_names_list, ucc_config_sprintf_names_list, ucc_config_clone_names_list, ucc_config_release_names_list, ucc_config_help_generic, "comma-separated list of component names or \"all\""}
#define UCC_CONFIG_TYPE_UINT_RANGED  {ucc_config_sscanf_uint_ranged, ucc_config_sprintf_uint_ranged, ucc_config_clone_uint_ranged, ucc_config_release_uint_ranged, ucc_config_help_generic, "[<msg_start>-<msg_end>:[mtype:]value,...,]default_value  (values: uint|auto|inf)"}
#define UCC_CONFIG_TYPE_PIPELINE_PARAMS {ucc_config_sscanf_pipeline_params, ucc_config_sprintf_pipeline_params, ucc_config_clone_pod, ucc_config_release_nop, ucc_config_help_generic, "P:thresh=<memunits>:fragsize=<memunits>:nfrags=<uint>:pdepth=<uint>:<ordered|parallel|sequential>"}
#define UCC_CONFIG_TYPE_TABLE(_t)    {ucc_config_sscanf_table, NULL, NULL, NULL, NULL, (_t)}

size_t ucc_config_pod_size(const ucc_config_parser_t *p);

/* API */
ucc_status_t ucc_config_parser_fill_opts(void *opts, ucc_config_global_list_entry_t *entry, const char *env_prefix,
                                         int ignore_errors);
ucc_status_t ucc_config_parser_fill_opts_table(void *opts, ucc_config_field_t *table, const char *env_prefix,
                                               const char *table_prefix, int ignore_errors);
void         ucc_config_parser_release_opts(void *opts, ucc_config_field_t *table);
ucc_status_t ucc_config_parser_set_value(void *opts, ucc_config_field_t *table, const char *name, const char *value);
ucc_status_t ucc_config_parser_get_value(void *opts, ucc_config_field_t *table, const char *name, char *value, size_t max);
ucc_status_t ucc_config_parser_clone_opts(const void *src, void *dst, ucc_config_field_t *table);
void         ucc_config_parser_print_opts(FILE *stream, const char *title, const void *opts, ucc_config_field_t *table,
                                          const char *table_prefix, const char *env_prefix,
                                          ucc_config_print_flags_t flags);
void         ucc_config_parser_print_all_opts(FILE *stream, const char *env_prefix, ucc_config_print_flags_t flags);
int          ucc_config_names_search(const ucc_config_names_array_t *arr, const char *name); /* index or -1 */
ucc_status_t ucc_config_names_array_dup(ucc_config_names_array_t *dst, const ucc_config_names_array_t *src);
void         ucc_config_names_array_free(ucc_config_names_array_t *a);
ucc_status_t ucc_config_names_array_merge(ucc_config_names_array_t *dst, const ucc_config_names_array_t *src);
ucc_status_t ucc_config_allow_list_process(const ucc_config_allow_list_t *list, const ucc_config_names_array_t *all,
                                           ucc_config_names_list_t *out);

/* ---- ucc.conf file ---- */
typedef struct ucc_file_section_filter { /* runtime facts a [section] may be predicated on */
    const char *vendor; const char *model; unsigned team_size; unsigned ppn; unsigned nnodes; unsigned sock;
} ucc_file_section_filter_t;
typedef struct ucc_file_config ucc_file_config_t;
ucc_status_t ucc_parse_file_config(const char *filename, ucc_file_config_t **cfg);
void         ucc_release_file_config(ucc_file_config_t *cfg);
/* look `var` (full name e.g. "UCC_TL_SHM_TUNE") up; sections only considered if filter!=NULL and matches */
const char  *ucc_file_config_lookup(const ucc_file_config_t *cfg, const char *var, const ucc_file_section_filter_t *filter);
/* apply section-scoped overrides of a team-level table (called at team create) */
ucc_status_t ucc_apply_file_cfg_sections(void *opts, ucc_config_field_t *table, const char *env_prefix,
                                         const char *table_prefix, const ucc_file_section_filter_t *filter);
#endif
