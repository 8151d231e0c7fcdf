/* Configuration parser implementation — see ucc_parser.h. */
#include "ucc_parser.h"
#include "ucc_string.h"
#include "ucc_math.h"
#include "core/ucc_global_opts.h"
#include <ctype.h>
#include <strings.h>
#include <limits.h>

ucc_list_link_t ucc_config_global_list = UCC_LIST_INITIALIZER(&ucc_config_global_list);

void ucc_config_table_register(ucc_config_global_list_entry_t *e)
{
    ucc_config_global_list_entry_t *it;
    ucc_list_for_each(it, &ucc_config_global_list, list) { if (it == e) return; }
    ucc_list_add_tail(&ucc_config_global_list, &e->list);
}

/* Does UCC_<rest> name a field of `table` (directly as <table prefix><field> or through the short-name fall-back)? */
static int env_names_field(const char *rest, const ucc_config_field_t *table, const char *tprefix)
{
    size_t pl = strlen(tprefix);
    for (const ucc_config_field_t *f = table; f->name; f++) {
        if (f->parser.read == ucc_config_sscanf_table) { if (env_names_field(rest, (const ucc_config_field_t *)f->parser.arg, tprefix)) return 1; continue; }
        if (!f->name[0]) continue;
        if (!strcmp(rest, f->name)) return 1;
        if (pl && !strncmp(rest, tprefix, pl) && !strcmp(rest + pl, f->name)) return 1;
    }
    return 0;
}

/* One-time warning about UCC_* environment variables that no registered configuration table consumes (typos are the
 * usual reason).  Same service as the UCS parser gives the reference (UCC_WARN_UNUSED_ENV_VARS). */
void ucc_config_parser_warn_unused_env_vars_once(void)
{
    static int done = 0;
    static const char *direct[] = {"UCC_DEBUGGER_WAIT", "UCC_TL_NCCL_LIB", "UCC_MODULE_DIR", NULL}; /* read with getenv() */
    extern char **environ;
    char unused[1024]; size_t len = 0; int n = 0;
    if (__atomic_exchange_n(&done, 1, __ATOMIC_ACQ_REL)) return;
    for (char **e = environ; e && *e; e++) {
        const char *eq = strchr(*e, '='), *p = strstr(*e, "UCC_");
        char name[256]; ucc_config_global_list_entry_t *it; int used = 0;
        if (!eq || !p || p > eq || (size_t)(eq - *e) >= sizeof(name)) continue;
        if (p != *e && p[-1] != '_') continue; /* "<APP>_UCC_..." is fine, "FOOUCC_" is not ours */
        memcpy(name, *e, (size_t)(eq - *e)); name[eq - *e] = 0;
        for (int i = 0; direct[i]; i++) if (!strcmp(name + (p - *e), direct[i])) used = 1;
        ucc_list_for_each(it, &ucc_config_global_list, list) { if (used) break; used = env_names_field(name + (p - *e) + 4, it->table, it->prefix ? it->prefix : ""); }
        if (!used && len + strlen(name) + 3 < sizeof(unused)) { len += (size_t)snprintf(unused + len, sizeof(unused) - len, "%s%s", n ? "; " : "", name); n++; }
    }
    if (n) ucc_warn("unused environment variable%s: %s (set UCC_WARN_UNUSED_ENV_VARS=n to suppress this warning)", n > 1 ? "s" : "", unused);
}

/* ------------------------------------------------------------------ */
/* scalar types                                                        */
/* ------------------------------------------------------------------ */
int ucc_config_sscanf_string(const char *buf, void *dest, const void *arg)
{ (void)arg; *(char **)dest = strdup(buf); return *(char **)dest != NULL; }
int ucc_config_sprintf_string(char *buf, size_t max, const void *src, const void *arg)
{ (void)arg; snprintf(buf, max, "%s", *(char *const *)src ? *(char *const *)src : ""); return 1; }
ucc_status_t ucc_config_clone_string(const void *src, void *dest, const void *arg)
{
    (void)arg;
    const char *s = *(char *const *)src;
    *(char **)dest = s ? strdup(s) : NULL;
    return (s && !*(char **)dest) ? UCC_ERR_NO_MEMORY : UCC_OK;
}
void ucc_config_release_string(void *ptr, const void *arg) { (void)arg; free(*(char **)ptr); *(char **)ptr = NULL; }

int ucc_config_sscanf_int(const char *buf, void *dest, const void *arg)
{
    (void)arg; char *e; long v = strtol(buf, &e, 0);
    if (e == buf || *e) return 0;
    *(int *)dest = (int)v; return 1;
}
int ucc_config_sprintf_int(char *buf, size_t max, const void *src, const void *arg)
{ (void)arg; snprintf(buf, max, "%d", *(const int *)src); return 1; }

int ucc_config_sscanf_uint(const char *buf, void *dest, const void *arg)
{
    (void)arg; char *e; unsigned long v;
    if (!strcasecmp(buf, "inf")) { *(unsigned *)dest = UINT_MAX; return 1; }
    if (!strcasecmp(buf, "auto")) { *(unsigned *)dest = UCC_UUNITS_AUTO; return 1; }
    if (*buf == '-') return 0;
    v = strtoul(buf, &e, 0);
    if (e == buf || *e) return 0;
    *(unsigned *)dest = (unsigned)v; return 1;
}
int ucc_config_sprintf_uint(char *buf, size_t max, const void *src, const void *arg)
{
    (void)arg; unsigned v = *(const unsigned *)src;
    if (v == UINT_MAX) snprintf(buf, max, "inf"); else if (v == UCC_UUNITS_AUTO) snprintf(buf, max, "auto");
    else snprintf(buf, max, "%u", v);
    return 1;
}
int ucc_config_sscanf_ulong(const char *buf, void *dest, const void *arg)
{
    (void)arg; char *e; unsigned long v;
    if (*buf == '-') return 0;
    v = strtoul(buf, &e, 0);
    if (e == buf || *e) return 0;
    *(unsigned long *)dest = v; return 1;
}
int ucc_config_sprintf_ulong(char *buf, size_t max, const void *src, const void *arg)
{ (void)arg; snprintf(buf, max, "%lu", *(const unsigned long *)src); return 1; }

int ucc_config_sscanf_double(const char *buf, void *dest, const void *arg)
{
    (void)arg; char *e; double v = strtod(buf, &e);
    if (e == buf || *e) return 0;
    *(double *)dest = v; return 1;
}
int ucc_config_sprintf_double(char *buf, size_t max, const void *src, const void *arg)
{ (void)arg; snprintf(buf, max, "%g", *(const double *)src); return 1; }

int ucc_config_sscanf_bool(const char *buf, void *dest, const void *arg)
{
    (void)arg;
    if (!strcasecmp(buf, "y") || !strcasecmp(buf, "yes") || !strcmp(buf, "1") || !strcasecmp(buf, "on") ||
        !strcasecmp(buf, "true")) { *(int *)dest = 1; return 1; }
    if (!strcasecmp(buf, "n") || !strcasecmp(buf, "no") || !strcmp(buf, "0") || !strcasecmp(buf, "off") ||
        !strcasecmp(buf, "false")) { *(int *)dest = 0; return 1; }
    return 0;
}
int ucc_config_sprintf_bool(char *buf, size_t max, const void *src, const void *arg)
{ (void)arg; snprintf(buf, max, "%c", *(const int *)src ? 'y' : 'n'); return 1; }

int ucc_config_sscanf_ternary(const char *buf, void *dest, const void *arg)
{
    if (!strcasecmp(buf, "try") || !strcasecmp(buf, "maybe")) { *(int *)dest = UCC_TRY; return 1; }
    if (!strcasecmp(buf, "auto")) { *(int *)dest = UCC_AUTO; return 1; }
    return ucc_config_sscanf_bool(buf, dest, arg);
}
int ucc_config_sprintf_ternary(char *buf, size_t max, const void *src, const void *arg)
{
    static const char *n[] = {"no", "yes", "try", "auto"};
    (void)arg; snprintf(buf, max, "%s", n[(*(const int *)src) & 3]); return 1;
}

int ucc_config_sscanf_memunits(const char *buf, void *dest, const void *arg)
{ (void)arg; return ucc_str_to_memunits(buf, (size_t *)dest) == UCC_OK; }
int ucc_config_sprintf_memunits(char *buf, size_t max, const void *src, const void *arg)
{ (void)arg; ucc_memunits_to_str(*(const size_t *)src, buf, max); return 1; }

int ucc_config_sscanf_ulunits(const char *buf, void *dest, const void *arg)
{
    if (!strcasecmp(buf, "inf")) { *(unsigned long *)dest = ULONG_MAX; return 1; }
    if (!strcasecmp(buf, "auto")) { *(unsigned long *)dest = UCC_ULUNITS_AUTO; return 1; }
    return ucc_config_sscanf_ulong(buf, dest, arg);
}
int ucc_config_sprintf_ulunits(char *buf, size_t max, const void *src, const void *arg)
{
    unsigned long v = *(const unsigned long *)src;
    if (v == ULONG_MAX) snprintf(buf, max, "inf"); else if (v == UCC_ULUNITS_AUTO) snprintf(buf, max, "auto");
    else return ucc_config_sprintf_ulong(buf, max, src, arg);
    return 1;
}

int ucc_config_sscanf_enum(const char *buf, void *dest, const void *arg)
{
    int i = ucc_str_find_in_list(buf, (const char **)arg);
    if (i < 0) return 0;
    *(unsigned *)dest = (unsigned)i; return 1;
}
int ucc_config_sprintf_enum(char *buf, size_t max, const void *src, const void *arg)
{ snprintf(buf, max, "%s", ((const char *const *)arg)[*(const unsigned *)src]); return 1; }
void ucc_config_help_enum(char *buf, size_t max, const void *arg)
{
    size_t o = snprintf(buf, max, "[");
    for (const char *const *n = (const char *const *)arg; *n && o < max; n++)
        o += snprintf(buf + o, max - o, "%s%s", n == (const char *const *)arg ? "" : "|", *n);
    if (o < max) snprintf(buf + o, max - o, "]");
}

int ucc_config_sscanf_time(const char *buf, void *dest, const void *arg)
{
    (void)arg; char *e; double v = strtod(buf, &e), m = 1.0;
    if (e == buf) return 0;
    if (!strcmp(e, "") || !strcmp(e, "s")) m = 1.0;
    else if (!strcmp(e, "ms")) m = 1e-3; else if (!strcmp(e, "us")) m = 1e-6;
    else if (!strcmp(e, "ns")) m = 1e-9; else if (!strcmp(e, "m")) m = 60.0; else return 0;
    *(double *)dest = v * m; return 1;
}
int ucc_config_sprintf_time(char *buf, size_t max, const void *src, const void *arg)
{ (void)arg; snprintf(buf, max, "%.2fus", *(const double *)src * 1e6); return 1; }

ucc_status_t ucc_config_clone_pod(const void *src, void *dest, const void *arg)
{
    const char *h = (const char *)arg;
    size_t      sz;
    switch (h ? h[0] : 'I') {
    case 'L': case 'D': sz = 8; break;
    case 'P': sz = sizeof(ucc_pipeline_params_t); break;
    default: sz = 4; break;
    }
    memcpy(dest, src, sz);
    return UCC_OK;
}
ucc_status_t ucc_config_clone_enum(const void *src, void *dest, const void *arg)
{ (void)arg; memcpy(dest, src, sizeof(unsigned)); return UCC_OK; }
void ucc_config_release_nop(void *ptr, const void *arg) { (void)ptr; (void)arg; }
void ucc_config_help_generic(char *buf, size_t max, const void *arg)
{
    const char *h = (const char *)arg;
    if (h && h[0] && h[1] == ':') h += 2;
    snprintf(buf, max, "%s", h ? h : "");
}

/* ------------------------------------------------------------------ */
/* arrays / lists                                                      */
/* ------------------------------------------------------------------ */
void ucc_config_names_array_free(ucc_config_names_array_t *a)
{
    for (unsigned i = 0; i < a->count; i++) free(a->names[i]);
    free(a->names); a->names = NULL; a->count = 0;
}
ucc_status_t ucc_config_names_array_dup(ucc_config_names_array_t *dst, const ucc_config_names_array_t *src)
{
    dst->count = src->count; dst->pad = 0;
    dst->names = src->count ? (char **)calloc(src->count, sizeof(char *)) : NULL;
    if (src->count && !dst->names) return UCC_ERR_NO_MEMORY;
    for (unsigned i = 0; i < src->count; i++) dst->names[i] = strdup(src->names[i]);
    return UCC_OK;
}
int ucc_config_names_search(const ucc_config_names_array_t *arr, const char *name)
{
    for (unsigned i = 0; i < arr->count; i++) if (!strcmp(arr->names[i], name)) return (int)i;
    return -1;
}
ucc_status_t ucc_config_names_array_merge(ucc_config_names_array_t *dst, const ucc_config_names_array_t *src)
{
    for (unsigned i = 0; i < src->count; i++) {
        if (ucc_config_names_search(dst, src->names[i]) >= 0) continue;
        dst->names = (char **)realloc(dst->names, (dst->count + 1) * sizeof(char *));
        if (!dst->names) return UCC_ERR_NO_MEMORY;
        dst->names[dst->count++] = strdup(src->names[i]);
    }
    return UCC_OK;
}
int ucc_config_sscanf_array(const char *buf, void *dest, const void *arg)
{
    (void)arg;
    ucc_config_names_array_t *a = (ucc_config_names_array_t *)dest;
    char **s = ucc_str_split(buf, ",");
    if (!s) return 0;
    a->count = ucc_str_split_count(s); a->names = s; a->pad = 0;
    for (unsigned i = 0; i < a->count; i++) ucc_str_trim(a->names[i]);
    return 1;
}
int ucc_config_sprintf_array(char *buf, size_t max, const void *src, const void *arg)
{
    (void)arg;
    const ucc_config_names_array_t *a = (const ucc_config_names_array_t *)src;
    size_t o = 0; buf[0] = 0;
    for (unsigned i = 0; i < a->count && o < max; i++) o += snprintf(buf + o, max - o, "%s%s", i ? "," : "", a->names[i]);
    return 1;
}
ucc_status_t ucc_config_clone_array(const void *src, void *dest, const void *arg)
{ (void)arg; return ucc_config_names_array_dup((ucc_config_names_array_t *)dest, (const ucc_config_names_array_t *)src); }
void ucc_config_release_array(void *ptr, const void *arg) { (void)arg; ucc_config_names_array_free((ucc_config_names_array_t *)ptr); }

int ucc_config_sscanf_allow_list(const char *buf, void *dest, const void *arg)
{
    ucc_config_allow_list_t *l = (ucc_config_allow_list_t *)dest;
    const char *p = buf;
    l->mode = UCC_CONFIG_ALLOW_LIST_ALLOW;
    if (*p == '^') { l->mode = UCC_CONFIG_ALLOW_LIST_NEGATE; p++; }
    if (!ucc_config_sscanf_array(p, &l->array, arg)) return 0;
    if (l->array.count == 1 && !strcasecmp(l->array.names[0], "all")) {
        if (l->mode == UCC_CONFIG_ALLOW_LIST_NEGATE) { ucc_config_names_array_free(&l->array); return 0; }
        l->mode = UCC_CONFIG_ALLOW_LIST_ALLOW_ALL; ucc_config_names_array_free(&l->array);
    }
    return 1;
}
int ucc_config_sprintf_allow_list(char *buf, size_t max, const void *src, const void *arg)
{
    const ucc_config_allow_list_t *l = (const ucc_config_allow_list_t *)src;
    if (l->mode == UCC_CONFIG_ALLOW_LIST_ALLOW_ALL) { snprintf(buf, max, "all"); return 1; }
    if (l->mode == UCC_CONFIG_ALLOW_LIST_NEGATE && max > 1) { *buf++ = '^'; max--; }
    return ucc_config_sprintf_array(buf, max, &l->array, arg);
}
ucc_status_t ucc_config_clone_allow_list(const void *src, void *dest, const void *arg)
{
    (void)arg;
    ((ucc_config_allow_list_t *)dest)->mode = ((const ucc_config_allow_list_t *)src)->mode;
    return ucc_config_names_array_dup(&((ucc_config_allow_list_t *)dest)->array, &((const ucc_config_allow_list_t *)src)->array);
}
void ucc_config_release_allow_list(void *ptr, const void *arg)
{ (void)arg; ucc_config_names_array_free(&((ucc_config_allow_list_t *)ptr)->array); }

ucc_status_t ucc_config_allow_list_process(const ucc_config_allow_list_t *list, const ucc_config_names_array_t *all,
                                           ucc_config_names_list_t *out)
{
    out->array.names = NULL; out->array.count = 0; out->array.pad = 0;
    out->requested = (list->mode == UCC_CONFIG_ALLOW_LIST_ALLOW);
    if (list->mode == UCC_CONFIG_ALLOW_LIST_ALLOW_ALL) return ucc_config_names_array_dup(&out->array, all);
    if (list->mode == UCC_CONFIG_ALLOW_LIST_ALLOW) return ucc_config_names_array_dup(&out->array, &list->array);
    for (unsigned i = 0; i < all->count; i++) {
        if (ucc_config_names_search(&list->array, all->names[i]) >= 0) continue;
        out->array.names = (char **)realloc(out->array.names, (out->array.count + 1) * sizeof(char *));
        out->array.names[out->array.count++] = strdup(all->names[i]);
    }
    return UCC_OK;
}

int ucc_config_sscanf_names_list(const char *buf, void *dest, const void *arg)
{
    ucc_config_names_list_t *l = (ucc_config_names_list_t *)dest;
    l->requested = 1;
    return ucc_config_sscanf_array(buf, &l->array, arg);
}
int ucc_config_sprintf_names_list(char *buf, size_t max, const void *src, const void *arg)
{ return ucc_config_sprintf_array(buf, max, &((const ucc_config_names_list_t *)src)->array, arg); }
ucc_status_t ucc_config_clone_names_list(const void *src, void *dest, const void *arg)
{
    (void)arg;
    ((ucc_config_names_list_t *)dest)->requested = ((const ucc_config_names_list_t *)src)->requested;
    return ucc_config_names_array_dup(&((ucc_config_names_list_t *)dest)->array, &((const ucc_config_names_list_t *)src)->array);
}
void ucc_config_release_names_list(void *ptr, const void *arg)
{ (void)arg; ucc_config_names_array_free(&((ucc_config_names_list_t *)ptr)->array); }

/* ------------------------------------------------------------------ */
/* ranged uint: "0-4k:host:8,4k-inf:4,auto"                            */
/* ------------------------------------------------------------------ */
static const char *mtype_names[] = {"host", "cuda", "cuda_managed", "rocm", "rocm_managed", NULL};

static int mtype_from_str(const char *s)
{
    if (!strcasecmp(s, "cpu")) return UCC_MEMORY_TYPE_HOST;
    if (!strcasecmp(s, "cudamanaged")) return UCC_MEMORY_TYPE_CUDA_MANAGED;
    return ucc_str_find_in_list(s, mtype_names);
}

void ucc_config_release_uint_ranged(void *ptr, const void *arg)
{
    (void)arg;
    ucc_mrange_uint_t *r = (ucc_mrange_uint_t *)ptr;
    ucc_mrange_t *m, *t;
    if (!r->ranges.next) return;
    ucc_list_for_each_safe(m, t, &r->ranges, list) { ucc_list_del(&m->list); free(m); }
}
int ucc_config_sscanf_uint_ranged(const char *buf, void *dest, const void *arg)
{
    ucc_mrange_uint_t *r = (ucc_mrange_uint_t *)dest;
    char **toks = ucc_str_split(buf, ",");
    unsigned n = ucc_str_split_count(toks);
    int ok = 1, have_default = 0;
    ucc_list_head_init(&r->ranges);
    r->default_value = UCC_UUNITS_AUTO;
    for (unsigned i = 0; i < n && ok; i++) {
        char **parts = ucc_str_split(toks[i], ":");
        unsigned np = ucc_str_split_count(parts);
        if (np == 1) { /* default value */
            ok = !have_default && ucc_config_sscanf_uint(parts[0], &r->default_value, arg);
            have_default = 1;
        } else if (np == 2 || np == 3) {
            ucc_mrange_t *m = (ucc_mrange_t *)calloc(1, sizeof(*m));
            m->mtypes = 0xffffffffu;
            ok = ucc_str_memunits_range_to_ulong(parts[0], &m->start, &m->end) == UCC_OK;
            if (ok && np == 3) { int mt = mtype_from_str(parts[1]); ok = mt >= 0; if (ok) m->mtypes = 1u << mt; }
            if (ok) ok = ucc_config_sscanf_uint(parts[np - 1], &m->value, arg);
            if (ok) ucc_list_add_tail(&r->ranges, &m->list); else free(m);
        } else ok = 0;
        ucc_str_split_free(parts);
    }
    ucc_str_split_free(toks);
    if (!ok) ucc_config_release_uint_ranged(r, arg);
    return ok;
}
int ucc_config_sprintf_uint_ranged(char *buf, size_t max, const void *src, const void *arg)
{
    const ucc_mrange_uint_t *r = (const ucc_mrange_uint_t *)src;
    ucc_mrange_t *m; size_t o = 0; char a[32], b[32], v[32];
    buf[0] = 0;
    if (r->ranges.next) ucc_list_for_each(m, &r->ranges, list) {
        ucc_memunits_to_str(m->start, a, sizeof(a)); ucc_memunits_to_str(m->end, b, sizeof(b));
        ucc_config_sprintf_uint(v, sizeof(v), &m->value, arg);
        if (m->mtypes != 0xffffffffu) o += snprintf(buf + o, max - o, "%s-%s:%s:%s,", a, b, mtype_names[ucc_ilog2(m->mtypes)], v);
        else o += snprintf(buf + o, max - o, "%s-%s:%s,", a, b, v);
        if (o >= max) return 1;
    }
    ucc_config_sprintf_uint(v, sizeof(v), &r->default_value, arg);
    snprintf(buf + o, max - o, "%s", v);
    return 1;
}
ucc_status_t ucc_config_clone_uint_ranged(const void *src, void *dest, const void *arg)
{
    (void)arg;
    const ucc_mrange_uint_t *s = (const ucc_mrange_uint_t *)src;
    ucc_mrange_uint_t *d = (ucc_mrange_uint_t *)dest;
    ucc_mrange_t *m;
    ucc_list_head_init(&d->ranges);
    d->default_value = s->default_value;
    if (s->ranges.next) ucc_list_for_each(m, &s->ranges, list) {
        ucc_mrange_t *c = (ucc_mrange_t *)malloc(sizeof(*c));
        if (!c) return UCC_ERR_NO_MEMORY;
        *c = *m; ucc_list_add_tail(&d->ranges, &c->list);
    }
    return UCC_OK;
}
unsigned ucc_mrange_uint_get(const ucc_mrange_uint_t *r, size_t msgsize, ucc_memory_type_t mt)
{
    ucc_mrange_t *m;
    if (r->ranges.next) ucc_list_for_each(m, &r->ranges, list) {
        if (msgsize >= m->start && msgsize <= m->end && (m->mtypes & (1u << mt))) return m->value;
    }
    return r->default_value;
}

/* pipeline params: thresh=…:fragsize=…:nfrags=…:pdepth=…:<order> | n | auto */
static const char *pipeline_order_names[] = {"parallel", "ordered", "sequential", NULL};
int ucc_config_sscanf_pipeline_params(const char *buf, void *dest, const void *arg)
{
    ucc_pipeline_params_t *p = (ucc_pipeline_params_t *)dest;
    char **toks; unsigned n; int ok = 1;
    (void)arg;
    p->threshold = UCC_MEMUNITS_INF; p->frag_size = UCC_MEMUNITS_INF; p->n_frags = 2; p->pdepth = 2;
    p->order = UCC_PIPELINE_PARALLEL;
    if (!strcasecmp(buf, "n") || !strcasecmp(buf, "no")) { p->n_frags = 0; p->pdepth = 0; return 1; }
    if (!strcasecmp(buf, "auto")) { p->threshold = UCC_MEMUNITS_AUTO; p->frag_size = UCC_MEMUNITS_AUTO; return 1; }
    toks = ucc_str_split(buf, ":"); n = ucc_str_split_count(toks);
    for (unsigned i = 0; i < n && ok; i++) {
        char *eq = strchr(toks[i], '=');
        if (!eq) { int o = ucc_str_find_in_list(toks[i], pipeline_order_names); ok = o >= 0; if (ok) p->order = (ucc_pipeline_order_t)o; continue; }
        *eq++ = 0;
        if (!strcasecmp(toks[i], "thresh")) ok = ucc_str_to_memunits(eq, &p->threshold) == UCC_OK;
        else if (!strcasecmp(toks[i], "fragsize")) ok = ucc_str_to_memunits(eq, &p->frag_size) == UCC_OK;
        else if (!strcasecmp(toks[i], "nfrags")) ok = ucc_config_sscanf_uint(eq, &p->n_frags, NULL);
        else if (!strcasecmp(toks[i], "pdepth")) ok = ucc_config_sscanf_uint(eq, &p->pdepth, NULL);
        else ok = 0;
    }
    ucc_str_split_free(toks);
    return ok;
}
int ucc_config_sprintf_pipeline_params(char *buf, size_t max, const void *src, const void *arg)
{
    const ucc_pipeline_params_t *p = (const ucc_pipeline_params_t *)src; char a[32], b[32];
    (void)arg;
    if (p->n_frags == 0 && p->pdepth == 0) { snprintf(buf, max, "n"); return 1; }
    ucc_memunits_to_str(p->threshold, a, sizeof(a)); ucc_memunits_to_str(p->frag_size, b, sizeof(b));
    snprintf(buf, max, "thresh=%s:fragsize=%s:nfrags=%u:pdepth=%u:%s", a, b, p->n_frags, p->pdepth, pipeline_order_names[p->order]);
    return 1;
}
int ucc_config_sscanf_table(const char *buf, void *dest, const void *arg) { (void)buf; (void)dest; (void)arg; return 1; }

/* ------------------------------------------------------------------ */
/* ucc.conf                                                            */
/* ------------------------------------------------------------------ */
typedef struct file_kv { char *key, *val; int section; } file_kv_t;
typedef struct file_section { char *name; char *vendor, *model; unsigned ts_lo, ts_hi, ppn_lo, ppn_hi, nn_lo, nn_hi, sock_lo, sock_hi; } file_section_t;
struct ucc_file_config { file_kv_t *kvs; unsigned n_kvs; file_section_t *sections; unsigned n_sections; };

static void parse_urange(const char *s, unsigned *lo, unsigned *hi)
{
    char *dash; unsigned long a = strtoul(s, &dash, 10);
    *lo = (unsigned)a; *hi = (unsigned)a;
    if (*dash == '-') { if (!strcasecmp(dash + 1, "inf")) *hi = UINT_MAX; else *hi = (unsigned)strtoul(dash + 1, NULL, 10); }
}

ucc_status_t ucc_parse_file_config(const char *filename, ucc_file_config_t **cfg_p)
{
    FILE *f = fopen(filename, "r");
    char line[4096];
    ucc_file_config_t *cfg;
    int cur = -1;
    if (!f) return UCC_ERR_NOT_FOUND;
    cfg = (ucc_file_config_t *)calloc(1, sizeof(*cfg));
    while (fgets(line, sizeof(line), f)) {
        char *p = line, *eq, *hash;
        hash = strchr(p, '#');
        /* '#' starts a comment only at line start or after whitespace (TUNE strings use '#') */
        if (hash && (hash == p || isspace((unsigned char)hash[-1]))) *hash = 0;
        ucc_str_trim(p);
        if (!*p || *p == ';') continue;
        if (*p == '[') { /* [name key=val key=val] */
            char *end = strchr(p, ']'); char **tk; unsigned n;
            if (!end) continue;
            *end = 0;
            tk = ucc_str_split(p + 1, " \t"); n = ucc_str_split_count(tk);
            cfg->sections = (file_section_t *)realloc(cfg->sections, (cfg->n_sections + 1) * sizeof(file_section_t));
            file_section_t *s = &cfg->sections[cfg->n_sections];
            memset(s, 0, sizeof(*s));
            s->ts_hi = s->ppn_hi = s->nn_hi = s->sock_hi = UINT_MAX;
            s->name = strdup(n ? tk[0] : "");
            for (unsigned i = 1; i < n; i++) {
                char *e = strchr(tk[i], '='); if (!e) continue; *e++ = 0;
                if (!strcasecmp(tk[i], "vendor")) s->vendor = strdup(e);
                else if (!strcasecmp(tk[i], "model")) s->model = strdup(e);
                else if (!strcasecmp(tk[i], "team_size")) parse_urange(e, &s->ts_lo, &s->ts_hi);
                else if (!strcasecmp(tk[i], "ppn")) parse_urange(e, &s->ppn_lo, &s->ppn_hi);
                else if (!strcasecmp(tk[i], "nnodes")) parse_urange(e, &s->nn_lo, &s->nn_hi);
                else if (!strcasecmp(tk[i], "sock")) parse_urange(e, &s->sock_lo, &s->sock_hi);
            }
            ucc_str_split_free(tk);
            cur = (int)cfg->n_sections++;
            continue;
        }
        eq = strchr(p, '=');
        if (!eq) continue;
        *eq++ = 0;
        ucc_str_trim(p); ucc_str_trim(eq);
        cfg->kvs = (file_kv_t *)realloc(cfg->kvs, (cfg->n_kvs + 1) * sizeof(file_kv_t));
        cfg->kvs[cfg->n_kvs].key = strdup(p); cfg->kvs[cfg->n_kvs].val = strdup(eq); cfg->kvs[cfg->n_kvs].section = cur;
        cfg->n_kvs++;
    }
    fclose(f);
    *cfg_p = cfg;
    return UCC_OK;
}

void ucc_release_file_config(ucc_file_config_t *cfg)
{
    if (!cfg) return;
    for (unsigned i = 0; i < cfg->n_kvs; i++) { free(cfg->kvs[i].key); free(cfg->kvs[i].val); }
    for (unsigned i = 0; i < cfg->n_sections; i++) { free(cfg->sections[i].name); free(cfg->sections[i].vendor); free(cfg->sections[i].model); }
    free(cfg->kvs); free(cfg->sections); free(cfg);
}

static int section_matches(const file_section_t *s, const ucc_file_section_filter_t *f)
{
    if (s->vendor && (!f->vendor || strcasecmp(s->vendor, f->vendor))) return 0;
    if (s->model && (!f->model || strcasecmp(s->model, f->model))) return 0;
    if (f->team_size < s->ts_lo || f->team_size > s->ts_hi) return 0;
    if (f->ppn < s->ppn_lo || f->ppn > s->ppn_hi) return 0;
    if (f->nnodes < s->nn_lo || f->nnodes > s->nn_hi) return 0;
    if (f->sock < s->sock_lo || f->sock > s->sock_hi) return 0;
    return 1;
}

const char *ucc_file_config_lookup(const ucc_file_config_t *cfg, const char *var, const ucc_file_section_filter_t *filter)
{
    const char *found = NULL;
    if (!cfg) return NULL;
    for (unsigned i = 0; i < cfg->n_kvs; i++) {
        if (strcmp(cfg->kvs[i].key, var)) continue;
        if (cfg->kvs[i].section < 0) { if (!filter) found = cfg->kvs[i].val; }
        else if (filter && section_matches(&cfg->sections[cfg->kvs[i].section], filter)) found = cfg->kvs[i].val;
    }
    return found;
}

/* ------------------------------------------------------------------ */
/* table operations                                                    */
/* ------------------------------------------------------------------ */
static int field_is_table(const ucc_config_field_t *f) { return f->parser.read == ucc_config_sscanf_table; }

static ucc_status_t set_defaults(void *opts, ucc_config_field_t *table)
{
    for (ucc_config_field_t *f = table; f->name; f++) {
        void *var = (char *)opts + f->offset;
        if (field_is_table(f)) { UCC_CHECK_RET(set_defaults(var, (ucc_config_field_t *)f->parser.arg)); continue; }
        if (!f->parser.read(f->dfl_value, var, f->parser.arg)) {
            ucc_error("config: invalid default '%s' for %s", f->dfl_value, f->name);
            return UCC_ERR_INVALID_PARAM;
        }
    }
    return UCC_OK;
}

static ucc_status_t set_field(void *opts, ucc_config_field_t *f, const char *value)
{
    void *var = (char *)opts + f->offset;
    char  saved[sizeof(ucc_pipeline_params_t) + 64];
    /* parse into a temporary first so a bad value keeps the old one */
    union { char b[128]; void *p; double d; } tmp;
    memset(&tmp, 0, sizeof(tmp));
    (void)saved;
    if (!f->parser.read(value, &tmp, f->parser.arg)) return UCC_ERR_INVALID_PARAM;
    f->parser.release(var, f->parser.arg);
    /* move tmp into place: sizes of all supported field types are <= 128 bytes */
    {
        size_t sz;
        if (f->parser.read == ucc_config_sscanf_string) sz = sizeof(char *);
        else if (f->parser.read == ucc_config_sscanf_array) sz = sizeof(ucc_config_names_array_t);
        else if (f->parser.read == ucc_config_sscanf_allow_list) sz = sizeof(ucc_config_allow_list_t);
        else if (f->parser.read == ucc_config_sscanf_names_list) sz = sizeof(ucc_config_names_list_t);
        else if (f->parser.read == ucc_config_sscanf_uint_ranged) {
            /* list head is self-referential: re-parse directly in place */
            ucc_config_release_uint_ranged(&tmp, NULL);
            return f->parser.read(value, var, f->parser.arg) ? UCC_OK : UCC_ERR_INVALID_PARAM;
        } else if (f->parser.read == ucc_config_sscanf_enum) sz = sizeof(unsigned);
        else { const char *h = (const char *)f->parser.arg; sz = (h && (h[0] == 'L' || h[0] == 'D')) ? 8 : (h && h[0] == 'P') ? sizeof(ucc_pipeline_params_t) : 4; }
        memcpy(var, &tmp, sz);
    }
    return UCC_OK;
}

static ucc_config_field_t *find_field(void *opts, ucc_config_field_t *table, const char *name, void **base)
{
    for (ucc_config_field_t *f = table; f->name; f++) {
        if (field_is_table(f)) {
            ucc_config_field_t *r = find_field((char *)opts + f->offset, (ucc_config_field_t *)f->parser.arg, name, base);
            if (r) return r;
        } else if (!strcmp(f->name, name)) { *base = opts; return f; }
    }
    return NULL;
}

ucc_status_t ucc_config_parser_set_value(void *opts, ucc_config_field_t *table, const char *name, const char *value)
{
    void *base = NULL;
    ucc_config_field_t *f = find_field(opts, table, name, &base);
    if (!f) return UCC_ERR_NOT_FOUND;
    return set_field(base, f, value);
}

ucc_status_t ucc_config_parser_get_value(void *opts, ucc_config_field_t *table, const char *name, char *value, size_t max)
{
    void *base = NULL;
    ucc_config_field_t *f = find_field(opts, table, name, &base);
    if (!f) return UCC_ERR_NOT_FOUND;
    f->parser.write(value, max, (char *)base + f->offset, f->parser.arg);
    return UCC_OK;
}

static ucc_status_t apply_source(void *opts, ucc_config_field_t *table, const char *prefix, int from_file,
                                 const ucc_file_section_filter_t *filter, int ignore_errors)
{
    char var[256];
    for (ucc_config_field_t *f = table; f->name; f++) {
        const char *val;
        if (field_is_table(f)) {
            UCC_CHECK_RET(apply_source((char *)opts + f->offset, (ucc_config_field_t *)f->parser.arg, prefix, from_file,
                                       filter, ignore_errors));
            continue;
        }
        if (!f->name[0]) continue;
        snprintf(var, sizeof(var), "%s%s", prefix, f->name);
        val = from_file ? ucc_file_config_lookup(ucc_global_config.file_cfg, var, filter) : getenv(var);
        if (!val) continue;
        if (set_field(opts, f, val) != UCC_OK) {
            if (ignore_errors) ucc_warn("config: invalid value '%s' for %s, keeping previous", val, var);
            else { ucc_error("config: invalid value '%s' for %s", val, var); return UCC_ERR_INVALID_PARAM; }
        }
    }
    return UCC_OK;
}

ucc_status_t ucc_config_parser_fill_opts_table(void *opts, ucc_config_field_t *table, const char *env_prefix,
                                               const char *table_prefix, int ignore_errors)
{
    char full[128];
    ucc_status_t st;
    const char *ep = env_prefix ? env_prefix : "UCC_";
    const char *tp = table_prefix ? table_prefix : "";
    st = set_defaults(opts, table);
    if (st != UCC_OK) return st;
    /* priority low->high: file(base prefix), file(full), env(base), env(full); then the same
     * with the user's env prefix (e.g. OMPI_UCC_) on top */
    for (int pass = 0; pass < 2; pass++) {
        const char *pfx = pass == 0 ? "UCC_" : ep;
        if (pass == 1 && !strcmp(ep, "UCC_")) break;
        for (int from_file = 1; from_file >= 0; from_file--) {
            if (from_file && !ucc_global_config.file_cfg) continue;
            if (tp[0]) {
                snprintf(full, sizeof(full), "%s", pfx);
                UCC_CHECK_RET(apply_source(opts, table, full, from_file, NULL, ignore_errors));
            }
            snprintf(full, sizeof(full), "%s%s", pfx, tp);
            UCC_CHECK_RET(apply_source(opts, table, full, from_file, NULL, ignore_errors));
        }
    }
    return UCC_OK;
}

ucc_status_t ucc_config_parser_fill_opts(void *opts, ucc_config_global_list_entry_t *entry, const char *env_prefix,
                                         int ignore_errors)
{
    ucc_config_table_register(entry);
    return ucc_config_parser_fill_opts_table(opts, entry->table, env_prefix, entry->prefix, ignore_errors);
}

ucc_status_t ucc_apply_file_cfg_sections(void *opts, ucc_config_field_t *table, const char *env_prefix,
                                         const char *table_prefix, const ucc_file_section_filter_t *filter)
{
    char full[128];
    if (!ucc_global_config.file_cfg) return UCC_OK;
    snprintf(full, sizeof(full), "%s%s", env_prefix ? env_prefix : "UCC_", table_prefix ? table_prefix : "");
    return apply_source(opts, table, full, 1, filter, 1);
}

void ucc_config_parser_release_opts(void *opts, ucc_config_field_t *table)
{
    for (ucc_config_field_t *f = table; f->name; f++) {
        void *var = (char *)opts + f->offset;
        if (field_is_table(f)) ucc_config_parser_release_opts(var, (ucc_config_field_t *)f->parser.arg);
        else f->parser.release(var, f->parser.arg);
    }
}

ucc_status_t ucc_config_parser_clone_opts(const void *src, void *dst, ucc_config_field_t *table)
{
    for (ucc_config_field_t *f = table; f->name; f++) {
        if (field_is_table(f)) {
            UCC_CHECK_RET(ucc_config_parser_clone_opts((const char *)src + f->offset, (char *)dst + f->offset,
                                                       (ucc_config_field_t *)f->parser.arg));
        } else UCC_CHECK_RET(f->parser.clone((const char *)src + f->offset, (char *)dst + f->offset, f->parser.arg));
    }
    return UCC_OK;
}

static void print_table(FILE *stream, const void *opts, ucc_config_field_t *table, const char *table_prefix,
                        const char *env_prefix, ucc_config_print_flags_t flags)
{
    char val[1024], help[512];
    for (ucc_config_field_t *f = table; f->name; f++) {
        if (field_is_table(f)) {
            print_table(stream, opts ? (const char *)opts + f->offset : NULL, (ucc_config_field_t *)f->parser.arg,
                        table_prefix, env_prefix, flags);
            continue;
        }
        if (!f->name[0]) continue;
        if (opts) f->parser.write(val, sizeof(val), (const char *)opts + f->offset, f->parser.arg);
        else snprintf(val, sizeof(val), "%s", f->dfl_value);
        if (flags & UCC_CONFIG_PRINT_DOC) {
            const char *d = f->doc ? f->doc : "";
            fprintf(stream, "#\n");
            while (*d) { const char *nl = strchr(d, '\n'); size_t l = nl ? (size_t)(nl - d) : strlen(d);
                fprintf(stream, "# %.*s\n", (int)l, d); d += l + (nl ? 1 : 0); }
            f->parser.help(help, sizeof(help), f->parser.arg);
            fprintf(stream, "#\n# syntax:    %s\n#\n", help);
        }
        fprintf(stream, "%s%s%s=%s\n", env_prefix, table_prefix, f->name, val);
        if (flags & UCC_CONFIG_PRINT_DOC) fprintf(stream, "\n");
    }
}

void ucc_config_parser_print_opts(FILE *stream, const char *title, const void *opts, ucc_config_field_t *table,
                                  const char *table_prefix, const char *env_prefix, ucc_config_print_flags_t flags)
{
    if (flags & UCC_CONFIG_PRINT_HEADER) fprintf(stream, "#\n# %s\n#\n\n", title ? title : "");
    if (flags & UCC_CONFIG_PRINT_CONFIG)
        print_table(stream, opts, table, table_prefix ? table_prefix : "", env_prefix ? env_prefix : "UCC_", flags);
}

void ucc_config_parser_print_all_opts(FILE *stream, const char *env_prefix, ucc_config_print_flags_t flags)
{
    ucc_config_global_list_entry_t *e;
    ucc_list_for_each(e, &ucc_config_global_list, list) {
        void *opts;
        if (!e->table || !e->table[0].name) continue;
        opts = calloc(1, e->size);
        if (ucc_config_parser_fill_opts_table(opts, e->table, env_prefix, e->prefix, 1) == UCC_OK) {
            ucc_config_parser_print_opts(stream, e->name, opts, e->table, e->prefix, env_prefix, flags);
            ucc_config_parser_release_opts(opts, e->table);
        }
        free(opts);
    }
}
