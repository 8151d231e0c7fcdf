/* coll_score implementation: see ucc_coll_score.h. */
#include "ucc_coll_score.h"
#include "utils/ucc_string.h"
#include "utils/ucc_log.h"
#include "components/base/ucc_base_iface.h"
#include "core/ucc_team.h"
#include <ctype.h>
#include <dlfcn.h>
#include <stdio.h>
#include <strings.h>

struct ucc_score_map {
    ucc_coll_score_t *score;
};

/* ------------------------------------------------------------------ */
/* range helpers                                                       */
/* ------------------------------------------------------------------ */
static ucc_coll_entry_t *entry_new(ucc_score_t score, ucc_base_coll_init_fn_t init, ucc_base_team_t *team)
{
    ucc_coll_entry_t *e = (ucc_coll_entry_t *)malloc(sizeof(*e));
    if (e) { e->score = score; e->init = init; e->team = team; }
    return e;
}
static ucc_msg_range_t *range_new(size_t start, size_t end, ucc_score_t score, ucc_base_coll_init_fn_t init, ucc_base_team_t *team)
{
    ucc_msg_range_t *r = (ucc_msg_range_t *)malloc(sizeof(*r));
    if (!r) return NULL;
    r->start = start; r->end = end; r->super.score = score; r->super.init = init; r->super.team = team;
    ucc_list_head_init(&r->fallback);
    return r;
}
static void range_free(ucc_msg_range_t *r)
{
    ucc_coll_entry_t *e, *t;
    ucc_list_for_each_safe(e, t, &r->fallback, list_elem) { ucc_list_del(&e->list_elem); free(e); }
    free(r);
}
/* insert into a fallback list keeping descending score order, skipping exact duplicates */
static ucc_status_t fb_insert(ucc_list_link_t *fb, ucc_score_t score, ucc_base_coll_init_fn_t init, ucc_base_team_t *team)
{
    ucc_coll_entry_t *e, *n;
    ucc_list_for_each(e, fb, list_elem) if (e->init == init && e->team == team) return UCC_OK;
    n = entry_new(score, init, team);
    if (!n) return UCC_ERR_NO_MEMORY;
    ucc_list_for_each(e, fb, list_elem) if (e->score < score) { ucc_list_insert_before(&e->list_elem, &n->list_elem); return UCC_OK; }
    ucc_list_add_tail(fb, &n->list_elem);
    return UCC_OK;
}
static ucc_status_t fb_copy(ucc_list_link_t *dst, const ucc_list_link_t *src)
{
    ucc_coll_entry_t *e;
    ucc_list_for_each(e, (ucc_list_link_t *)src, list_elem) UCC_CHECK_RET(fb_insert(dst, e->score, e->init, e->team));
    return UCC_OK;
}
static int fb_equal(const ucc_list_link_t *a, const ucc_list_link_t *b)
{
    const ucc_list_link_t *x = a->next, *y = b->next;
    while (x != a && y != b) {
        const ucc_coll_entry_t *ex = ucc_container_of(x, ucc_coll_entry_t, list_elem), *ey = ucc_container_of(y, ucc_coll_entry_t, list_elem);
        if (ex->init != ey->init || ex->team != ey->team || ex->score != ey->score) return 0;
        x = x->next; y = y->next;
    }
    return x == a && y == b;
}
static ucc_msg_range_t *range_dup(const ucc_msg_range_t *r, size_t start, size_t end)
{
    ucc_msg_range_t *n = range_new(start, end, r->super.score, r->super.init, r->super.team);
    if (n && fb_copy(&n->fallback, &r->fallback) != UCC_OK) { range_free(n); return NULL; }
    return n;
}
static void list_free(ucc_list_link_t *l)
{
    ucc_msg_range_t *r, *t;
    ucc_list_for_each_safe(r, t, l, super.list_elem) { ucc_list_del(&r->super.list_elem); range_free(r); }
}
/* drop score-0 ranges and glue neighbours that are indistinguishable */
static void list_normalize(ucc_list_link_t *l)
{
    ucc_msg_range_t *r, *t;
    ucc_list_for_each_safe(r, t, l, super.list_elem) if (r->super.score == 0 || r->start >= r->end) { ucc_list_del(&r->super.list_elem); range_free(r); }
    ucc_list_for_each_safe(r, t, l, super.list_elem) {
        if (&t->super.list_elem == l) break;
        if (r->end == t->start && r->super.score == t->super.score && r->super.init == t->super.init &&
            r->super.team == t->super.team && fb_equal(&r->fallback, &t->fallback)) {
            t->start = r->start;
            ucc_list_del(&r->super.list_elem); range_free(r);
        }
    }
}
static const ucc_msg_range_t *list_find(const ucc_list_link_t *l, size_t point)
{
    const ucc_msg_range_t *r;
    ucc_list_for_each(r, (ucc_list_link_t *)l, super.list_elem) if (point >= r->start && (point < r->end || (r->end == UCC_MSG_MAX && point == UCC_MSG_MAX))) return r;
    return NULL;
}
/* sorted, de-duplicated boundaries of two lists */
static size_t *collect_bounds(const ucc_list_link_t *a, const ucc_list_link_t *b, unsigned *n_out)
{
    unsigned n = 0, cap = 2 * (unsigned)(ucc_list_length(a) + ucc_list_length(b)) + 2;
    size_t *v = (size_t *)malloc(cap * sizeof(size_t));
    const ucc_msg_range_t *r;
    if (!v) return NULL;
    ucc_list_for_each(r, (ucc_list_link_t *)a, super.list_elem) { v[n++] = r->start; v[n++] = r->end; }
    ucc_list_for_each(r, (ucc_list_link_t *)b, super.list_elem) { v[n++] = r->start; v[n++] = r->end; }
    for (unsigned i = 1; i < n; i++) { size_t x = v[i]; unsigned j = i; while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; j--; } v[j] = x; }
    unsigned m = 0;
    for (unsigned i = 0; i < n; i++) if (m == 0 || v[m - 1] != v[i]) v[m++] = v[i];
    *n_out = m;
    return v;
}

/* ------------------------------------------------------------------ */
/* alloc / add / dup / set                                             */
/* ------------------------------------------------------------------ */
ucc_status_t ucc_coll_score_alloc(ucc_coll_score_t **score)
{
    ucc_coll_score_t *s = (ucc_coll_score_t *)malloc(sizeof(*s));
    if (!s) return UCC_ERR_NO_MEMORY;
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) for (int j = 0; j < UCC_MEMORY_TYPE_LAST; j++) ucc_list_head_init(&s->scores[i][j]);
    *score = s;
    return UCC_OK;
}
void ucc_coll_score_free(ucc_coll_score_t *s)
{
    if (!s) return;
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) for (int j = 0; j < UCC_MEMORY_TYPE_LAST; j++) list_free(&s->scores[i][j]);
    free(s);
}

static ucc_status_t list_merge(const ucc_list_link_t *a, const ucc_list_link_t *b, ucc_list_link_t *out);

ucc_status_t ucc_coll_score_add_range(ucc_coll_score_t *score, ucc_coll_type_t coll_type, ucc_memory_type_t mem_type,
                                      size_t start, size_t end, ucc_score_t msg_score, ucc_base_coll_init_fn_t init,
                                      ucc_base_team_t *team)
{
    ucc_list_link_t *l, tmp, res;
    ucc_msg_range_t *r;
    ucc_status_t st;
    if (start >= end) return UCC_ERR_INVALID_PARAM;
    if (msg_score == 0) return UCC_OK;
    l = &score->scores[ucc_coll_type_index(coll_type)][mem_type];
    r = range_new(start, end, msg_score, init, team);
    if (!r) return UCC_ERR_NO_MEMORY;
    /* adding = merging a one-range list: overlaps resolve to the higher score */
    ucc_list_head_init(&tmp); ucc_list_head_init(&res);
    ucc_list_add_tail(&tmp, &r->super.list_elem);
    st = list_merge(l, &tmp, &res);
    list_free(&tmp);
    if (st != UCC_OK) { list_free(&res); return st; }
    list_free(l);
    ucc_list_splice_tail(l, &res);
    return UCC_OK;
}

ucc_status_t ucc_coll_score_dup(const ucc_coll_score_t *in, ucc_coll_score_t **out)
{
    ucc_coll_score_t *s;
    const ucc_msg_range_t *r;
    UCC_CHECK_RET(ucc_coll_score_alloc(&s));
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) for (int j = 0; j < UCC_MEMORY_TYPE_LAST; j++)
        ucc_list_for_each(r, (ucc_list_link_t *)&in->scores[i][j], super.list_elem) {
            ucc_msg_range_t *n = range_dup(r, r->start, r->end);
            if (!n) { ucc_coll_score_free(s); return UCC_ERR_NO_MEMORY; }
            ucc_list_add_tail(&s->scores[i][j], &n->super.list_elem);
        }
    *out = s;
    return UCC_OK;
}

void ucc_coll_score_set(ucc_coll_score_t *score, ucc_score_t value)
{
    ucc_msg_range_t *r;
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) for (int j = 0; j < UCC_MEMORY_TYPE_LAST; j++)
        ucc_list_for_each(r, &score->scores[i][j], super.list_elem) r->super.score = value;
}

ucc_status_t ucc_coll_score_build_default(ucc_base_team_t *team, ucc_score_t default_score, ucc_base_coll_init_fn_t default_init,
                                          uint64_t coll_types, ucc_memory_type_t *mem_types, int mt_n, ucc_coll_score_t **score_p)
{
    ucc_coll_score_t *s;
    ucc_memory_type_t all[UCC_MEMORY_TYPE_LAST];
    UCC_CHECK_RET(ucc_coll_score_alloc(&s));
    if (!mem_types) { for (int j = 0; j < UCC_MEMORY_TYPE_LAST; j++) all[j] = (ucc_memory_type_t)j; mem_types = all; mt_n = UCC_MEMORY_TYPE_LAST; }
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) {
        if (!(coll_types & UCC_BIT(i))) continue;
        for (int j = 0; j < mt_n; j++) {
            ucc_status_t st = ucc_coll_score_add_range(s, (ucc_coll_type_t)UCC_BIT(i), mem_types[j], 0, UCC_MSG_MAX, default_score, default_init, team);
            if (st != UCC_OK) { ucc_coll_score_free(s); return st; }
        }
    }
    *score_p = s;
    return UCC_OK;
}

/* ------------------------------------------------------------------ */
/* merge                                                               */
/* ------------------------------------------------------------------ */
static ucc_status_t list_merge(const ucc_list_link_t *a, const ucc_list_link_t *b, ucc_list_link_t *out)
{
    unsigned nb;
    size_t *bounds = collect_bounds(a, b, &nb);
    if (!bounds) return UCC_ERR_NO_MEMORY;
    for (unsigned i = 0; i + 1 < nb; i++) {
        size_t lo = bounds[i], hi = bounds[i + 1];
        const ucc_msg_range_t *ra = list_find(a, lo), *rb = list_find(b, lo), *win, *lose;
        ucc_msg_range_t *n;
        if (!ra && !rb) continue;
        win = ra; lose = rb;
        if (!ra || (rb && rb->super.score > ra->super.score)) { win = rb; lose = ra; }
        n = range_dup(win, lo, hi);
        if (!n) { free(bounds); return UCC_ERR_NO_MEMORY; }
        if (lose && !(lose->super.init == win->super.init && lose->super.team == win->super.team)) {
            if (fb_insert(&n->fallback, lose->super.score, lose->super.init, lose->super.team) != UCC_OK ||
                fb_copy(&n->fallback, &lose->fallback) != UCC_OK) { range_free(n); free(bounds); return UCC_ERR_NO_MEMORY; }
        }
        ucc_list_add_tail(out, &n->super.list_elem);
    }
    free(bounds);
    list_normalize(out);
    return UCC_OK;
}

ucc_status_t ucc_coll_score_merge(ucc_coll_score_t *s1, ucc_coll_score_t *s2, ucc_coll_score_t **rst, int free_inputs)
{
    ucc_coll_score_t *out;
    ucc_status_t st = ucc_coll_score_alloc(&out);
    if (st != UCC_OK) goto done;
    for (int i = 0; i < UCC_COLL_TYPE_NUM && st == UCC_OK; i++)
        for (int j = 0; j < UCC_MEMORY_TYPE_LAST && st == UCC_OK; j++)
            st = list_merge(&s1->scores[i][j], &s2->scores[i][j], &out->scores[i][j]);
    if (st != UCC_OK) { ucc_coll_score_free(out); out = NULL; }
    *rst = out;
done:
    if (free_inputs) { ucc_coll_score_free(s1); ucc_coll_score_free(s2); }
    return st;
}

ucc_status_t ucc_coll_score_merge_in(ucc_coll_score_t **dst, ucc_coll_score_t *src)
{
    ucc_coll_score_t *out = NULL;
    ucc_status_t st = ucc_coll_score_merge(*dst, src, &out, 1);
    *dst = out;
    return st;
}

/* ------------------------------------------------------------------ */
/* update: user/TUNE overrides                                         */
/* ------------------------------------------------------------------ */
static ucc_status_t list_update(ucc_list_link_t *dest, const ucc_list_link_t *src, ucc_score_t default_score)
{
    unsigned nb;
    size_t *bounds;
    ucc_list_link_t out;
    if (ucc_list_is_empty(src)) return UCC_OK;
    bounds = collect_bounds(dest, src, &nb);
    if (!bounds) return UCC_ERR_NO_MEMORY;
    ucc_list_head_init(&out);
    for (unsigned i = 0; i + 1 < nb; i++) {
        size_t lo = bounds[i], hi = bounds[i + 1];
        const ucc_msg_range_t *rd = list_find(dest, lo), *rs = list_find(src, lo);
        ucc_msg_range_t *n = NULL;
        if (rd && !rs) n = range_dup(rd, lo, hi);
        else if (!rd && rs) {
            if (!rs->super.init) continue; /* score-only token over a hole: nothing to re-score */
            n = range_dup(rs, lo, hi);
            if (n && n->super.score == UCC_SCORE_INVALID) n->super.score = default_score;
        } else if (rd && rs) {
            n = range_dup(rd, lo, hi);
            if (n) {
                if (rs->super.score != UCC_SCORE_INVALID) n->super.score = rs->super.score;
                if (rs->super.init && (rs->super.init != rd->super.init || rs->super.team != rd->super.team)) {
                    /* the user's choice replaces the init fn; the previous one stays reachable as a fallback */
                    if (fb_insert(&n->fallback, rd->super.score, rd->super.init, rd->super.team) != UCC_OK) { range_free(n); n = NULL; }
                    else { n->super.init = rs->super.init; n->super.team = rs->super.team; }
                }
            }
        } else continue;
        if (!n) { free(bounds); list_free(&out); return UCC_ERR_NO_MEMORY; }
        ucc_list_add_tail(&out, &n->super.list_elem);
    }
    free(bounds);
    list_normalize(&out);
    list_free(dest);
    ucc_list_splice_tail(dest, &out);
    return UCC_OK;
}

ucc_status_t ucc_coll_score_update(ucc_coll_score_t *score, ucc_coll_score_t *update, ucc_score_t default_score,
                                   ucc_memory_type_t *mtypes, int mt_n, uint64_t colls)
{
    if (mt_n == 0 || !mtypes) mt_n = UCC_MEMORY_TYPE_LAST;
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) {
        if (!(colls & UCC_BIT(i))) continue;
        for (int j = 0; j < mt_n; j++) {
            ucc_memory_type_t mt = mtypes ? mtypes[j] : (ucc_memory_type_t)j;
            UCC_CHECK_RET(list_update(&score->scores[i][mt], &update->scores[i][mt], default_score));
        }
    }
    return UCC_OK;
}

/* ------------------------------------------------------------------ */
/* TUNE string parser                                                  */
/*   token[#token...]; token = qualifier[:qualifier...]                 */
/*   qualifiers: coll list | a-b[,c-d] msg ranges | mem types |         */
/*               [a-b,...] team sizes | score (uint|inf) | @alg         */
/* ------------------------------------------------------------------ */
typedef struct tune_token {
    uint64_t colls; uint32_t mtypes;
    size_t (*ranges)[2]; unsigned n_ranges;
    ucc_score_t score; int has_alg; int alg_id; char alg_str[64];
    int team_size_ok;
} tune_token_t;

static int parse_colls(const char *s, uint64_t *colls)
{
    char **t = ucc_str_split(s, ","); unsigned n = ucc_str_split_count(t); uint64_t m = 0; int ok = n > 0;
    for (unsigned i = 0; i < n && ok; i++) { ucc_coll_type_t c = ucc_coll_type_from_str(t[i]); if (c == UCC_COLL_TYPE_LAST) ok = 0; else m |= (uint64_t)c; }
    ucc_str_split_free(t);
    if (ok) *colls = m;
    return ok;
}
static int parse_mtypes(const char *s, uint32_t *mtypes)
{
    char **t = ucc_str_split(s, ","); unsigned n = ucc_str_split_count(t); uint32_t m = 0; int ok = n > 0;
    for (unsigned i = 0; i < n && ok; i++) { ucc_memory_type_t mt = ucc_mem_type_from_str(t[i]); if (mt == UCC_MEMORY_TYPE_LAST) ok = 0; else m |= 1u << mt; }
    ucc_str_split_free(t);
    if (ok) *mtypes = m;
    return ok;
}
static int parse_ranges(const char *s, tune_token_t *tk)
{
    char **t = ucc_str_split(s, ","); unsigned n = ucc_str_split_count(t); int ok = n > 0;
    size_t (*r)[2] = (size_t(*)[2])malloc(sizeof(size_t[2]) * (n ? n : 1));
    for (unsigned i = 0; i < n && ok; i++) {
        ok = ucc_str_memunits_range_to_ulong(t[i], &r[i][0], &r[i][1]) == UCC_OK;
        if (ok && r[i][1] != UCC_MSG_MAX && r[i][1] == r[i][0]) ok = 0;
    }
    ucc_str_split_free(t);
    if (ok) { tk->ranges = r; tk->n_ranges = n; } else free(r);
    return ok;
}
static int parse_team_sizes(const char *s, ucc_rank_t team_size, int *match)
{
    size_t len = strlen(s); char buf[128]; char **t; unsigned n; int ok = 1, m = 0;
    if (len < 3 || s[0] != '[' || s[len - 1] != ']' || len >= sizeof(buf)) return 0;
    memcpy(buf, s + 1, len - 2); buf[len - 2] = 0;
    t = ucc_str_split(buf, ","); n = ucc_str_split_count(t);
    for (unsigned i = 0; i < n && ok; i++) {
        size_t a, b;
        if (strchr(t[i], '-')) ok = ucc_str_memunits_range_to_ulong(t[i], &a, &b) == UCC_OK;
        else { ok = ucc_str_to_memunits(t[i], &a) == UCC_OK; b = a; }
        if (ok && team_size >= a && team_size <= b) m = 1;
    }
    ucc_str_split_free(t);
    *match = m;
    return ok && n > 0;
}
static int parse_score(const char *s, ucc_score_t *score)
{
    char *e; unsigned long v;
    if (!strcasecmp(s, UCC_SCORE_MAX_STR)) { *score = UCC_SCORE_MAX; return 1; }
    if (!isdigit((unsigned char)s[0])) return 0;
    v = strtoul(s, &e, 10);
    if (*e || v > UCC_SCORE_MAX) return 0;
    *score = (ucc_score_t)v;
    return 1;
}

static ucc_status_t parse_token(const char *str, ucc_rank_t team_size, tune_token_t *tk)
{
    char **q = ucc_str_split(str, ":");
    unsigned n = ucc_str_split_count(q);
    ucc_status_t st = UCC_OK;
    memset(tk, 0, sizeof(*tk));
    tk->score = UCC_SCORE_INVALID; tk->team_size_ok = 1;
    for (unsigned i = 0; i < n; i++) {
        const char *s = q[i];
        int m;
        if (s[0] == '@') {
            if (tk->has_alg || !s[1]) { st = UCC_ERR_INVALID_PARAM; break; }
            tk->has_alg = 1;
            if (ucc_str_is_number(s + 1) == UCC_OK) { tk->alg_id = atoi(s + 1); tk->alg_str[0] = 0; }
            else { tk->alg_id = -1; snprintf(tk->alg_str, sizeof(tk->alg_str), "%s", s + 1); }
        } else if (s[0] == '[') {
            if (!parse_team_sizes(s, team_size, &m)) { st = UCC_ERR_INVALID_PARAM; break; }
            tk->team_size_ok = m;
        } else if (parse_score(s, &tk->score)) {
        } else if (!tk->colls && parse_colls(s, &tk->colls)) {
        } else if (!tk->mtypes && parse_mtypes(s, &tk->mtypes)) {
        } else if (!tk->ranges && parse_ranges(s, tk)) {
        } else { st = UCC_ERR_INVALID_PARAM; break; }
    }
    ucc_str_split_free(q);
    if (st == UCC_OK && tk->score == UCC_SCORE_INVALID && !tk->has_alg) st = UCC_ERR_INVALID_PARAM;
    if (st != UCC_OK) { free(tk->ranges); tk->ranges = NULL; }
    return st;
}

/* token entries are *appended*; later tokens override earlier ones (applied as successive updates) */
static ucc_status_t score_from_token(const tune_token_t *tk, ucc_coll_score_t *score, ucc_base_coll_init_fn_t dflt_init,
                                     ucc_base_team_t *team, ucc_alg_id_to_init_fn_t alg_fn)
{
    size_t full[1][2] = {{0, UCC_MSG_MAX}};
    size_t (*ranges)[2] = tk->n_ranges ? tk->ranges : full;
    unsigned n_ranges = tk->n_ranges ? tk->n_ranges : 1;
    uint64_t colls = tk->colls ? tk->colls : UCC_COLL_TYPE_ALL;
    uint32_t mtypes = tk->mtypes ? tk->mtypes : UCC_MEM_TYPE_MASK_FULL;
    (void)dflt_init;
    for (int c = 0; c < UCC_COLL_TYPE_NUM; c++) {
        if (!(colls & UCC_BIT(c))) continue;
        for (int m = 0; m < UCC_MEMORY_TYPE_LAST; m++) {
            ucc_base_coll_init_fn_t init = NULL;
            if (!(mtypes & (1u << m))) continue;
            if (tk->has_alg) {
                ucc_status_t st;
                if (!alg_fn) { ucc_warn("tune: this component does not support algorithm selection"); return UCC_ERR_NOT_SUPPORTED; }
                st = alg_fn(tk->alg_id, tk->alg_str[0] ? tk->alg_str : NULL, (ucc_coll_type_t)UCC_BIT(c), (ucc_memory_type_t)m, &init);
                if (st == UCC_ERR_INVALID_PARAM) return st; /* unknown alg name/id for an explicitly listed coll */
                if (st != UCC_OK) { if (tk->colls) return st; continue; /* "all colls" token: skip those without such alg */ }
            }
            for (unsigned r = 0; r < n_ranges; r++) {
                ucc_msg_range_t *nr = range_new(ranges[r][0], ranges[r][1], tk->score, init, init ? team : NULL), *it;
                ucc_list_link_t *l = &score->scores[c][m], upd;
                if (!nr) return UCC_ERR_NO_MEMORY;
                /* keep the partial list non-overlapping: apply as an update that may carry INVALID/NULL */
                ucc_list_head_init(&upd); ucc_list_add_tail(&upd, &nr->super.list_elem);
                if (ucc_list_is_empty(l)) { ucc_list_del(&nr->super.list_elem); ucc_list_add_tail(l, &nr->super.list_elem); continue; }
                {
                    /* partial x partial: overlay nr on top of l, interval by interval */
                    unsigned nb; size_t *b = collect_bounds(l, &upd, &nb); ucc_list_link_t out; ucc_list_head_init(&out);
                    if (!b) { range_free(nr); return UCC_ERR_NO_MEMORY; }
                    for (unsigned i = 0; i + 1 < nb; i++) {
                        const ucc_msg_range_t *ro = list_find(l, b[i]), *rn = list_find(&upd, b[i]); ucc_msg_range_t *x;
                        if (!ro && !rn) continue;
                        x = range_dup(ro ? ro : rn, b[i], b[i + 1]);
                        if (!x) { free(b); return UCC_ERR_NO_MEMORY; }
                        if (ro && rn) {
                            if (rn->super.score != UCC_SCORE_INVALID) x->super.score = rn->super.score;
                            if (rn->super.init) { x->super.init = rn->super.init; x->super.team = rn->super.team; }
                        }
                        ucc_list_add_tail(&out, &x->super.list_elem);
                    }
                    free(b);
                    /* do NOT normalize away score 0 here: 0 means "disable" and must survive until update */
                    list_free(l); ucc_list_splice_tail(l, &out);
                    ucc_list_for_each(it, l, super.list_elem) { (void)it; }
                }
                list_free(&upd);
            }
        }
    }
    return UCC_OK;
}

ucc_status_t ucc_coll_score_alloc_from_str(const char *str, ucc_coll_score_t **score_p, ucc_rank_t team_size,
                                           ucc_base_coll_init_fn_t init, ucc_base_team_t *team, ucc_alg_id_to_init_fn_t alg_fn)
{
    ucc_coll_score_t *score;
    char **tokens;
    unsigned n;
    ucc_status_t st;
    UCC_CHECK_RET(ucc_coll_score_alloc(&score));
    tokens = ucc_str_split(str, "#");
    n = ucc_str_split_count(tokens);
    st = UCC_OK;
    for (unsigned i = 0; i < n && st == UCC_OK; i++) {
        tune_token_t tk;
        st = parse_token(tokens[i], team_size, &tk);
        if (st != UCC_OK) { ucc_error("failed to parse tune token \"%s\"", tokens[i]); break; }
        if (tk.team_size_ok) st = score_from_token(&tk, score, init, team, alg_fn);
        free(tk.ranges);
    }
    ucc_str_split_free(tokens);
    if (st != UCC_OK) { ucc_coll_score_free(score); return st; }
    *score_p = score;
    return UCC_OK;
}

ucc_status_t ucc_coll_score_update_from_str(const char *str, const ucc_coll_score_team_info_t *info, ucc_base_team_t *team,
                                            ucc_coll_score_t *score)
{
    ucc_coll_score_t *upd;
    ucc_status_t st = ucc_coll_score_alloc_from_str(str, &upd, info->size, info->init, team, info->alg_fn);
    if (st != UCC_OK) return st;
    st = ucc_coll_score_update(score, upd, info->default_score, info->supported_mem_types, info->num_mem_types, info->supported_colls);
    ucc_coll_score_free(upd);
    return st;
}

/* ------------------------------------------------------------------ */
/* map                                                                 */
/* ------------------------------------------------------------------ */
ucc_status_t ucc_coll_score_build_map(ucc_coll_score_t *score, ucc_score_map_t **map_p)
{
    ucc_score_map_t *map = (ucc_score_map_t *)malloc(sizeof(*map));
    if (!map) return UCC_ERR_NO_MEMORY;
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) for (int j = 0; j < UCC_MEMORY_TYPE_LAST; j++) list_normalize(&score->scores[i][j]);
    map->score = score;
    *map_p = map;
    return UCC_OK;
}
void ucc_coll_score_free_map(ucc_score_map_t *map) { if (!map) return; ucc_coll_score_free(map->score); free(map); }

static ucc_status_t map_lookup_as(ucc_score_map_t *map, ucc_base_coll_args_t *bargs, ucc_rank_t rank, ucc_rank_t size, ucc_msg_range_t **range)
{
    ucc_coll_args_t *a = &bargs->args;
    ucc_memory_type_t mt = ucc_coll_args_mem_type(a, rank);
    size_t msgsize = ucc_coll_args_msgsize(a, rank, size);
    const ucc_msg_range_t *r;
    int ct = ucc_coll_type_index(a->coll_type);
    if (mt >= UCC_MEMORY_TYPE_LAST || ct >= UCC_COLL_TYPE_NUM) return UCC_ERR_NOT_SUPPORTED;
    r = list_find(&map->score->scores[ct][mt], msgsize);
    if (!r) return UCC_ERR_NOT_SUPPORTED;
    *range = (ucc_msg_range_t *)r;
    return UCC_OK;
}

ucc_status_t ucc_coll_score_map_lookup(ucc_score_map_t *map, ucc_base_coll_args_t *bargs, ucc_msg_range_t **range)
{
    return map_lookup_as(map, bargs, bargs->team ? ucc_team_rank_(bargs->team) : 0, bargs->team ? ucc_team_size_(bargs->team) : 1, range);
}

ucc_status_t ucc_coll_init(ucc_score_map_t *map, ucc_base_coll_args_t *bargs, ucc_coll_task_t **task)
{
    return ucc_coll_init_as(map, bargs, bargs->team ? ucc_team_rank_(bargs->team) : 0, bargs->team ? ucc_team_size_(bargs->team) : 1, task);
}

/* selection for a collective that runs on a sub-group (cl/hier): root / sizes are relative to that group */
ucc_status_t ucc_coll_init_as(ucc_score_map_t *map, ucc_base_coll_args_t *bargs, ucc_rank_t rank, ucc_rank_t size, ucc_coll_task_t **task)
{
    ucc_msg_range_t *r;
    ucc_coll_entry_t *fb;
    ucc_status_t st = map_lookup_as(map, bargs, rank, size, &r);
    if (st != UCC_OK) return st;
    st = r->super.init(bargs, r->super.team, task);
    if (st == UCC_OK && !(*task)->init_fn && (*task)->team == r->super.team) (*task)->init_fn = r->super.init; /* innermost (TL) entry wins */
    if (st != UCC_ERR_NOT_SUPPORTED && st != UCC_ERR_NOT_IMPLEMENTED) return st;
    ucc_list_for_each(fb, &r->fallback, list_elem) {
        ucc_debug("coll_init: falling back to the next candidate (score %u)", fb->score);
        st = fb->init(bargs, fb->team, task);
        if (st == UCC_OK && !(*task)->init_fn && (*task)->team == fb->team) (*task)->init_fn = fb->init;
        if (st != UCC_ERR_NOT_SUPPORTED && st != UCC_ERR_NOT_IMPLEMENTED) return st;
    }
    return st;
}

static ucc_coll_score_name_fn_t name_resolver = NULL;
void ucc_coll_score_set_name_resolver(ucc_coll_score_name_fn_t fn) { __atomic_store_n(&name_resolver, fn, __ATOMIC_RELAXED); /* every context sets the same function */ }

static const char *init_name(ucc_base_coll_init_fn_t init, ucc_base_team_t *team, char *tmp, size_t max)
{
    Dl_info info;
    const char *n = name_resolver ? name_resolver(init, team) : NULL;
    if (n) return n;
    if (dladdr((void *)(uintptr_t)init, &info) && info.dli_sname) return info.dli_sname;
    snprintf(tmp, max, "%p", (void *)(uintptr_t)init);
    return tmp;
}

void ucc_coll_score_map_str(const ucc_score_map_t *map, char *buf, size_t len)
{
    size_t o = 0;
    char a[32], b[32], t1[32];
    buf[0] = 0;
    for (int i = 0; i < UCC_COLL_TYPE_NUM; i++) for (int j = 0; j < UCC_MEMORY_TYPE_LAST; j++) {
        const ucc_list_link_t *l = &map->score->scores[i][j];
        const ucc_msg_range_t *r;
        if (ucc_list_is_empty(l)) continue;
        o += snprintf(buf + o, o < len ? len - o : 0, "%s %s:", ucc_coll_type_str((ucc_coll_type_t)UCC_BIT(i)), ucc_mem_type_str((ucc_memory_type_t)j));
        ucc_list_for_each(r, (ucc_list_link_t *)l, super.list_elem) {
            const ucc_coll_entry_t *fb;
            ucc_memunits_to_str(r->start, a, sizeof(a)); ucc_memunits_to_str(r->end, b, sizeof(b));
            if (o < len) o += snprintf(buf + o, len - o, " {%s..%s}:%s:%u:%s", a, b, r->super.team ? ucc_base_team_name(r->super.team) : "?",
                          r->super.score, init_name(r->super.init, r->super.team, t1, sizeof(t1)));
            ucc_list_for_each(fb, (ucc_list_link_t *)&r->fallback, list_elem)
                if (o < len) o += snprintf(buf + o, len - o, " >%s:%u", fb->team ? ucc_base_team_name(fb->team) : "?", fb->score);
        }
        if (o < len) o += snprintf(buf + o, len - o, "\n");
    }
}

void ucc_coll_score_map_print_info(const ucc_score_map_t *map, int verbosity)
{
    size_t len = 1 << 16;
    char *buf = (char *)malloc(len), *line, *save = NULL;
    if (!buf) return;
    ucc_coll_score_map_str(map, buf, len);
    for (line = strtok_r(buf, "\n", &save); line; line = strtok_r(NULL, "\n", &save))
        ucc_log_core((ucc_log_level_t)verbosity, "%s", line);
    free(buf);
}
