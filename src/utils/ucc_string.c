#include "ucc_string.h"
#include <ctype.h>
#include <stdio.h>
#include <strings.h>

char **ucc_str_split(const char *str, const char *delim)
{
    char  *copy = strdup(str), *save = NULL, *tok;
    size_t cap = 8, n = 0;
    char **out = (char **)malloc(cap * sizeof(char *));
    if (!copy || !out) { free(copy); free(out); return NULL; }
    for (tok = strtok_r(copy, delim, &save); tok; tok = strtok_r(NULL, delim, &save)) {
        if (n + 2 > cap) { cap *= 2; out = (char **)realloc(out, cap * sizeof(char *)); }
        out[n++] = strdup(tok);
    }
    out[n] = NULL;
    free(copy);
    return out;
}
unsigned ucc_str_split_count(char **s) { unsigned n = 0; while (s && s[n]) n++; return n; }
void ucc_str_split_free(char **s) { if (!s) return; for (char **p = s; *p; p++) free(*p); free(s); }

ucc_status_t ucc_str_is_number(const char *str)
{
    if (!str || !*str) return UCC_ERR_INVALID_PARAM;
    for (; *str; str++) if (!isdigit((unsigned char)*str)) return UCC_ERR_INVALID_PARAM;
    return UCC_OK;
}

ucc_status_t ucc_str_to_memunits(const char *buf, size_t *dest)
{
    char  *end;
    double v;
    size_t mul = 1;
    while (isspace((unsigned char)*buf)) buf++;
    if (!strcasecmp(buf, "inf")) { *dest = UCC_MEMUNITS_INF; return UCC_OK; }
    if (!strcasecmp(buf, "auto")) { *dest = UCC_MEMUNITS_AUTO; return UCC_OK; }
    v = strtod(buf, &end);
    if (end == buf || v < 0) return UCC_ERR_INVALID_PARAM;
    switch (toupper((unsigned char)*end)) {
    case 'K': mul = 1ul << 10; end++; break;
    case 'M': mul = 1ul << 20; end++; break;
    case 'G': mul = 1ul << 30; end++; break;
    case 'T': mul = 1ul << 40; end++; break;
    default: break;
    }
    if (toupper((unsigned char)*end) == 'B') end++;
    while (isspace((unsigned char)*end)) end++;
    if (*end) return UCC_ERR_INVALID_PARAM;
    *dest = (size_t)(v * (double)mul + 0.5);
    return UCC_OK;
}

void ucc_memunits_to_str(size_t value, char *buf, size_t max)
{
    static const char *suf[] = {"", "K", "M", "G", "T"};
    int i = 0;
    if (value == UCC_MEMUNITS_INF) { snprintf(buf, max, "inf"); return; }
    if (value == UCC_MEMUNITS_AUTO) { snprintf(buf, max, "auto"); return; }
    while (i < 4 && value >= 1024 && (value % 1024) == 0) { value /= 1024; i++; }
    snprintf(buf, max, "%zu%s", value, suf[i]);
}

ucc_status_t ucc_str_memunits_range_to_ulong(const char *str, size_t *start, size_t *end)
{
    char  tmp[128];
    char *dash;
    if (strlen(str) >= sizeof(tmp)) return UCC_ERR_INVALID_PARAM;
    strcpy(tmp, str);
    dash = strchr(tmp, '-');
    if (!dash) return UCC_ERR_INVALID_PARAM;
    *dash = 0;
    if (ucc_str_to_memunits(tmp, start) != UCC_OK || ucc_str_to_memunits(dash + 1, end) != UCC_OK)
        return UCC_ERR_INVALID_PARAM;
    if (*start > *end) return UCC_ERR_INVALID_PARAM;
    return UCC_OK;
}

const char *ucc_strstr_last(const char *s, const char *pattern)
{
    const char *found = NULL, *p = s;
    while ((p = strstr(p, pattern)) != NULL) { found = p; p++; }
    return found;
}

ucc_status_t ucc_str_concat(const char *a, const char *b, char **out)
{
    size_t la = strlen(a), lb = strlen(b);
    char  *r = (char *)malloc(la + lb + 1);
    if (!r) return UCC_ERR_NO_MEMORY;
    memcpy(r, a, la); memcpy(r + la, b, lb + 1);
    *out = r;
    return UCC_OK;
}

void ucc_str_trim(char *s)
{
    char *p = s, *e;
    while (isspace((unsigned char)*p)) p++;
    if (p != s) memmove(s, p, strlen(p) + 1);
    e = s + strlen(s);
    while (e > s && isspace((unsigned char)e[-1])) *--e = 0;
}

int ucc_str_find_in_list(const char *s, const char **list)
{
    for (int i = 0; list[i]; i++) if (!strcasecmp(s, list[i])) return i;
    return -1;
}

unsigned long ucc_str_hash_djb2(const char *s)
{
    unsigned long h = 5381;
    int c;
    while ((c = (unsigned char)*s++)) h = ((h << 5) + h) + (unsigned long)c;
    return h;
}
