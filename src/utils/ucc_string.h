/* String helpers: split/join, memunits, case-insensitive find, safe concat. */
#ifndef UCC_STRING_H_
#define UCC_STRING_H_
#include "ucc_compiler_def.h"
char       **ucc_str_split(const char *str, const char *delim); /* NULL-terminated array, free with ucc_str_split_free */
unsigned     ucc_str_split_count(char **split);
void         ucc_str_split_free(char **split);
ucc_status_t ucc_str_is_number(const char *str);
ucc_status_t ucc_str_to_memunits(const char *buf, size_t *dest); /* 128, 4K, 1Mb, inf, auto(-2) */
void         ucc_memunits_to_str(size_t value, char *buf, size_t max);
ucc_status_t ucc_str_memunits_range_to_ulong(const char *str, size_t *start, size_t *end); /* "a-b" */
const char  *ucc_strstr_last(const char *s, const char *pattern);
ucc_status_t ucc_str_concat(const char *a, const char *b, char **out);
void         ucc_str_trim(char *s);
int          ucc_str_find_in_list(const char *s, const char **list); /* case-insensitive; -1 if absent */
unsigned long ucc_str_hash_djb2(const char *s);
#define UCC_MEMUNITS_INF  ((size_t)-1)
#define UCC_MEMUNITS_AUTO ((size_t)-2)
#define UCC_ULUNITS_AUTO  ((unsigned long)-2)
#define UCC_UUNITS_AUTO   ((unsigned)-2)
#endif
