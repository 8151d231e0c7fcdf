/* Basic compiler helpers shared by every translation unit of ucc_b200. */
#ifndef UCC_COMPILER_DEF_H_
#define UCC_COMPILER_DEF_H_

#include <ucc/api/ucc_status.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ucc_likely(x)   __builtin_expect(!!(x), 1)
#define ucc_unlikely(x) __builtin_expect(!!(x), 0)
#define ucc_offsetof(_t, _m) offsetof(_t, _m)
#define ucc_container_of(_ptr, _type, _member) \
    ((_type *)((char *)(_ptr) - offsetof(_type, _member)))
#define ucc_derived_of(_ptr, _type) ((_type *)(_ptr))
#define ucc_static_assert(_c) _Static_assert(_c, #_c)
#define UCC_PP_QUOTE_(x) #x
#define UCC_PP_QUOTE(x) UCC_PP_QUOTE_(x)
#define UCC_PP_CAT_(a, b) a##b
#define UCC_PP_CAT(a, b) UCC_PP_CAT_(a, b)
#define ucc_min(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a < _b ? _a : _b; })
#define ucc_max(a, b) ({ __typeof__(a) _a = (a); __typeof__(b) _b = (b); _a > _b ? _a : _b; })
#define ucc_array_size(a) (sizeof(a) / sizeof((a)[0]))
#define UCC_CACHE_LINE_SIZE 64
#define UCC_ALIGNED(n) __attribute__((aligned(n)))
#define UCC_EXPORT __attribute__((visibility("default")))
#define UCC_UNUSED __attribute__((unused))
#define UCC_CTOR __attribute__((constructor))
#define UCC_DTOR __attribute__((destructor))

#define UCC_CHECK_GOTO(_cmd, _label, _st) \
    do { _st = (_cmd); if (ucc_unlikely(_st != UCC_OK)) goto _label; } while (0)
#define UCC_CHECK_RET(_cmd) \
    do { ucc_status_t _s = (_cmd); if (ucc_unlikely(_s != UCC_OK)) return _s; } while (0)

static inline void *ucc_malloc(size_t s, const char *what) { (void)what; return malloc(s); }
static inline void *ucc_calloc(size_t n, size_t s, const char *what) { (void)what; return calloc(n, s); }
static inline void *ucc_realloc(void *p, size_t s, const char *what) { (void)what; return realloc(p, s); }
static inline void  ucc_free(void *p) { free(p); }
static inline int   ucc_posix_memalign(void **p, size_t a, size_t s, const char *what)
{ (void)what; return posix_memalign(p, a, s); }

typedef uint32_t ucc_rank_t;
#define UCC_RANK_INVALID ((ucc_rank_t)-1)
#define UCC_RANK_MAX     ((ucc_rank_t)-2)
typedef uint32_t ucc_score_t;
#define UCC_SCORE_MAX     ((ucc_score_t)0x7fffffff)
#define UCC_SCORE_MIN     0
#define UCC_SCORE_INVALID ((ucc_score_t)-1)
#define UCC_MSG_MAX       ((size_t)-1)

#endif
