/* Logging: 12 levels, per-component level, optional file sink with rotation.
 * Role of reference utils/debug/log.c + utils/ucc_log.h (UCS free). */
#ifndef UCC_LOG_H_
#define UCC_LOG_H_
#include "ucc_compiler_def.h"
#include <stdarg.h>
#include <stdio.h>

typedef enum {
    UCC_LOG_LEVEL_FATAL, UCC_LOG_LEVEL_ERROR, UCC_LOG_LEVEL_WARN, UCC_LOG_LEVEL_DIAG,
    UCC_LOG_LEVEL_INFO, UCC_LOG_LEVEL_DEBUG, UCC_LOG_LEVEL_TRACE, UCC_LOG_LEVEL_TRACE_REQ,
    UCC_LOG_LEVEL_TRACE_DATA, UCC_LOG_LEVEL_TRACE_ASYNC, UCC_LOG_LEVEL_TRACE_FUNC,
    UCC_LOG_LEVEL_TRACE_POLL, UCC_LOG_LEVEL_LAST, UCC_LOG_LEVEL_PRINT
} ucc_log_level_t;

typedef struct ucc_log_component_config {
    ucc_log_level_t log_level;
    char            name[24];
} ucc_log_component_config_t;

extern const char *ucc_log_level_names[];

void ucc_log_init(void);          /* reads global opts: file, size, rotate */
void ucc_log_cleanup(void);
void ucc_log_dispatch(const char *file, unsigned line, const char *func, ucc_log_level_t level,
                      const ucc_log_component_config_t *comp, const char *fmt, ...)
    __attribute__((format(printf, 6, 7)));
const char *ucc_get_host_name(void);

#define ucc_log_component_is_enabled(_lvl, _comp) (ucc_unlikely((_lvl) <= (_comp)->log_level))
#define ucc_log_component(_lvl, _comp, _fmt, ...) \
    do { if (ucc_log_component_is_enabled(_lvl, _comp)) \
        ucc_log_dispatch(__FILE__, __LINE__, __func__, (_lvl), (_comp), _fmt, ##__VA_ARGS__); } while (0)

/* core ("UCC") component logging */
extern ucc_log_component_config_t ucc_global_log_component;
#define ucc_log_core(_lvl, _fmt, ...) ucc_log_component(_lvl, &ucc_global_log_component, _fmt, ##__VA_ARGS__)
#define ucc_error(_f, ...) ucc_log_core(UCC_LOG_LEVEL_ERROR, _f, ##__VA_ARGS__)
#define ucc_warn(_f, ...)  ucc_log_core(UCC_LOG_LEVEL_WARN, _f, ##__VA_ARGS__)
#define ucc_diag(_f, ...)  ucc_log_core(UCC_LOG_LEVEL_DIAG, _f, ##__VA_ARGS__)
#define ucc_info(_f, ...)  ucc_log_core(UCC_LOG_LEVEL_INFO, _f, ##__VA_ARGS__)
#define ucc_debug(_f, ...) ucc_log_core(UCC_LOG_LEVEL_DEBUG, _f, ##__VA_ARGS__)
#define ucc_trace(_f, ...) ucc_log_core(UCC_LOG_LEVEL_TRACE, _f, ##__VA_ARGS__)
#define ucc_trace_req(_f, ...)  ucc_log_core(UCC_LOG_LEVEL_TRACE_REQ, _f, ##__VA_ARGS__)
#define ucc_trace_poll(_f, ...) ucc_log_core(UCC_LOG_LEVEL_TRACE_POLL, _f, ##__VA_ARGS__)
#define ucc_fatal(_f, ...) do { ucc_log_dispatch(__FILE__, __LINE__, __func__, UCC_LOG_LEVEL_FATAL, \
        &ucc_global_log_component, _f, ##__VA_ARGS__); abort(); } while (0)
#define ucc_print(_f, ...) ucc_log_dispatch(__FILE__, __LINE__, __func__, UCC_LOG_LEVEL_PRINT, \
        &ucc_global_log_component, _f, ##__VA_ARGS__)

#ifdef UCC_ENABLE_ASSERT
#define ucc_assert(_c) do { if (ucc_unlikely(!(_c))) ucc_fatal("assertion failed: %s", #_c); } while (0)
#else
#define ucc_assert(_c) do { (void)sizeof(_c); } while (0)
#endif
#define ucc_assert_always(_c) do { if (ucc_unlikely(!(_c))) ucc_fatal("assertion failed: %s", #_c); } while (0)
#endif
