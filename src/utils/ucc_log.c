#include "ucc_log.h"
#include "ucc_atomic.h"
#include "ucc_time.h"
#include "core/ucc_global_opts.h"
#include <unistd.h>
#include <sys/syscall.h>
#include <sys/stat.h>

const char *ucc_log_level_names[] = {
    "FATAL", "ERROR", "WARN", "DIAG", "INFO", "DEBUG", "TRACE", "REQ", "DATA", "ASYNC", "FUNC", "POLL", NULL,
    "PRINT", NULL};

ucc_log_component_config_t ucc_global_log_component = {UCC_LOG_LEVEL_WARN, "UCC"};

static FILE          *log_stream    = NULL;
static char           log_path[512] = "";
static size_t         log_written   = 0;
static ucc_spinlock_t log_lock      = {0};
static char           host_name[128] = "";

const char *ucc_get_host_name(void)
{
    if (!host_name[0]) {
        if (gethostname(host_name, sizeof(host_name) - 1) != 0) strcpy(host_name, "localhost");
        host_name[sizeof(host_name) - 1] = 0;
    }
    return host_name;
}

/* expand %p -> pid, %h -> host in a file name template */
static void expand_name(const char *tmpl, char *out, size_t max)
{
    size_t o = 0;
    for (const char *p = tmpl; *p && o + 1 < max; p++) {
        if (p[0] == '%' && p[1] == 'p') { o += snprintf(out + o, max - o, "%d", (int)getpid()); p++; }
        else if (p[0] == '%' && p[1] == 'h') { o += snprintf(out + o, max - o, "%s", ucc_get_host_name()); p++; }
        else out[o++] = *p;
    }
    out[o < max ? o : max - 1] = 0;
}

void ucc_log_init(void)
{
    const char *f = ucc_global_config.log_file;
    if (log_stream && log_stream != stdout && log_stream != stderr) fclose(log_stream);
    log_stream = NULL;
    if (f && f[0]) {
        if (!strcmp(f, "stdout")) log_stream = stdout;
        else if (!strcmp(f, "stderr")) log_stream = stderr;
        else {
            expand_name(f, log_path, sizeof(log_path));
            log_stream = fopen(log_path, "a");
            if (!log_stream) { log_path[0] = 0; }
        }
    }
    if (!log_stream) log_stream = stdout;
    if (ucc_global_config.log_buffer_size > 0 && log_stream != stdout && log_stream != stderr)
        setvbuf(log_stream, NULL, _IOFBF, ucc_global_config.log_buffer_size);
    log_written = 0;
}

void ucc_log_cleanup(void)
{
    if (log_stream && log_stream != stdout && log_stream != stderr) fclose(log_stream);
    log_stream = NULL;
}

static void log_rotate(void)
{
    char from[600], to[600];
    int  n = (int)ucc_global_config.log_file_rotate;
    fclose(log_stream);
    for (int i = n - 1; i >= 0; i--) {
        if (i == 0) snprintf(from, sizeof(from), "%s", log_path);
        else snprintf(from, sizeof(from), "%s.%d", log_path, i);
        snprintf(to, sizeof(to), "%s.%d", log_path, i + 1);
        if (i + 1 > n) remove(from); else rename(from, to);
    }
    if (n == 0) remove(log_path);
    log_stream  = fopen(log_path, "a");
    if (!log_stream) log_stream = stdout;
    log_written = 0;
}

void ucc_log_dispatch(const char *file, unsigned line, const char *func, ucc_log_level_t level,
                      const ucc_log_component_config_t *comp, const char *fmt, ...)
{
    char        buf[2048];
    va_list     ap;
    const char *base = strrchr(file, '/');
    double      t    = ucc_get_wall_time();
    FILE       *out;
    int         n;
    (void)func;
    if (level != UCC_LOG_LEVEL_PRINT && level != UCC_LOG_LEVEL_FATAL && !ucc_global_config.log_print_enable &&
        level > UCC_LOG_LEVEL_WARN && 0) return;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    base = base ? base + 1 : file;
    ucc_spin_lock(&log_lock);
    out = log_stream ? log_stream : stdout;
    if (level == UCC_LOG_LEVEL_PRINT) n = fprintf(out, "%s\n", buf);
    else
        n = fprintf(out, "[%.6f] [%s:%d:%ld] %16s:%-4u %-8s %-5s %s\n", t, ucc_get_host_name(), (int)getpid(),
                    (long)syscall(SYS_gettid) - (long)getpid(), base, line, comp->name, ucc_log_level_names[level], buf);
    if (level <= UCC_LOG_LEVEL_WARN || ucc_global_config.log_buffer_size == 0) fflush(out);
    if (n > 0) log_written += (size_t)n;
    if (log_path[0] && ucc_global_config.log_file_size != (size_t)-1 && log_written >= ucc_global_config.log_file_size)
        log_rotate();
    ucc_spin_unlock(&log_lock);
    if (level <= ucc_global_config.log_level_trigger && level != UCC_LOG_LEVEL_PRINT &&
        ucc_global_config.log_level_trigger != UCC_LOG_LEVEL_FATAL) {
        fflush(out);
        abort(); /* LOG_LEVEL_TRIGGER: stop at the first message of that severity */
    }
}
