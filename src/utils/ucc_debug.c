/* Debug helpers: fatal-signal backtraces (UCC_HANDLE_ERRORS=bt), in the spirit of the
 * error handler UCS provides to the reference (UCX_HANDLE_ERRORS). */
#include "ucc_compiler_def.h"
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <unistd.h>

static void fatal_handler(int sig)
{
    void *bt[64];
    int   n = backtrace(bt, 64);
    char  msg[64];
    int   l = snprintf(msg, sizeof(msg), "==== ucc_b200: caught signal %d, backtrace ====\n", sig);
    if (write(2, msg, (size_t)l) < 0) {}
    backtrace_symbols_fd(bt, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
void ucc_debug_install_handlers(void)
{
    const char *e = getenv("UCC_HANDLE_ERRORS");
    if (!e || !strstr(e, "bt")) return;
    signal(SIGSEGV, fatal_handler); signal(SIGBUS, fatal_handler); signal(SIGABRT, fatal_handler); signal(SIGFPE, fatal_handler);
}
