// error.cu — thread-local error string + version (no exceptions cross the C ABI)
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void tmla_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {
int tmla_version(void) { return TMLA_VERSION; }
const char *tmla_last_error(void) { return g_err; }
}
