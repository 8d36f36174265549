// capi.cu -- error plumbing and version of the C ABI (include/b2second.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void b2s_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *b2s_last_error(void) { return g_err; }
extern "C" int b2s_version(void) { return B2S_VERSION; }
