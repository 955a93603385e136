// Error reporting and version of the spikesort_b200 C ABI.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>

namespace {
thread_local char g_err[512] = "";
}

void ss_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int ss_version(void) { return SS_VERSION; }
extern "C" const char *ss_last_error(void) { return g_err; }
