// Error plumbing + version for the C ABI.
#include "common.cuh"
#include "../../include/phi3_b200.h"
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";

void p3_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* p3_last_error(void) { return g_err; }
extern "C" int p3_version(void) { return 1; }
