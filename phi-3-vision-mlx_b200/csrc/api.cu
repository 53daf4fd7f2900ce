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

// ---- in-kernel timeline (tools/chain_trace.py; off unless a buffer is set): every traced launch takes the next slot of
// P3_TRACE_CTAS x 8 u64 and its CTAs' thread 0 store {globaltimer at start, after the dependency wait, first operands ready,
// main loop done, exit, smid, kind}. Diagnostic only; costs one null test per launch when off.
static unsigned long long* g_trace = nullptr;
static int g_trace_slots = 0, g_trace_next = 0;
extern "C" int p3_trace_set(void* buf, int n_slots) { g_trace = (unsigned long long*)buf; g_trace_slots = n_slots; g_trace_next = 0; return 0; }
extern "C" int p3_trace_count(void) { return g_trace_next; }
unsigned long long* p3_trace_slot() {
    if (!g_trace || g_trace_next >= g_trace_slots) return nullptr;
    return g_trace + (size_t)(g_trace_next++) * P3_TRACE_CTAS * 8;
}
