// Error reporting, launch counter, version.
#include <stdarg.h>

#include "common.cuh"

namespace rcn {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace rcn

extern "C" const char* rcn_last_error(void) { return rcn::g_err; }
extern "C" int rcn_version(void) { return 100; }
extern "C" unsigned long long rcn_launch_count(void) { return rcn::g_launches; }
