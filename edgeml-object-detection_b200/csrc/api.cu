// Error reporting and version of the C ABI (include/orie_b200.h).
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace orie {
static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace orie

extern "C" const char *orie_last_error(void) { return orie::g_error; }
extern "C" int orie_version(void) { return 100; }
extern "C" long long orie_launch_count(void) { return orie::g_launches.load(std::memory_order_relaxed); }
