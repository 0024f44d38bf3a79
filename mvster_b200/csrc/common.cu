// Error string, version and launch counter of libmvster_b200.
#include "common.cuh"
#include <atomic>
#include <string.h>

namespace mvster {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
}  // namespace mvster

extern "C" {
int mvster_version(void) { return 100; }  // 0.1.0
const char* mvster_last_error(void) { return mvster::g_err; }
uint64_t mvster_launch_count(void) { return mvster::g_launches.load(std::memory_order_relaxed); }
}
