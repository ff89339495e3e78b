// Error slot, version and launch accounting of the C-ABI (include/gg_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace gg {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace gg

extern "C" {

int gg_version(void) { return 100; }

const char* gg_last_error(void) { return gg::g_error; }

int64_t gg_launch_count(void) { return gg::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
