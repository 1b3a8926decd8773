// Error string, ABI version and launch counter of libnkbk.so.
#include "nkbk_common.cuh"

#include <atomic>

namespace nkbk {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace nkbk

extern "C" int nkbk_abi_version(void) { return NKBK_ABI_VERSION; }
extern "C" const char* nkbk_last_error(void) { return nkbk::g_err; }
extern "C" int64_t nkbk_launch_count(void) { return nkbk::g_launches.load(std::memory_order_relaxed); }
