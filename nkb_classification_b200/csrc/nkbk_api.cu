// Error string, ABI version and launch counter of libnkbk.so.
#include "nkbk_common.cuh"

#include <atomic>

namespace nkbk {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static thread_local int g_k1_overlap = 0;
static thread_local int g_heads_one_launch = 1;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool k1_overlap_previous() { return g_k1_overlap != 0; }
bool heads_one_launch() { return g_heads_one_launch != 0; }

}  // namespace nkbk

extern "C" int nkbk_abi_version(void) { return NKBK_ABI_VERSION; }
extern "C" const char* nkbk_last_error(void) { return nkbk::g_err; }
extern "C" int64_t nkbk_launch_count(void) { return nkbk::g_launches.load(std::memory_order_relaxed); }
extern "C" int nkbk_k1_overlap_previous(int enable) {
    const int prev = nkbk::g_k1_overlap;
    nkbk::g_k1_overlap = enable ? 1 : 0;
    return prev;
}
extern "C" int nkbk_heads_one_launch(int enable) {
    const int prev = nkbk::g_heads_one_launch;
    nkbk::g_heads_one_launch = enable ? 1 : 0;
    return prev;
}
