// C-ABI plumbing shared by all entry points: thread-local error text, launch accounting.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace mmu {
namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error((int)e, "%s: launch failed: %s", what, cudaGetErrorString(e));
    return MMU_OK;
}
}  // namespace mmu

extern "C" int mmu_version(void) { return MMU_VERSION; }
extern "C" const char *mmu_last_error(void) { return mmu::g_err; }
extern "C" uint64_t mmu_launch_count(void) { return mmu::g_launches.load(std::memory_order_relaxed); }
