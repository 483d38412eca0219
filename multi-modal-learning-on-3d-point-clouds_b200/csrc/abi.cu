// abi.cu -- status / error plumbing of the C ABI (include/pn2_abi.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace pn2 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            cached = sms;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace pn2

extern "C" const char *pn2_last_error(void) { return pn2::g_err; }
extern "C" int pn2_abi_version(void) { return PN2_ABI_VERSION; }
extern "C" uint64_t pn2_launch_count(void) { return pn2::g_launches.load(std::memory_order_relaxed); }
