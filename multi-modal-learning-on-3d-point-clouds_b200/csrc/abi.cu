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

cudaError_t scratch_alloc(void **ptr, size_t bytes, cudaStream_t stream) {
    static std::atomic<unsigned long long> configured{0};  // bit per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_relaxed) & bit)) {
        cudaMemPool_t pool;
        e = cudaDeviceGetDefaultMemPool(&pool, dev);
        if (e != cudaSuccess) return e;
        unsigned long long keep = ~0ull;
        e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        configured.fetch_or(bit, std::memory_order_relaxed);
    }
    return cudaMallocAsync(ptr, bytes, stream);
}

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
