// Library-level entry points of libmaskplanner_b200.so: version, error reporting, device queries.
#include <stdarg.h>

#include "common.cuh"

namespace mpb {

static thread_local char g_err[512] = "no error";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace mpb

extern "C" int mpb_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char *mpb_last_error_string(void) { return mpb::g_err; }

extern "C" int mpb_device_sm_count(void)
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MPB_ERR_CUDA;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return MPB_ERR_CUDA;
    return n;
}

extern "C" int mpb_device_arch(void)
{
    int dev = 0, maj = 0, min = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MPB_ERR_CUDA;
    if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return MPB_ERR_CUDA;
    if (cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return MPB_ERR_CUDA;
    return maj * 10 + min;
}
