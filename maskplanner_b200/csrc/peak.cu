// Measurement helper: the FP32 SIMT roof of this device, measured, for the kernels of the path that are
// bound by the fma pipe (ball query / kNN / chamfer nearest-neighbour search; SURVEY.md 8d, BASELINE.md 2:
// "must be measured by the builder").  Two variants: scalar FFMA (one fma per lane per instruction) and packed
// FFMA2 (fma.rn.f32x2, two per lane per instruction -- what chamfer_nn_kernel issues).  Each thread keeps 16
// independent accumulator chains so the pipe, not the dependency latency, is the limit; the result is
// stored so nothing is optimised away.  bench.py times the launch with CUDA events:
//   flops = 2 * threads * iters * 16 (* 2 for the packed variant).
#include "common.cuh"

namespace mpb {

constexpr int kPeakThreads = 256;
constexpr int kPeakChains = 16;

template <bool PACKED>
__global__ void __launch_bounds__(kPeakThreads)
peak_ffma_kernel(int iters, float a, float b, float *__restrict__ out)
{
    if (PACKED) {
        float2 acc[kPeakChains];
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) acc[i] = make_float2((float)(threadIdx.x + i), (float)(blockIdx.x - i));
        const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kPeakChains; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) s += acc[i].x + acc[i].y;
        out[(size_t)blockIdx.x * kPeakThreads + threadIdx.x] = s;
    } else {
        float acc[kPeakChains];
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) acc[i] = (float)(threadIdx.x + i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < kPeakChains; ++i) acc[i] = fmaf(acc[i], a, b);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) s += acc[i];
        out[(size_t)blockIdx.x * kPeakThreads + threadIdx.x] = s;
    }
}

}  // namespace mpb

extern "C" int64_t mpb_peak_fp32_threads(int ctas_per_sm)
{
    return (int64_t)mpb::sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 8) * mpb::kPeakThreads;
}

extern "C" int mpb_peak_fp32_ffma(int packed, int iters, int ctas_per_sm, float *out, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(out && iters > 0, "bad argument");
    const int grid = sm_count() * (ctas_per_sm > 0 ? ctas_per_sm : 8);
    if (packed)
        peak_ffma_kernel<true><<<grid, kPeakThreads, 0, (cudaStream_t)stream>>>(iters, 0.999f, 0.001f, out);
    else
        peak_ffma_kernel<false><<<grid, kPeakThreads, 0, (cudaStream_t)stream>>>(iters, 0.999f, 0.001f, out);
    return check_launch("peak_ffma_kernel");
}
