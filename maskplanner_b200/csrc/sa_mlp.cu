// BatchNorm / ReLU / max-pool stages of PointNetSetAbstraction's shared MLP on bf16 activations (sm_100a).
// Reference: models/pointnet2_utils.py:210-214 -- per layer `relu(bn(conv(x)))`, then `max` over the K
// neighbours of each group.  The 1x1 conv runs in sa_gemm.cu and leaves Z[M,C] (bf16) in HBM; everything
// here is HBM-bound streaming over [M,C] with 16-byte vector accesses (8 bf16 per thread per row):
//
//   colstats      : per-channel sum / sum of squares of Z (training-mode batch statistics)
//   bn_finalize   : -> scale = gamma*rstd, shift = beta - mean*scale; running-stat update (momentum, unbiased var)
//   bn_relu       : A = relu(scale*Z + shift)                                   (bf16 -> bf16)
//   bn_relu_max   : out[g,c] = max_k relu(scale*Z[g,k,c] + shift) + arg-max     (last layer, fused max-pool)
//   bwd_stats     : sum dY, sum dY*zhat with dY = dA * [scale*Z+shift > 0]      (dense or pooled upstream)
//   bwd_finalize  : dgamma, dbeta and the two per-channel means BN backward needs
//   bwd_apply     : dZ = gamma*rstd*(dY - mean(dY) - zhat*mean(dY*zhat))        (bf16 out)
//
// Batch statistics are accumulated in fp32 per CTA and combined in fp64 by the finalize kernels
// (deterministic: fixed partial order, no atomics across CTAs).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mpb {

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8])
{
    const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(p[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8])
{
    uint4 v;
    __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

constexpr int kEwThreads = 256;

// Generic "two per-channel sums over all rows" skeleton: each thread owns one 8-channel group and walks
// rows with stride; CTA-level combine through shared-memory atomics; one partial row per CTA.
template <class RowFn>
__device__ __forceinline__ void column_sums(int64_t M, int C, float *__restrict__ partials, RowFn fn)
{
    extern __shared__ float sm_acc[];  // [2][C]
    const int cg = C >> 3;
    const int rows_per_pass = kEwThreads / cg;
    const int tr = threadIdx.x / cg, tc = threadIdx.x - tr * cg;
    for (int i = threadIdx.x; i < 2 * C; i += kEwThreads) sm_acc[i] = 0.f;
    __syncthreads();
    if (tr < rows_per_pass) {
        float s0[8], s1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
        for (int64_t r = (int64_t)blockIdx.x * rows_per_pass + tr; r < M; r += (int64_t)gridDim.x * rows_per_pass) fn(r, tc * 8, s0, s1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&sm_acc[tc * 8 + i], s0[i]);
            atomicAdd(&sm_acc[C + tc * 8 + i], s1[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kEwThreads) partials[(size_t)blockIdx.x * 2 * C + i] = sm_acc[i];
}

__global__ void __launch_bounds__(kEwThreads)
colstats_kernel(const __nv_bfloat16 *__restrict__ Z, int64_t M, int C, float *__restrict__ partials)
{
    column_sums(M, C, partials, [&](int64_t r, int c0, float(&s0)[8], float(&s1)[8]) {
        float z[8];
        unpack8(*reinterpret_cast<const uint4 *>(Z + r * C + c0), z);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s0[i] += z[i];
            s1[i] = fmaf(z[i], z[i], s1[i]);
        }
    });
}

// Combine the per-CTA partial sums of 8 channels (one CTA = 8 channels x 32 partial lanes) in fp64, fixed order.
// Returns true on the thread that holds the totals of channel `c`.
constexpr int kFinThreads = 256;
__device__ __forceinline__ bool reduce_partials(const float *__restrict__ partials, int nparts, int C, int &c, double &s, double &q)
{
    __shared__ double red[2][32][8];
    const int pl = threadIdx.x >> 3, ci = threadIdx.x & 7;
    c = blockIdx.x * 8 + ci;
    s = 0.0, q = 0.0;
    if (c < C)
        for (int p = pl; p < nparts; p += 32) {
            s += (double)partials[(size_t)p * 2 * C + c];
            q += (double)partials[(size_t)p * 2 * C + C + c];
        }
    red[0][pl][ci] = s;
    red[1][pl][ci] = q;
    __syncthreads();
    if (pl != 0 || c >= C) return false;
    s = 0.0, q = 0.0;
    for (int p = 0; p < 32; ++p) s += red[0][p][ci], q += red[1][p][ci];
    return true;
}

// Training: batch statistics -> scale/shift/mean/rstd, running stats updated in place.
// (conv bias only moves the mean: BN(z + b) == BN(z); it enters the running mean, reference :210-212.)
__global__ void __launch_bounds__(kFinThreads)
bn_finalize_kernel(const float *__restrict__ partials, int nparts, int C, int C_valid, double M,
                                   const float *__restrict__ bias, const float *__restrict__ gamma, const float *__restrict__ beta,
                                   float *__restrict__ running_mean, float *__restrict__ running_var, float momentum, float eps,
                                   float *__restrict__ scale, float *__restrict__ shift, float *__restrict__ mean_out,
                                   float *__restrict__ rstd_out)
{
    int c;
    double s, q;
    if (reduce_partials(partials, nparts, C, c, s, q)) {
        if (c >= C_valid) {  // zero-padded channel (GEMM alignment): contributes exact zeros downstream
            scale[c] = shift[c] = mean_out[c] = rstd_out[c] = 0.f;
            return;
        }
        const double mean = s / M;
        double var = q / M - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
        const float sc = g * rstd;
        scale[c] = sc;
        shift[c] = b - (float)mean * sc;
        mean_out[c] = (float)mean;
        rstd_out[c] = rstd;
        if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * ((float)mean + (bias ? bias[c] : 0.f));
        if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(M > 1.0 ? var * M / (M - 1.0) : var);
    }
}

__global__ void __launch_bounds__(kEwThreads)
bn_relu_kernel(const __nv_bfloat16 *__restrict__ Z, const float *__restrict__ scale, const float *__restrict__ shift, int64_t total_vec,
               int cg, __nv_bfloat16 *__restrict__ A)
{
    for (int64_t v = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; v < total_vec; v += (int64_t)gridDim.x * kEwThreads) {
        const int c0 = (int)(v % cg) * 8;
        float z[8];
        unpack8(reinterpret_cast<const uint4 *>(Z)[v], z);
        const float4 s0 = *reinterpret_cast<const float4 *>(scale + c0), s1 = *reinterpret_cast<const float4 *>(scale + c0 + 4);
        const float4 t0 = *reinterpret_cast<const float4 *>(shift + c0), t1 = *reinterpret_cast<const float4 *>(shift + c0 + 4);
        const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = fmaxf(fmaf(z[i], sc[i], sh[i]), 0.f);
        reinterpret_cast<uint4 *>(A)[v] = pack8(z);
    }
}

// out[g, c] = max_k relu(scale*Z[g*K + k, c] + shift); arg[g, c] = first k attaining it (torch.max semantics).
__global__ void __launch_bounds__(kEwThreads)
bn_relu_max_kernel(const __nv_bfloat16 *__restrict__ Z, const float *__restrict__ scale, const float *__restrict__ shift, int64_t G,
                   int K, int C, float *__restrict__ out, int *__restrict__ arg)
{
    const int cg = C >> 3;
    for (int64_t v = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; v < G * cg; v += (int64_t)gridDim.x * kEwThreads) {
        const int64_t g = v / cg;
        const int c0 = (int)(v - g * cg) * 8;
        float sc[8], sh[8], best[8];
        int bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) sc[i] = scale[c0 + i], sh[i] = shift[c0 + i], best[i] = -INFINITY, bi[i] = 0;
        const __nv_bfloat16 *zp = Z + (g * K) * C + c0;
        for (int k = 0; k < K; ++k) {
            float z[8];
            unpack8(*reinterpret_cast<const uint4 *>(zp + (int64_t)k * C), z);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = fmaxf(fmaf(z[i], sc[i], sh[i]), 0.f);
                if (a > best[i]) best[i] = a, bi[i] = k;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) out[g * C + c0 + i] = best[i], arg[g * C + c0 + i] = bi[i];
    }
}

__global__ void __launch_bounds__(kEwThreads)
bwd_stats_dense_kernel(const __nv_bfloat16 *__restrict__ dA, const __nv_bfloat16 *__restrict__ Z, const float *__restrict__ scale,
                       const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd, int64_t M,
                       int C, float *__restrict__ partials)
{
    column_sums(M, C, partials, [&](int64_t r, int c0, float(&s0)[8], float(&s1)[8]) {
        float z[8], d[8];
        unpack8(*reinterpret_cast<const uint4 *>(Z + r * C + c0), z);
        unpack8(*reinterpret_cast<const uint4 *>(dA + r * C + c0), d);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float dy = fmaf(z[i], scale[c0 + i], shift[c0 + i]) > 0.f ? d[i] : 0.f;
            s0[i] += dy;
            s1[i] = fmaf(dy, (z[i] - mean[c0 + i]) * rstd[c0 + i], s1[i]);
        }
    });
}

// Upstream gradient is the pooled one: only the arg-max row of each (group, channel) carries dOut.
__global__ void __launch_bounds__(kEwThreads)
bwd_stats_pooled_kernel(const float *__restrict__ dOut, const int *__restrict__ arg, const __nv_bfloat16 *__restrict__ Z,
                        const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                        const float *__restrict__ rstd, int64_t G, int K, int C, float *__restrict__ partials)
{
    column_sums(G, C, partials, [&](int64_t g, int c0, float(&s0)[8], float(&s1)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            const float z = __bfloat162float(Z[(g * K + arg[g * C + c]) * C + c]);
            const float dy = fmaf(z, scale[c], shift[c]) > 0.f ? dOut[g * C + c] : 0.f;
            s0[i] += dy;
            s1[i] = fmaf(dy, (z - mean[c]) * rstd[c], s1[i]);
        }
    });
}

// dgamma = sum dY*zhat, dbeta = sum dY; coef[0][c] = gamma*rstd, coef[1][c] = mean(dY), coef[2][c] = mean(dY*zhat)
__global__ void __launch_bounds__(kFinThreads)
bwd_finalize_kernel(const float *__restrict__ partials, int nparts, int C, int C_valid, double M,
                                    const float *__restrict__ gamma, const float *__restrict__ rstd, float *__restrict__ dgamma,
                                    float *__restrict__ dbeta, float *__restrict__ coef)
{
    int c;
    double s, q;
    if (reduce_partials(partials, nparts, C, c, s, q)) {
        if (c >= C_valid) {
            coef[c] = coef[C + c] = coef[2 * C + c] = 0.f;
            return;
        }
        if (dbeta) dbeta[c] = (float)s;
        if (dgamma) dgamma[c] = (float)q;
        coef[c] = (gamma ? gamma[c] : 1.f) * rstd[c];
        coef[C + c] = (float)(s / M);
        coef[2 * C + c] = (float)(q / M);
    }
}

template <bool POOLED>
__global__ void __launch_bounds__(kEwThreads)
bwd_apply_kernel(const __nv_bfloat16 *__restrict__ dA, const float *__restrict__ dOut, const int *__restrict__ arg, int K,
                 const __nv_bfloat16 *__restrict__ Z, const float *__restrict__ scale, const float *__restrict__ shift,
                 const float *__restrict__ mean, const float *__restrict__ rstd, const float *__restrict__ coef, int64_t M, int C,
                 __nv_bfloat16 *__restrict__ dZ)
{
    const int cg = C >> 3;
    for (int64_t v = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; v < M * cg; v += (int64_t)gridDim.x * kEwThreads) {
        const int64_t r = v / cg;
        const int c0 = (int)(v - r * cg) * 8;
        float z[8], d[8];
        unpack8(reinterpret_cast<const uint4 *>(Z)[v], z);
        if (POOLED) {
            const int64_t g = r / K;
            const int k = (int)(r - g * K);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = arg[g * C + c0 + i] == k ? dOut[g * C + c0 + i] : 0.f;
        } else {
            unpack8(reinterpret_cast<const uint4 *>(dA)[v], d);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            const float dy = fmaf(z[i], scale[c], shift[c]) > 0.f ? d[i] : 0.f;
            const float zh = (z[i] - mean[c]) * rstd[c];
            d[i] = coef[c] * (dy - coef[C + c] - zh * coef[2 * C + c]);
        }
        reinterpret_cast<uint4 *>(dZ)[v] = pack8(d);
    }
}

// Pooled upstream gradient: one thread owns (group, 8 channels), loads arg-max / dOut once and walks the K rows.
__global__ void __launch_bounds__(kEwThreads)
bwd_apply_pooled_kernel(const float *__restrict__ dOut, const int *__restrict__ arg, int K, const __nv_bfloat16 *__restrict__ Z,
                        const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                        const float *__restrict__ rstd, const float *__restrict__ coef, int64_t G, int C,
                        __nv_bfloat16 *__restrict__ dZ)
{
    const int cg = C >> 3;
    for (int64_t v = (int64_t)blockIdx.x * kEwThreads + threadIdx.x; v < G * cg; v += (int64_t)gridDim.x * kEwThreads) {
        const int64_t g = v / cg;
        const int c0 = (int)(v - g * cg) * 8;
        float sc[8], sh[8], mu[8], rs[8], k0[8], k1[8], k2[8], go[8];
        int am[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            sc[i] = scale[c], sh[i] = shift[c], mu[i] = mean[c], rs[i] = rstd[c];
            k0[i] = coef[c], k1[i] = coef[C + c], k2[i] = coef[2 * C + c];
            go[i] = dOut[g * C + c], am[i] = arg[g * C + c];
        }
        const __nv_bfloat16 *zp = Z + (g * K) * C + c0;
        __nv_bfloat16 *dp = dZ + (g * K) * C + c0;
        for (int k = 0; k < K; ++k) {
            float z[8], d[8];
            unpack8(*reinterpret_cast<const uint4 *>(zp + (int64_t)k * C), z);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float dy = (am[i] == k && fmaf(z[i], sc[i], sh[i]) > 0.f) ? go[i] : 0.f;
                d[i] = k0[i] * (dy - k1[i] - (z[i] - mu[i]) * rs[i] * k2[i]);
            }
            *reinterpret_cast<uint4 *>(dp + (int64_t)k * C) = pack8(d);
        }
    }
}

static inline int stat_parts(int64_t rows, int C)
{
    const int rows_per_pass = kEwThreads / (C >> 3);
    int64_t want = (rows + rows_per_pass - 1) / rows_per_pass;
    const int cap = 2 * sm_count();
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}
static inline unsigned ew_blocks(int64_t total)
{
    int64_t b = (total + kEwThreads - 1) / kEwThreads;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace mpb

#define MPB_CHECK_C(C) MPB_REQUIRE((C) > 0 && (C) % 8 == 0 && (C) <= 2048, "C must be a positive multiple of 8, at most 2048")

extern "C" int mpb_bn_stat_partials(int64_t rows, int C)
{
    if (rows <= 0 || C <= 0 || C % 8) return 0;
    return mpb::stat_parts(rows, C);
}

extern "C" int mpb_bn_colstats_bf16(const void *Z, int64_t M, int C, float *partials, int nparts, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_REQUIRE(M > 0 && Z && partials && nparts == stat_parts(M, C), "bad argument");
    colstats_kernel<<<nparts, kEwThreads, 2 * C * sizeof(float), (cudaStream_t)stream>>>((const __nv_bfloat16 *)Z, M, C, partials);
    return check_launch("colstats_kernel");
}

extern "C" int mpb_bn_finalize_f32(const float *partials, int nparts, int C, int C_valid, int64_t M, const float *bias,
                                   const float *gamma, const float *beta, float *running_mean, float *running_var, float momentum,
                                   float eps, float *scale, float *shift, float *mean, float *rstd, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(partials && scale && shift && mean && rstd && C > 0 && nparts > 0 && M > 0, "bad argument");
    MPB_REQUIRE(C_valid >= 0 && C_valid <= C, "C_valid out of range");
    bn_finalize_kernel<<<(C + 7) / 8, kFinThreads, 0, (cudaStream_t)stream>>>(partials, nparts, C, C_valid, (double)M, bias, gamma, beta, running_mean,
                                                            running_var, momentum, eps, scale, shift, mean, rstd);
    return check_launch("bn_finalize_kernel");
}

extern "C" int mpb_bn_relu_bf16(const void *Z, const float *scale, const float *shift, int64_t M, int C, void *A, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_REQUIRE(M >= 0 && Z && scale && shift && A, "bad argument");
    if (M == 0) return MPB_OK;
    const int64_t total = M * (C >> 3);
    bn_relu_kernel<<<ew_blocks(total), kEwThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)Z, scale, shift, total, C >> 3,
                                                                              (__nv_bfloat16 *)A);
    return check_launch("bn_relu_kernel");
}

extern "C" int mpb_bn_relu_max_bf16(const void *Z, const float *scale, const float *shift, int64_t G, int K, int C, float *out,
                                    int32_t *argmax, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_REQUIRE(G >= 0 && K > 0 && Z && scale && shift && out && argmax, "bad argument");
    if (G == 0) return MPB_OK;
    bn_relu_max_kernel<<<ew_blocks(G * (C >> 3)), kEwThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)Z, scale, shift, G, K, C,
                                                                                        out, argmax);
    return check_launch("bn_relu_max_kernel");
}

extern "C" int mpb_bn_bwd_stats_bf16(const void *dA, const float *dOut, const int32_t *argmax, int K, const void *Z, const float *scale,
                                     const float *shift, const float *mean, const float *rstd, int64_t M, int C, float *partials,
                                     int nparts, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_REQUIRE(M > 0 && Z && scale && shift && mean && rstd && partials, "bad argument");
    MPB_REQUIRE((dA != nullptr) != (dOut != nullptr), "exactly one of dA (dense) / dOut (pooled) must be given");
    cudaStream_t st = (cudaStream_t)stream;
    if (dA) {
        MPB_REQUIRE(nparts == stat_parts(M, C), "nparts mismatch");
        bwd_stats_dense_kernel<<<nparts, kEwThreads, 2 * C * sizeof(float), st>>>((const __nv_bfloat16 *)dA, (const __nv_bfloat16 *)Z, scale,
                                                                                 shift, mean, rstd, M, C, partials);
    } else {
        MPB_REQUIRE(argmax && K > 0 && M % K == 0 && nparts == stat_parts(M / K, C), "pooled: bad argmax/K/nparts");
        bwd_stats_pooled_kernel<<<nparts, kEwThreads, 2 * C * sizeof(float), st>>>(dOut, argmax, (const __nv_bfloat16 *)Z, scale, shift, mean,
                                                                                  rstd, M / K, K, C, partials);
    }
    return check_launch("bwd_stats kernel");
}

extern "C" int mpb_bn_bwd_finalize_f32(const float *partials, int nparts, int C, int C_valid, int64_t M, const float *gamma,
                                       const float *rstd, float *dgamma, float *dbeta, float *coef, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(partials && rstd && coef && C > 0 && nparts > 0 && M > 0, "bad argument");
    MPB_REQUIRE(C_valid >= 0 && C_valid <= C, "C_valid out of range");
    bwd_finalize_kernel<<<(C + 7) / 8, kFinThreads, 0, (cudaStream_t)stream>>>(partials, nparts, C, C_valid, (double)M, gamma, rstd, dgamma, dbeta, coef);
    return check_launch("bwd_finalize_kernel");
}

extern "C" int mpb_bn_bwd_apply_bf16(const void *dA, const float *dOut, const int32_t *argmax, int K, const void *Z, const float *scale,
                                     const float *shift, const float *mean, const float *rstd, const float *coef, int64_t M, int C,
                                     void *dZ, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_REQUIRE(M > 0 && Z && scale && shift && mean && rstd && coef && dZ, "bad argument");
    MPB_REQUIRE((dA != nullptr) != (dOut != nullptr), "exactly one of dA (dense) / dOut (pooled) must be given");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = ew_blocks(M * (C >> 3));
    if (dA)
        bwd_apply_kernel<false><<<blocks, kEwThreads, 0, st>>>((const __nv_bfloat16 *)dA, nullptr, nullptr, 1, (const __nv_bfloat16 *)Z, scale,
                                                              shift, mean, rstd, coef, M, C, (__nv_bfloat16 *)dZ);
    else {
        MPB_REQUIRE(argmax && K > 0 && M % K == 0, "pooled: bad argmax/K");
        bwd_apply_pooled_kernel<<<ew_blocks((M / K) * (C >> 3)), kEwThreads, 0, st>>>(dOut, argmax, K, (const __nv_bfloat16 *)Z, scale, shift,
                                                                                     mean, rstd, coef, M / K, C, (__nv_bfloat16 *)dZ);
    }
    return check_launch("bwd_apply_kernel");
}
