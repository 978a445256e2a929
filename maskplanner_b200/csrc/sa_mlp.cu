// BatchNorm / ReLU / max-pool stages of PointNetSetAbstraction's shared MLP on bf16 activations (sm_100a).
// Reference: models/pointnet2_utils.py:210-214 -- per layer `relu(bn(conv(x)))`, then `max` over the K
// neighbours of each group.  The 1x1 conv runs in sa_gemm.cu and leaves Z[M,C] (bf16) in HBM; everything
// here is HBM-bound streaming over [M,C] with 16-byte vector accesses:
//
//   colstats      : per-channel sum / sum of squares of Z (training-mode batch statistics)
//   bn_finalize   : -> scale = gamma*rstd, shift = beta - mean*scale; running-stat update (momentum, unbiased var)
//   bn_relu       : A = relu(scale*Z + shift)                                   (bf16 -> bf16)
//   bn_relu_max   : out[g,c] = max_k relu(scale*Z[g,k,c] + shift) + arg-max     (last layer, fused max-pool)
//   bwd_stats     : sum dY, sum dY*zhat with dY = dA * [scale*Z+shift > 0]      (dense or pooled upstream)
//   bwd_finalize  : dgamma, dbeta and the three per-channel coefficients BN backward needs
//   bwd_apply     : dZ = gamma*rstd*(dY - mean(dY) - zhat*mean(dY*zhat))        (bf16 out)
//
// Thread mapping of the streaming kernels: a thread owns ONE 8-channel group (16 bytes of a row) for its
// whole life and walks rows with a grid stride, so all per-channel constants live in registers and the only
// per-row work is two or three 16-byte loads, a handful of FMAs and one 16-byte store; rows are processed
// four at a time to keep four independent loads in flight per thread.
// Batch statistics are accumulated in fp32 per CTA and combined in fp64 by the finalize kernels
// (deterministic: fixed partial order, no atomics across CTAs).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mpb {

// Storage type of the activations: bf16 (tensor-core path, tolerance 1e-2) or fp32 (TF32 / 3xTF32 path, tolerance 1e-4).
// A thread always handles 8 consecutive channels of a row: one 16-byte access for bf16, two for fp32.
template <class T>
struct Act;
template <>
struct Act<__nv_bfloat16> {
    typedef uint4 raw_t;
    static __device__ __forceinline__ raw_t ld(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint4 *>(p); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, const raw_t &v) { *reinterpret_cast<uint4 *>(p) = v; }
    static __device__ __forceinline__ void unpack(const raw_t &v, float (&f)[8])
    {
        const __nv_bfloat162 *p = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __bfloat1622float2(p[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    static __device__ __forceinline__ raw_t pack(const float (&f)[8])
    {
        uint4 v;
        __nv_bfloat162 *p = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
    static __device__ __forceinline__ float round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }   // what a store keeps
    static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
};
template <>
struct Act<float> {
    struct raw_t {
        float4 a, b;
    };
    static __device__ __forceinline__ raw_t ld(const float *p)
    {
        raw_t r;
        r.a = *reinterpret_cast<const float4 *>(p), r.b = *reinterpret_cast<const float4 *>(p + 4);
        return r;
    }
    static __device__ __forceinline__ void st(float *p, const raw_t &v)
    {
        *reinterpret_cast<float4 *>(p) = v.a, *reinterpret_cast<float4 *>(p + 4) = v.b;
    }
    static __device__ __forceinline__ void unpack(const raw_t &v, float (&f)[8])
    {
        f[0] = v.a.x, f[1] = v.a.y, f[2] = v.a.z, f[3] = v.a.w, f[4] = v.b.x, f[5] = v.b.y, f[6] = v.b.z, f[7] = v.b.w;
    }
    static __device__ __forceinline__ raw_t pack(const float (&f)[8])
    {
        raw_t r;
        r.a = make_float4(f[0], f[1], f[2], f[3]), r.b = make_float4(f[4], f[5], f[6], f[7]);
        return r;
    }
    static __device__ __forceinline__ float round(float v) { return v; }
    static __device__ __forceinline__ float to_float(float v) { return v; }
};
__device__ __forceinline__ void load8(const float *__restrict__ p, float (&f)[8])
{
    const float4 a = *reinterpret_cast<const float4 *>(p), b = *reinterpret_cast<const float4 *>(p + 4);
    f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}

constexpr int kEwThreads = 256;
constexpr int kRowUnroll = 4;
// Grids are ONE wave of resident CTAs (each streams a contiguous slab of rows): the statistics kernels are
// compiled for 4 CTAs/SM (<= 64 registers) and launched 4 per SM; the elementwise kernels ask the occupancy
// calculator.  (Before: 8 CTAs per SM requested with 3 resident => 2.67 waves, an 11 % tail.)
constexpr int kStatCtasPerSm = 4;

// Row walker: thread (tr, tc) of a CTA handles channel group tc (channels 8*tc .. 8*tc+7) of rows
// blockIdx.x*rpp + tr, + gridDim.x*rpp, ...; `body(r0, nr, stride)` is called with up to kRowUnroll rows.
struct RowWalk {
    int cg, rpp, tr, tc;
    bool active;
    __device__ __forceinline__ RowWalk(int C)
    {
        cg = C >> 3;
        rpp = kEwThreads / cg;
        tr = threadIdx.x / cg;
        tc = threadIdx.x - tr * cg;
        active = tr < rpp;
    }
};

// Contiguous slab of rows owned by this CTA: [slab_begin, slab_end), slab size a multiple of rpp.
__device__ __forceinline__ int64_t slab_rows(int64_t M, int rpp)
{
    const int64_t per = (M + gridDim.x - 1) / gridDim.x;
    return (per + rpp - 1) / rpp * rpp;
}
__device__ __forceinline__ int64_t slab_begin(int64_t M, int rpp) { return (int64_t)blockIdx.x * slab_rows(M, rpp); }
__device__ __forceinline__ int64_t slab_end(int64_t M, int rpp)
{
    const int64_t e = slab_begin(M, rpp) + slab_rows(M, rpp);
    return e < M ? e : M;
}

// ---- training statistics ---------------------------------------------------------------------------
template <class RowFn>
__device__ __forceinline__ void column_sums(int64_t M, int C, float *__restrict__ partials, RowFn fn)
{
    extern __shared__ float sm_acc[];  // [rpp][2][C]: one slot per thread, combined in a FIXED order (bitwise reproducible)
    const RowWalk w(C);
    float s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
    if (w.active) {
        // each CTA streams one contiguous slab of rows (consecutive 16-byte x cg x rpp blocks: DRAM-page friendly)
        const int64_t r_end = slab_end(M, w.rpp);
        const int64_t stride = w.rpp;
        int64_t r = slab_begin(M, w.rpp) + w.tr;
        for (; r + (kRowUnroll - 1) * stride < r_end; r += kRowUnroll * stride) fn(r, stride, kRowUnroll, w.tc * 8, s0, s1);
        for (; r < r_end; r += stride) fn(r, stride, 1, w.tc * 8, s0, s1);
        float4 *d0 = reinterpret_cast<float4 *>(sm_acc + (size_t)w.tr * 2 * C + w.tc * 8);
        float4 *d1 = reinterpret_cast<float4 *>(sm_acc + (size_t)w.tr * 2 * C + C + w.tc * 8);
        d0[0] = make_float4(s0[0], s0[1], s0[2], s0[3]), d0[1] = make_float4(s0[4], s0[5], s0[6], s0[7]);
        d1[0] = make_float4(s1[0], s1[1], s1[2], s1[3]), d1[1] = make_float4(s1[4], s1[5], s1[6], s1[7]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kEwThreads) {
        float t = 0.f;
        for (int tr = 0; tr < w.rpp; ++tr) t += sm_acc[(size_t)tr * 2 * C + i];
        partials[(size_t)blockIdx.x * 2 * C + i] = t;
    }
}

template <class T>
__global__ void __launch_bounds__(kEwThreads, kStatCtasPerSm)
colstats_kernel(const T *__restrict__ Z, int64_t M, int C, float *__restrict__ partials)
{
    column_sums(M, C, partials, [&](int64_t r, int64_t stride, int nr, int c0, float(&s0)[8], float(&s1)[8]) {
        typename Act<T>::raw_t raw[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (u < nr) raw[u] = Act<T>::ld(Z + (r + u * stride) * C + c0);
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (u < nr) {
                float z[8];
                Act<T>::unpack(raw[u], z);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s0[i] += z[i];
                    s1[i] = fmaf(z[i], z[i], s1[i]);
                }
            }
    });
}

// Combine the per-CTA partial sums of 8 channels (one CTA = 8 channels x 32 partial lanes) in fp64, fixed order.
// Returns true on the thread that holds the totals of channel `c`.
constexpr int kFinThreads = 512;
// 8 columns per block, kFinThreads/8 partial lanes per column.  The loads of a column are independent, so
// they are issued four deep before the fp64 adds (the old one-load-per-iteration loop was pure L2 latency:
// ~16 us for 1184 partials); lanes combine by warp shuffles, then one add per warp.
__device__ __forceinline__ bool reduce_partials(const float *__restrict__ partials, int nparts, int C, int &c, double &s, double &q)
{
    constexpr int LANES = kFinThreads / 8, WARPS = kFinThreads / 32;
    __shared__ double red[2][WARPS][8];
    const int pl = threadIdx.x >> 3, ci = threadIdx.x & 7;
    c = blockIdx.x * 8 + ci;
    s = 0.0, q = 0.0;
    if (c < C)
        for (int p = pl; p < nparts; p += 4 * LANES) {
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int pp = p + u * LANES;
                const bool ok = pp < nparts;
                a[u] = ok ? partials[(size_t)pp * 2 * C + c] : 0.f;
                b[u] = ok ? partials[(size_t)pp * 2 * C + C + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) s += (double)a[u], q += (double)b[u];
        }
    // a warp holds 4 partial lanes x 8 columns: fold lanes 8 and 16 apart
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    q += __shfl_xor_sync(0xffffffffu, q, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    q += __shfl_xor_sync(0xffffffffu, q, 16);
    if ((threadIdx.x & 31) < 8) {
        red[0][threadIdx.x >> 5][ci] = s;
        red[1][threadIdx.x >> 5][ci] = q;
    }
    __syncthreads();
    if (pl != 0 || c >= C) return false;
    s = 0.0, q = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += red[0][w][ci], q += red[1][w][ci];
    return true;
}

// Training: batch statistics -> scale/shift/mean/rstd, running stats updated in place.
// (conv bias only moves the mean: BN(z + b) == BN(z); it enters the running mean, reference :210-212.)
__global__ void __launch_bounds__(kFinThreads)
bn_finalize_kernel(const float *__restrict__ partials, int nparts, int C, int C_valid, double M, const float *__restrict__ bias,
                   const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ running_mean,
                   float *__restrict__ running_var, float momentum, float eps, float *__restrict__ scale, float *__restrict__ shift,
                   float *__restrict__ mean_out, float *__restrict__ rstd_out, int64_t *__restrict__ num_batches_tracked)
{
    if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0) *num_batches_tracked += 1;   // BatchNorm's own counter (one launch less)
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

// ---- forward elementwise ---------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kEwThreads)
bn_relu_kernel(const T *__restrict__ Z, const float *__restrict__ scale, const float *__restrict__ shift, int64_t M, int C,
               T *__restrict__ A)
{
    const RowWalk w(C);
    if (!w.active) return;
    const int c0 = w.tc * 8;
    float sc[8], sh[8];
    load8(scale + c0, sc);
    load8(shift + c0, sh);
    const int64_t stride = w.rpp, r_end = slab_end(M, w.rpp);
    for (int64_t r = slab_begin(M, w.rpp) + w.tr; r < r_end; r += kRowUnroll * stride) {
        typename Act<T>::raw_t raw[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (r + u * stride < r_end) raw[u] = Act<T>::ld(Z + (r + u * stride) * C + c0);
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (r + u * stride < r_end) {
                float z[8];
                Act<T>::unpack(raw[u], z);
#pragma unroll
                for (int i = 0; i < 8; ++i) z[i] = fmaxf(fmaf(z[i], sc[i], sh[i]), 0.f);
                Act<T>::st(A + (r + u * stride) * C + c0, Act<T>::pack(z));
            }
    }
}

constexpr int kMaxUnroll = 4;   // rows in flight per thread in the pooling kernel (8 at 3 CTAs/SM: 66-69 us, 4 at 4 CTAs/SM: 63-66 us)
// out[g, c] = max_k relu(scale*Z[g*K + k, c] + shift); arg[g, c] = first k attaining it (torch.max semantics).
// relu(scale*z + shift) is monotone in z (non-decreasing for scale >= 0, non-increasing otherwise, and a correctly rounded fma
// keeps that), so the pooled row is the first maximum of s*z with s = sign(scale): the scan runs on the RAW pre-activations --
// for bf16 in packed form: one XOR flips the sign of the decreasing channels, one HSETP2 compares two channels against the
// running maxima, one HMNMX2 updates them, two selects keep the row index; ~20 instructions per 16 bytes and 12 registers of
// state instead of ~50 (unpack, fma, max, compare, two selects per element) and 24 -- and the affine + ReLU is applied ONCE
// per (group, channel) at the end.  Groups that the ReLU kills entirely (pooled value 0) report the arg-max of s*z instead of
// row 0; no gradient flows through them either way.  Round-1 form: 69-71 us for 268 MB (3.8 TB/s), this one 63-66 us (4.2 TB/s);
// a value-only variant (max / min of z, no index: 8 instructions per 16 bytes) ran at 55 us but moved the arg-max search into
// bwd_apply_pooled (+9 us each), so the index stays here.
__device__ __forceinline__ void setp_gt_bf16x2(uint32_t a, uint32_t b, bool &lo, bool &hi)
{
    uint32_t l, h;
    asm("{\n\t.reg .pred p, q;\n\t"
        "setp.gt.bf16x2 p|q, %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "selp.u32 %1, 1, 0, q;\n\t}"
        : "=r"(l), "=r"(h)
        : "r"(a), "r"(b));
    lo = l != 0, hi = h != 0;
}
struct PoolScanBf16 {
    uint32_t best[4], sgn[4];
    int bi[8];
    __device__ __forceinline__ void init(const float (&sc)[8])
    {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            best[i] = 0xFF80FF80u;                                   // (-inf, -inf)
            sgn[i] = (sc[2 * i] < 0.f ? 0x00008000u : 0u) | (sc[2 * i + 1] < 0.f ? 0x80000000u : 0u);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) bi[i] = 0;
    }
    __device__ __forceinline__ void update(const uint4 &raw, int k)
    {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t v = w[i] ^ sgn[i];
            bool lo, hi;
            setp_gt_bf16x2(v, best[i], lo, hi);
            bi[2 * i] = lo ? k : bi[2 * i];
            bi[2 * i + 1] = hi ? k : bi[2 * i + 1];
            __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&v), *reinterpret_cast<const __nv_bfloat162 *>(&best[i]));
            best[i] = *reinterpret_cast<uint32_t *>(&m);
        }
    }
    __device__ __forceinline__ void get(float (&z)[8]) const     // the selected pre-activations, sign restored
    {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t v = best[i] ^ sgn[i];
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&v));
            z[2 * i] = f.x, z[2 * i + 1] = f.y;
        }
    }
};
struct PoolScanF32 {
    float best[8], sg[8];
    int bi[8];
    __device__ __forceinline__ void init(const float (&sc)[8])
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) best[i] = -INFINITY, sg[i] = sc[i] < 0.f ? -1.f : 1.f, bi[i] = 0;
    }
    __device__ __forceinline__ void update(const Act<float>::raw_t &raw, int k)
    {
        float z[8];
        Act<float>::unpack(raw, z);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = z[i] * sg[i];
            if (v > best[i]) best[i] = v, bi[i] = k;
        }
    }
    __device__ __forceinline__ void get(float (&z)[8]) const
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = best[i] * sg[i];
    }
};
template <class T>
struct PoolScanOf {
    typedef PoolScanF32 type;
};
template <>
struct PoolScanOf<__nv_bfloat16> {
    typedef PoolScanBf16 type;
};

template <class T>
__global__ void __launch_bounds__(kEwThreads, sizeof(T) == 2 ? 4 : 2)
bn_relu_max_kernel(const T *__restrict__ Z, const float *__restrict__ scale, const float *__restrict__ shift, int64_t G,
                   int K, int C, float *__restrict__ out, int *__restrict__ arg, float *__restrict__ zmax)
{
    const RowWalk w(C);
    if (!w.active) return;
    const int c0 = w.tc * 8;
    for (int64_t g = (int64_t)blockIdx.x * w.rpp + w.tr; g < G; g += (int64_t)gridDim.x * w.rpp) {
        typename PoolScanOf<T>::type scan;
        {
            float sc[8];
            load8(scale + c0, sc);
            scan.init(sc);          // only the signs stay in registers during the scan
        }
        const T *zp = Z + (g * K) * C + c0;
        for (int k = 0; k < K; k += kMaxUnroll) {
            typename Act<T>::raw_t raw[kMaxUnroll];
#pragma unroll
            for (int u = 0; u < kMaxUnroll; ++u)
                if (k + u < K) raw[u] = Act<T>::ld(zp + (int64_t)(k + u) * C);
#pragma unroll
            for (int u = 0; u < kMaxUnroll; ++u)
                if (k + u < K) scan.update(raw[u], k + u);
        }
        float bz[8], best[8], sc[8], sh[8];
        scan.get(bz);
        load8(scale + c0, sc);
        load8(shift + c0, sh);
#pragma unroll
        for (int i = 0; i < 8; ++i) best[i] = fmaxf(fmaf(bz[i], sc[i], sh[i]), 0.f);
        float4 *op = reinterpret_cast<float4 *>(out + g * C + c0);
        op[0] = make_float4(best[0], best[1], best[2], best[3]);
        op[1] = make_float4(best[4], best[5], best[6], best[7]);
        int4 *ap = reinterpret_cast<int4 *>(arg + g * C + c0);
        ap[0] = make_int4(scan.bi[0], scan.bi[1], scan.bi[2], scan.bi[3]);
        ap[1] = make_int4(scan.bi[4], scan.bi[5], scan.bi[6], scan.bi[7]);
        if (zmax) {   // pre-activation at the arg-max row: lets the backward statistics skip a 2-byte gather per (g, c)
            float4 *zp4 = reinterpret_cast<float4 *>(zmax + g * C + c0);
            zp4[0] = make_float4(bz[0], bz[1], bz[2], bz[3]);
            zp4[1] = make_float4(bz[4], bz[5], bz[6], bz[7]);
        }
    }
}

// Few groups, many rows per group (SA3: G = batch size, K = 128): one 1024-thread CTA per group, the K rows are
// interleaved over 1024/(C/8) segments and the per-segment (max, arg, z) triples meet in shared memory.
constexpr int kWideThreads = 1024;
template <class T>
__global__ void __launch_bounds__(kWideThreads)
bn_relu_max_wide_kernel(const T *__restrict__ Z, const float *__restrict__ scale, const float *__restrict__ shift, int64_t G,
                        int K, int C, float *__restrict__ out, int *__restrict__ arg, float *__restrict__ zmax)
{
    extern __shared__ float sm_wide[];           // [3][segments][C]: best, arg (as int bits), z
    const int cg = C >> 3, nseg = kWideThreads / cg;
    const int seg = threadIdx.x / cg, tc = threadIdx.x - seg * cg;
    const bool active = seg < nseg;
    const int c0 = tc * 8;
    float *s_best = sm_wide, *s_z = sm_wide + 2 * nseg * C;
    int *s_arg = reinterpret_cast<int *>(sm_wide + nseg * C);
    float sc[8], sh[8];
    if (active) {
        load8(scale + c0, sc);
        load8(shift + c0, sh);
    }
    for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
        float best[8], bz[8];
        int bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) best[i] = -INFINITY, bi[i] = 0, bz[i] = 0.f;
        if (active) {
            const T *zp = Z + (g * K) * C + c0;
            for (int k = seg; k < K; k += kRowUnroll * nseg) {
                typename Act<T>::raw_t raw[kRowUnroll];
#pragma unroll
                for (int u = 0; u < kRowUnroll; ++u)
                    if (k + u * nseg < K) raw[u] = Act<T>::ld(zp + (int64_t)(k + u * nseg) * C);
#pragma unroll
                for (int u = 0; u < kRowUnroll; ++u)
                    if (k + u * nseg < K) {
                        float z[8];
                        Act<T>::unpack(raw[u], z);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float a = fmaxf(fmaf(z[i], sc[i], sh[i]), 0.f);
                            if (a > best[i]) best[i] = a, bi[i] = k + u * nseg, bz[i] = z[i];
                        }
                    }
            }
            if (seg > 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s_best[seg * C + c0 + i] = best[i];
                    s_arg[seg * C + c0 + i] = bi[i];
                    s_z[seg * C + c0 + i] = bz[i];
                }
            }
        }
        __syncthreads();
        if (active && seg == 0) {
            for (int s2 = 1; s2 < nseg; ++s2)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float a = s_best[s2 * C + c0 + i];
                    const int k2 = s_arg[s2 * C + c0 + i];
                    if (a > best[i] || (a == best[i] && k2 < bi[i])) best[i] = a, bi[i] = k2, bz[i] = s_z[s2 * C + c0 + i];  // first k wins ties
                }
            float4 *op = reinterpret_cast<float4 *>(out + g * C + c0);
            op[0] = make_float4(best[0], best[1], best[2], best[3]);
            op[1] = make_float4(best[4], best[5], best[6], best[7]);
            int4 *ap = reinterpret_cast<int4 *>(arg + g * C + c0);
            ap[0] = make_int4(bi[0], bi[1], bi[2], bi[3]);
            ap[1] = make_int4(bi[4], bi[5], bi[6], bi[7]);
            if (zmax) {
                float4 *zp4 = reinterpret_cast<float4 *>(zmax + g * C + c0);
                zp4[0] = make_float4(bz[0], bz[1], bz[2], bz[3]);
                zp4[1] = make_float4(bz[4], bz[5], bz[6], bz[7]);
            }
        }
        __syncthreads();
    }
}

// ---- backward --------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(kEwThreads, kStatCtasPerSm)
bwd_stats_dense_kernel(const T *__restrict__ dA, const T *__restrict__ Z, const float *__restrict__ scale,
                       const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd, int64_t M,
                       int C, float *__restrict__ partials)
{
    const int c00 = (threadIdx.x % (C >> 3)) * 8;
    float sc[8], sh[8];
    load8(scale + c00, sc);
    load8(shift + c00, sh);
    (void)mean, (void)rstd;
    column_sums(M, C, partials, [&](int64_t r, int64_t stride, int nr, int c0, float(&s0)[8], float(&s1)[8]) {
        typename Act<T>::raw_t rz[kRowUnroll], rd[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (u < nr) {
                rz[u] = Act<T>::ld(Z + (r + u * stride) * C + c0);
                rd[u] = Act<T>::ld(dA + (r + u * stride) * C + c0);
            }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (u < nr) {
                float z[8], d[8];
                Act<T>::unpack(rz[u], z);
                Act<T>::unpack(rd[u], d);
#pragma unroll
                for (int i = 0; i < 8; ++i) {   // raw moments sum dY, sum dY*z; the finalize kernel turns them into sum dY*zhat in fp64
                    const float dy = fmaf(z[i], sc[i], sh[i]) > 0.f ? d[i] : 0.f;
                    s0[i] += dy;
                    s1[i] = fmaf(dy, z[i], s1[i]);
                }
            }
    });
}

// Upstream gradient is the pooled one: only the arg-max row of each (group, channel) carries dOut.
template <class T>
__global__ void __launch_bounds__(kEwThreads, kStatCtasPerSm)
bwd_stats_pooled_kernel(const float *__restrict__ dOut, const int *__restrict__ arg, const T *__restrict__ Z,
                        const float *__restrict__ zmax, const float *__restrict__ scale, const float *__restrict__ shift,
                        const float *__restrict__ mean, const float *__restrict__ rstd, int64_t G, int K, int C,
                        float *__restrict__ partials, float *__restrict__ pgo)
{
    column_sums(G, C, partials, [&](int64_t g0, int64_t stride, int nr, int c0, float(&s0)[8], float(&s1)[8]) {
        for (int u = 0; u < nr; ++u) {
            const int64_t g = g0 + u * stride;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = c0 + i;
                const float z = zmax ? zmax[g * C + c] : Act<T>::to_float(Z[(g * K + arg[g * C + c]) * C + c]);
                const float dy = fmaf(z, scale[c], shift[c]) > 0.f ? dOut[g * C + c] : 0.f;
                s0[i] += dy;
                s1[i] = fmaf(dy, z, s1[i]);
                // the one non-zero of the pooled upstream gradient in (group g, channel c), pre-multiplied by p = gamma * rstd
                // (= scale): what the GEMMs' pooled operand transform adds at row arg[g, c] (mpb_sa_gemm_*_pool)
                if (pgo) pgo[g * C + c] = scale[c] * dy;
            }
        }
    });
}

// Partials hold the raw moments (sum dY, sum dY*z); sum dY*zhat = rstd*(sum dY*z - mean*sum dY), formed in fp64.
// dgamma = sum dY*zhat, dbeta = sum dY; coef[0][c] = gamma*rstd, coef[1][c] = mean(dY), coef[2][c] = mean(dY*zhat)
__global__ void __launch_bounds__(kFinThreads)
bwd_finalize_kernel(const float *__restrict__ partials, int nparts, int C, int C_valid, double M, const float *__restrict__ gamma,
                    const float *__restrict__ mean, const float *__restrict__ rstd, float *__restrict__ dgamma,
                    float *__restrict__ dbeta, float *__restrict__ coef, float4 *__restrict__ clear, int64_t clear_vec4,
                    float *__restrict__ negw_e)
{
    // side job: zero the accumulation buffer of the weight-gradient GEMM that follows (saves a fill launch)
    for (int64_t i = (int64_t)blockIdx.x * kFinThreads + threadIdx.x; i < clear_vec4; i += (int64_t)gridDim.x * kFinThreads)
        clear[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    int c;
    double s, q;
    if (reduce_partials(partials, nparts, C, c, s, q)) {
        if (c >= C_valid) {
            coef[c] = coef[C + c] = coef[2 * C + c] = 0.f;
            if (negw_e) negw_e[c] = negw_e[C + c] = 0.f;
            return;
        }
        q = (double)rstd[c] * (q - (double)mean[c] * s);
        if (dbeta) dbeta[c] = (float)s;
        if (dgamma) dgamma[c] = (float)q;
        coef[c] = (gamma ? gamma[c] : 1.f) * rstd[c];
        coef[C + c] = (float)(s / M);
        coef[2 * C + c] = (float)(q / M);
        if (negw_e) {   // dZ = p*dY - w*z + e (ApplyConst), handed to the GEMMs' pooled operand transform as (-w, e)
            const float w = coef[c] * coef[2 * C + c] * rstd[c];
            negw_e[c] = -w;
            negw_e[C + c] = fmaf(mean[c], w, -coef[c] * coef[C + c]);
        }
    }
}

// Per-thread constants of dZ = k0*(dY - k1 - zhat*k2) rewritten as dZ = p*dY - w*z + e:
//   p = k0, w = k0*k2*rstd, e = mean*w - k0*k1
struct ApplyConst {
    float sc[8], sh[8], p[8], w[8], e[8];
    __device__ __forceinline__ void load(const float *scale, const float *shift, const float *mean, const float *rstd, const float *coef,
                                         int C, int c0)
    {
        float mu[8], rs[8], k1[8], k2[8];
        load8(scale + c0, sc);
        load8(shift + c0, sh);
        load8(mean + c0, mu);
        load8(rstd + c0, rs);
        load8(coef + c0, p);
        load8(coef + C + c0, k1);
        load8(coef + 2 * C + c0, k2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            w[i] = p[i] * k2[i] * rs[i];
            e[i] = fmaf(mu[i], w[i], -p[i] * k1[i]);
        }
    }
};

template <class T>
__global__ void __launch_bounds__(kEwThreads)
bwd_apply_dense_kernel(const T *__restrict__ dA, const T *__restrict__ Z, const float *__restrict__ scale,
                       const float *__restrict__ shift, const float *__restrict__ mean, const float *__restrict__ rstd,
                       const float *__restrict__ coef, int64_t M, int C, T *__restrict__ dZ)
{
    const RowWalk w(C);
    if (!w.active) return;
    const int c0 = w.tc * 8;
    ApplyConst k;
    k.load(scale, shift, mean, rstd, coef, C, c0);
    const int64_t stride = w.rpp, r_end = slab_end(M, w.rpp);
    for (int64_t r = slab_begin(M, w.rpp) + w.tr; r < r_end; r += kRowUnroll * stride) {
        typename Act<T>::raw_t rz[kRowUnroll], rd[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (r + u * stride < r_end) {
                rz[u] = Act<T>::ld(Z + (r + u * stride) * C + c0);
                rd[u] = Act<T>::ld(dA + (r + u * stride) * C + c0);
            }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (r + u * stride < r_end) {
                float z[8], d[8];
                Act<T>::unpack(rz[u], z);
                Act<T>::unpack(rd[u], d);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dy = fmaf(z[i], k.sc[i], k.sh[i]) > 0.f ? d[i] : 0.f;
                    d[i] = fmaf(k.p[i], dy, fmaf(-k.w[i], z[i], k.e[i]));
                }
                Act<T>::st(dZ + (r + u * stride) * C + c0, Act<T>::pack(d));
            }
    }
}

// Pooled upstream gradient: a thread owns (group, 8 channels), loads arg-max / dOut once and walks the K rows.
template <class T>
__global__ void __launch_bounds__(kEwThreads)
bwd_apply_pooled_kernel(const float *__restrict__ dOut, const int *__restrict__ arg, int K, const T *__restrict__ Z,
                        const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                        const float *__restrict__ rstd, const float *__restrict__ coef, int64_t G, int C,
                        T *__restrict__ dZ)
{
    const RowWalk w(C);
    if (!w.active) return;
    const int c0 = w.tc * 8;
    ApplyConst k;
    k.load(scale, shift, mean, rstd, coef, C, c0);
    // few groups (SA3: G = batch size): gridDim.y splits the K rows of a group so the launch still fills the machine
    const int kseg = ((K + (int)gridDim.y - 1) / (int)gridDim.y + kRowUnroll - 1) / kRowUnroll * kRowUnroll;
    const int kb = (int)blockIdx.y * kseg, ke = min(K, kb + kseg);
    for (int64_t g = (int64_t)blockIdx.x * w.rpp + w.tr; g < G; g += (int64_t)gridDim.x * w.rpp) {
        float go[8];
        int am[8];
        load8(dOut + g * C + c0, go);
        {
            const int4 a = *reinterpret_cast<const int4 *>(arg + g * C + c0), b = *reinterpret_cast<const int4 *>(arg + g * C + c0 + 4);
            am[0] = a.x, am[1] = a.y, am[2] = a.z, am[3] = a.w, am[4] = b.x, am[5] = b.y, am[6] = b.z, am[7] = b.w;
        }
        const T *zp = Z + (g * K) * C + c0;
        T *dp = dZ + (g * K) * C + c0;
        for (int kk = kb; kk < ke; kk += kRowUnroll) {
            typename Act<T>::raw_t raw[kRowUnroll];
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (kk + u < ke) raw[u] = Act<T>::ld(zp + (int64_t)(kk + u) * C);
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (kk + u < ke) {
                    float z[8], d[8];
                    Act<T>::unpack(raw[u], z);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float dy = (am[i] == kk + u && fmaf(z[i], k.sc[i], k.sh[i]) > 0.f) ? go[i] : 0.f;
                        d[i] = fmaf(k.p[i], dy, fmaf(-k.w[i], z[i], k.e[i]));
                    }
                    Act<T>::st(dp + (int64_t)(kk + u) * C, Act<T>::pack(d));
                }
        }
    }
}

// The same pooled backward from PRE-DIGESTED inputs: pgo[g, c] = p * dY at the pooled row (ReLU gate included, written by
// bwd_stats_pooled_kernel) and (-w, e) per channel (written by bwd_finalize_kernel), so  dZ = pgo * [row == arg] - w*z + e  needs
// 16 per-channel registers instead of 40, no ReLU re-evaluation and no coefficient algebra per thread: 117 -> ~70 registers,
// one more resident CTA per SM.
template <class T>
__global__ void __launch_bounds__(kEwThreads, 3)
bwd_apply_pooled_lean_kernel(const float *__restrict__ pgo, const int *__restrict__ arg, int K, const T *__restrict__ Z,
                             const float *__restrict__ negw_e, int64_t G, int C, T *__restrict__ dZ)
{
    const RowWalk w(C);
    if (!w.active) return;
    const int c0 = w.tc * 8;
    float nw[8], e[8];
    load8(negw_e + c0, nw);
    load8(negw_e + C + c0, e);
    const int kseg = ((K + (int)gridDim.y - 1) / (int)gridDim.y + kRowUnroll - 1) / kRowUnroll * kRowUnroll;
    const int kb = (int)blockIdx.y * kseg, ke = min(K, kb + kseg);
    for (int64_t g = (int64_t)blockIdx.x * w.rpp + w.tr; g < G; g += (int64_t)gridDim.x * w.rpp) {
        float pg[8];
        int am[8];
        load8(pgo + g * C + c0, pg);
        {
            const int4 a = *reinterpret_cast<const int4 *>(arg + g * C + c0), b = *reinterpret_cast<const int4 *>(arg + g * C + c0 + 4);
            am[0] = a.x, am[1] = a.y, am[2] = a.z, am[3] = a.w, am[4] = b.x, am[5] = b.y, am[6] = b.z, am[7] = b.w;
        }
        const T *zp = Z + (g * K) * C + c0;
        T *dp = dZ + (g * K) * C + c0;
        for (int kk = kb; kk < ke; kk += kRowUnroll) {
            typename Act<T>::raw_t raw[kRowUnroll];
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (kk + u < ke) raw[u] = Act<T>::ld(zp + (int64_t)(kk + u) * C);
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (kk + u < ke) {
                    float z[8], d[8];
                    Act<T>::unpack(raw[u], z);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[i] = fmaf(nw[i], z[i], e[i]) + (am[i] == kk + u ? pg[i] : 0.f);
                    Act<T>::st(dp + (int64_t)(kk + u) * C, Act<T>::pack(d));
                }
        }
    }
}

// ---- narrow first layer (SA1: 3 or 6 input channels) ------------------------------------------------
// When the grouped row is only [feats(D) | centred xyz(3)] with 3 + D <= 8, padding it to a 64-column bf16 GEMM
// operand costs 128 B/row of HBM three times (write, GEMM read, weight-gradient read) for 6-16 real bytes.  These
// two kernels never materialise it: the forward gathers the row on the fly (same values, same bf16 rounding as
// group_rows_bf16_kernel), forms Z = row . W^T on CUDA cores (fp32 FMAs over <= 8 terms; bf16 operands, like the
// tensor-core path) and emits the BatchNorm statistic partials; the backward recomputes the row, forms dZ exactly
// like bwd_apply_dense_kernel and reduces dW = dZ^T . row directly (no dZ tensor, no weight-gradient GEMM).
struct NarrowRows {   // element strides are validated to fit 31 bits on the host: offsets are single IMAD.WIDEs
    const float *xyz;
    int32_t xsb, xsn, xsc;
    const float *feats;
    int32_t fsb, fsn, fsc;
    const float *new_xyz;
    const int64_t *idx;
    int N, S, K, D;
};

template <class T, int CIN>
__device__ __forceinline__ void gather_narrow_row(const NarrowRows &g, uint32_t row, float (&a)[CIN])
{
#pragma unroll
    for (int k = 0; k < CIN; ++k) a[k] = 0.f;
    const int64_t i = g.idx[row];
    if (i < 0 || i >= g.N) return;   // empty-ball sentinel (index N): an all-zero row, as in group_rows_bf16_kernel
    const uint32_t bs = row / (uint32_t)g.K;
    const int64_t b = bs / (uint32_t)g.S;
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
        float v = 0.f;
        if (k < g.D)
            v = g.feats[b * g.fsb + i * g.fsn + k * g.fsc];
        else if (k < g.D + 3)
            v = __fsub_rn(g.xyz[b * g.xsb + i * g.xsn + (k - g.D) * g.xsc], g.new_xyz[(int64_t)bs * 3 + (k - g.D)]);
        a[k] = Act<T>::round(v);
    }
}

// The rows one thread visits form an arithmetic progression (r, r + stride, ...), so the decomposition
// row = (b*S + s)*K + k is carried along instead of being divided out per row (integer division goes through the
// XU pipe: the first version of these kernels stalled 45 % of its cycles on it).
struct RowCursor {
    uint32_t bs, k, b, s;
    __device__ __forceinline__ void init(const NarrowRows &g, uint32_t row)
    {
        bs = row / (uint32_t)g.K, k = row - bs * (uint32_t)g.K;
        b = bs / (uint32_t)g.S, s = bs - b * (uint32_t)g.S;
    }
    __device__ __forceinline__ void advance(const NarrowRows &g, uint32_t step)
    {
        k += step;
        while (k >= (uint32_t)g.K) {
            k -= (uint32_t)g.K, ++bs;
            if (++s == (uint32_t)g.S) s = 0, ++b;
        }
    }
};

// The C/8 threads that share a row are adjacent lanes (C/8 = 8, 16 or 32): lane j < 3 + D of the group loads channel j of
// the grouped row, the group exchanges the values by shuffle.  Groups of one warp may sit in different loop
// iterations, hence the group-local mask.
// The gather is split into phases so that the loads of the kRowUnroll rows in flight are independent of each other
// (index loads, then coordinate loads, then the exchange): one dependent chain per row made the first version
// latency-bound at 2 CTAs/SM.
template <class T>
__device__ __forceinline__ float narrow_component(const NarrowRows &g, int64_t i, uint32_t b, uint32_t bs, int sub)
{
    float v = 0.f;
    if (sub < g.D + 3 && i >= 0 && i < g.N) {
        const int32_t i32 = (int32_t)i;
        if (sub < g.D)
            v = g.feats[(int64_t)(int32_t)b * g.fsb + (int64_t)i32 * g.fsn + sub * g.fsc];
        else
            v = __fsub_rn(g.xyz[(int64_t)(int32_t)b * g.xsb + (int64_t)i32 * g.xsn + (sub - g.D) * g.xsc], g.new_xyz[bs * 3u + (uint32_t)(sub - g.D)]);
        v = Act<T>::round(v);
    }
    return v;
}
template <int CIN>
__device__ __forceinline__ void narrow_exchange(float v, unsigned gmask, int gbase, float (&a)[CIN])
{
#pragma unroll
    for (int k = 0; k < CIN; ++k) a[k] = __shfl_sync(gmask, v, gbase + k);
}

// SHARE: C/8 divides the warp, the cooperative gather above applies; otherwise every thread gathers for itself.
template <class T, int CIN, bool SHARE>
__global__ void __launch_bounds__(kEwThreads)
narrow_first_layer_kernel(NarrowRows g, const T *__restrict__ W, int ldw, int64_t M, int C, T *__restrict__ Z,
                          float *__restrict__ partials)
{
    const int cg = C >> 3;
    const int c00 = (threadIdx.x % cg) * 8;
    const int lane = threadIdx.x & 31, sub = lane % cg, gbase = lane - sub;
    const unsigned gmask = cg >= 32 ? 0xffffffffu : (((1u << cg) - 1u) << gbase);
    float2 w2[4][CIN];   // channel pairs (c00 + 2j, c00 + 2j + 1) x input channel
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < CIN; ++k)
            w2[j][k] = make_float2(Act<T>::to_float(W[(size_t)(c00 + 2 * j) * ldw + k]), Act<T>::to_float(W[(size_t)(c00 + 2 * j + 1) * ldw + k]));
    RowCursor cur;
    bool started = false;
    column_sums(M, C, partials, [&](int64_t r, int64_t stride, int nr, int c0, float(&s0)[8], float(&s1)[8]) {
        if (SHARE && !started) cur.init(g, (uint32_t)r), started = true;
        int64_t ii[kRowUnroll];
        uint32_t cb[kRowUnroll], cbs[kRowUnroll];
        float comp[kRowUnroll];
        if (SHARE) {
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (u < nr) {
                    ii[u] = g.idx[r + u * stride];
                    cb[u] = cur.b, cbs[u] = cur.bs;
                    cur.advance(g, (uint32_t)stride);
                }
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (u < nr) comp[u] = narrow_component<T>(g, ii[u], cb[u], cbs[u], sub);
        }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (u < nr) {
                const int64_t row = r + u * stride;
                float a[CIN];
                if (SHARE)
                    narrow_exchange<CIN>(comp[u], gmask, gbase, a);
                else
                    gather_narrow_row<T, CIN>(g, (uint32_t)row, a);
                float zf[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 acc = __fmul2_rn(make_float2(a[0], a[0]), w2[j][0]);
#pragma unroll
                    for (int k = 1; k < CIN; ++k) acc = __ffma2_rn(make_float2(a[k], a[k]), w2[j][k], acc);
                    zf[2 * j] = Act<T>::round(acc.x), zf[2 * j + 1] = Act<T>::round(acc.y);
                }
                Act<T>::st(Z + row * C + c0, Act<T>::pack(zf));
#pragma unroll
                for (int i = 0; i < 8; ++i) {   // statistics of the STORED values, like the fused GEMM epilogue
                    s0[i] += zf[i];
                    s1[i] = fmaf(zf[i], zf[i], s1[i]);
                }
            }
    });
}

constexpr int kNarrowBwdUnroll = 2;   // 4 rows in flight spill at 128 registers (2 CTAs/SM)
template <class T, int CIN, bool SHARE>
__global__ void __launch_bounds__(kEwThreads, 2)
narrow_first_layer_bwd_kernel(NarrowRows g, const T *__restrict__ dA, const T *__restrict__ Z,
                              const float *__restrict__ scale, const float *__restrict__ shift, const float *__restrict__ mean,
                              const float *__restrict__ rstd, const float *__restrict__ coef, int64_t M, int C, float *__restrict__ dW,
                              int ldw)
{
    extern __shared__ float sm_dw[];  // [C][CIN]
    for (int i = threadIdx.x; i < C * CIN; i += kEwThreads) sm_dw[i] = 0.f;
    __syncthreads();
    const RowWalk wk(C);
    if (wk.active) {
        const int c0 = wk.tc * 8;
        const int lane = threadIdx.x & 31, sub = lane % wk.cg, gbase = lane - sub;
        const unsigned gmask = wk.cg >= 32 ? 0xffffffffu : (((1u << wk.cg) - 1u) << gbase);
        ApplyConst k;
        k.load(scale, shift, mean, rstd, coef, C, c0);
        float2 acc[4][CIN];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < CIN; ++q) acc[j][q] = make_float2(0.f, 0.f);
        const int64_t stride = wk.rpp, r_end = slab_end(M, wk.rpp);
        int64_t r = slab_begin(M, wk.rpp) + wk.tr;
        RowCursor cur;
        if (SHARE && r < r_end) cur.init(g, (uint32_t)r);
        for (; r < r_end; r += kNarrowBwdUnroll * stride) {
            typename Act<T>::raw_t rz[kNarrowBwdUnroll], rd[kNarrowBwdUnroll];
            int64_t ii[kNarrowBwdUnroll];
            uint32_t cb[kNarrowBwdUnroll], cbs[kNarrowBwdUnroll];
            float comp[kNarrowBwdUnroll];
#pragma unroll
            for (int u = 0; u < kNarrowBwdUnroll; ++u)
                if (r + u * stride < r_end) {
                    rz[u] = Act<T>::ld(Z + (r + u * stride) * C + c0);
                    rd[u] = Act<T>::ld(dA + (r + u * stride) * C + c0);
                    if (SHARE) {
                        ii[u] = g.idx[r + u * stride];
                        cb[u] = cur.b, cbs[u] = cur.bs;
                        cur.advance(g, (uint32_t)stride);
                    }
                }
            if (SHARE) {
#pragma unroll
                for (int u = 0; u < kNarrowBwdUnroll; ++u)
                    if (r + u * stride < r_end) comp[u] = narrow_component<T>(g, ii[u], cb[u], cbs[u], sub);
            }
#pragma unroll
            for (int u = 0; u < kNarrowBwdUnroll; ++u)
                if (r + u * stride < r_end) {
                    float z[8], d[8], a[CIN];
                    Act<T>::unpack(rz[u], z);
                    Act<T>::unpack(rd[u], d);
                    if (SHARE)
                        narrow_exchange<CIN>(comp[u], gmask, gbase, a);
                    else
                        gather_narrow_row<T, CIN>(g, (uint32_t)(r + u * stride), a);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float dy = fmaf(z[i], k.sc[i], k.sh[i]) > 0.f ? d[i] : 0.f;
                        d[i] = fmaf(k.p[i], dy, fmaf(-k.w[i], z[i], k.e[i]));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // the tensor-core path feeds dZ to its weight-gradient GEMM in the storage type: same rounding here
                        const float2 dz = make_float2(Act<T>::round(d[2 * j]), Act<T>::round(d[2 * j + 1]));
#pragma unroll
                        for (int q = 0; q < CIN; ++q) acc[j][q] = __ffma2_rn(dz, make_float2(a[q], a[q]), acc[j][q]);
                    }
                }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int q = 0; q < CIN; ++q) {
                atomicAdd(&sm_dw[(c0 + 2 * j) * CIN + q], acc[j][q].x);
                atomicAdd(&sm_dw[(c0 + 2 * j + 1) * CIN + q], acc[j][q].y);
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * CIN; i += kEwThreads) atomicAdd(dW + (size_t)(i / CIN) * ldw + (i % CIN), sm_dw[i]);
}

static inline int row_blocks(int64_t rows, int C, int per_sm)
{
    const int rpp = kEwThreads / (C >> 3);
    int64_t want = (rows + rpp - 1) / rpp;
    const int cap = per_sm * sm_count();
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}
static inline int stat_parts(int64_t rows, int C) { return row_blocks(rows, C, kStatCtasPerSm); }
// dynamic shared memory of a column_sums kernel: one [2][C] slot per row lane of the CTA (16 KB when C/8 divides 256)
static inline size_t stat_smem(int C) { return (size_t)(kEwThreads / (C >> 3)) * 2 * C * sizeof(float); }

}  // namespace mpb

#define MPB_CHECK_C(C) MPB_REQUIRE((C) > 0 && (C) % 8 == 0 && (C) <= 2048, "C must be a positive multiple of 8, at most 2048")
#define MPB_CHECK_DT(dt) MPB_REQUIRE((dt) == 0 || (dt) == 1, "dtype must be 0 (bf16 activations) or 1 (fp32 activations)")
// Run the statement(s) with T bound to the activation storage type selected by `dt`.
#define MPB_DISPATCH_ACT(dt, ...)    \
    do {                             \
        if ((dt) == 0) {             \
            typedef __nv_bfloat16 T; \
            __VA_ARGS__;             \
        } else {                     \
            typedef float T;         \
            __VA_ARGS__;             \
        }                            \
    } while (0)

namespace mpb {
// conv weight fp32 [Cout, Cin] -> packed operand W [cout_p, cin_p] (zero padded; xyz_last: the 3 leading input channels move
// behind the feature channels, matching mpb_group_points_bf16's row layout) and its transpose Wt [cin_p, cout_p].
// MODE 0: bf16.  MODE 1: fp32 split for the 3xTF32 GEMMs: hi = tf32(w) into (Wp, Wt), lo = w - hi into (Wp_lo, Wt_lo).
// First half of the grid writes W, second half Wt; both with coalesced stores (the fp32 source is L2-resident).
template <int MODE>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float *__restrict__ W, int cout, int cin, int cout_p, int cin_p, int xyz_last, void *__restrict__ Wp,
                   void *__restrict__ Wt, float *__restrict__ Wp_lo, float *__restrict__ Wt_lo)
{
    const int total = cout_p * cin_p;
    const int half = (total + 255) / 256;
    const bool transposed = (int)blockIdx.x >= half;
    const int e = ((int)blockIdx.x - (transposed ? half : 0)) * 256 + threadIdx.x;
    if (e >= total) return;
    const int r = transposed ? e % cout_p : e / cin_p;   // output channel
    const int c = transposed ? e / cout_p : e % cin_p;   // packed input channel
    float v = 0.f;
    if (r < cout && c < cin) {
        const int src = (xyz_last && cin > 3) ? (c < cin - 3 ? c + 3 : c - (cin - 3)) : c;
        v = W[(size_t)r * cin + src];
    }
    if (MODE == 0) {
        static_cast<__nv_bfloat16 *>(transposed ? Wt : Wp)[e] = __float2bfloat16_rn(v);
    } else {
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
        const float hi = __uint_as_float(hb);
        float *lo = transposed ? Wt_lo : Wp_lo;
        static_cast<float *>(transposed ? Wt : Wp)[e] = lo ? hi : v;      // no low-part buffer: the weight stays unsplit
        if (lo) lo[e] = v - hi;
    }
}

template <class T>
static int resident_per_sm(const void *kernel)
{
    int q = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kernel, kEwThreads, 0) != cudaSuccess) q = 1;
    return q < 1 ? 1 : q;
}
}  // namespace mpb

extern "C" int mpb_pack_weight_bf16(const float *W, int cout, int cin, int cout_p, int cin_p, int xyz_last, void *Wp, void *Wt,
                                    void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(W && Wp && Wt && cout > 0 && cin > 0 && cout_p >= cout && cin_p >= cin, "bad argument");
    const int total = cout_p * cin_p;
    pack_weight_kernel<0><<<2 * ((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, cout, cin, cout_p, cin_p, xyz_last, Wp, Wt,
                                                                                     nullptr, nullptr);
    return check_launch("pack_weight_kernel");
}

extern "C" int mpb_pack_weight_tf32(const float *W, int cout, int cin, int cout_p, int cin_p, int xyz_last, float *Wp_hi, float *Wp_lo,
                                    float *Wt_hi, float *Wt_lo, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(W && Wp_hi && Wt_hi && cout > 0 && cin > 0 && cout_p >= cout && cin_p >= cin, "bad argument");
    MPB_REQUIRE((Wp_lo != nullptr) == (Wt_lo != nullptr), "Wp_lo / Wt_lo must come together (both null: unsplit fp32 copy)");
    const int total = cout_p * cin_p;
    pack_weight_kernel<1><<<2 * ((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, cout, cin, cout_p, cin_p, xyz_last, Wp_hi, Wt_hi,
                                                                                     Wp_lo, Wt_lo);
    return check_launch("pack_weight_kernel");
}

extern "C" int mpb_bn_stat_partials(int64_t rows, int C)
{
    if (rows <= 0 || C <= 0 || C % 8) return 0;
    return mpb::stat_parts(rows, C);
}

extern "C" int mpb_bn_colstats(int dtype, const void *Z, int64_t M, int C, float *partials, int nparts, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_CHECK_DT(dtype);
    MPB_REQUIRE(M > 0 && Z && partials && nparts == stat_parts(M, C), "bad argument");
    MPB_DISPATCH_ACT(dtype, colstats_kernel<T><<<nparts, kEwThreads, stat_smem(C), (cudaStream_t)stream>>>((const T *)Z, M, C, partials));
    return check_launch("colstats_kernel");
}

extern "C" int mpb_bn_finalize_f32(const float *partials, int nparts, int C, int C_valid, int64_t M, const float *bias,
                                   const float *gamma, const float *beta, float *running_mean, float *running_var, float momentum,
                                   float eps, float *scale, float *shift, float *mean, float *rstd, int64_t *num_batches_tracked,
                                   void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(partials && scale && shift && mean && rstd && C > 0 && nparts > 0 && M > 0, "bad argument");
    MPB_REQUIRE(C_valid >= 0 && C_valid <= C, "C_valid out of range");
    bn_finalize_kernel<<<(C + 7) / 8, kFinThreads, 0, (cudaStream_t)stream>>>(partials, nparts, C, C_valid, (double)M, bias, gamma, beta,
                                                                              running_mean, running_var, momentum, eps, scale, shift, mean,
                                                                              rstd, num_batches_tracked);
    return check_launch("bn_finalize_kernel");
}

extern "C" int mpb_bn_relu(int dtype, const void *Z, const float *scale, const float *shift, int64_t M, int C, void *A, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_CHECK_DT(dtype);
    MPB_REQUIRE(M >= 0 && Z && scale && shift && A, "bad argument");
    if (M == 0) return MPB_OK;
    MPB_DISPATCH_ACT(dtype, {
        static const int occ = resident_per_sm<T>((const void *)bn_relu_kernel<T>);
        bn_relu_kernel<T><<<row_blocks((M + kRowUnroll - 1) / kRowUnroll, C, occ), kEwThreads, 0, (cudaStream_t)stream>>>((const T *)Z, scale, shift, M, C,
                                                                                                                     (T *)A);
    });
    return check_launch("bn_relu_kernel");
}

extern "C" int mpb_bn_relu_max(int dtype, const void *Z, const float *scale, const float *shift, int64_t G, int K, int C, float *out,
                               int32_t *argmax, float *zmax, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_CHECK_DT(dtype);
    MPB_REQUIRE(G >= 0 && K > 0 && Z && scale && shift && out && argmax, "bad argument");
    if (G == 0) return MPB_OK;
    const int nseg = kWideThreads / (C >> 3);
    if (G <= 2 * sm_count() && nseg >= 2 && K >= 4 * nseg) {
        const size_t smem = (size_t)3 * nseg * C * sizeof(float);
        MPB_DISPATCH_ACT(dtype, {
            MPB_ENSURE_DYN_SMEM(bn_relu_max_wide_kernel<T>, smem);
            bn_relu_max_wide_kernel<T><<<(unsigned)G, kWideThreads, smem, (cudaStream_t)stream>>>((const T *)Z, scale, shift, G, K, C, out, argmax, zmax);
        });
        return check_launch("bn_relu_max_wide_kernel");
    }
    MPB_DISPATCH_ACT(dtype, {
        static const int occ = resident_per_sm<T>((const void *)bn_relu_max_kernel<T>);
        bn_relu_max_kernel<T><<<row_blocks(G, C, occ), kEwThreads, 0, (cudaStream_t)stream>>>((const T *)Z, scale, shift, G, K, C, out, argmax, zmax);
    });
    return check_launch("bn_relu_max_kernel");
}

extern "C" int mpb_bn_bwd_stats(int dtype, const void *dA, const float *dOut, const int32_t *argmax, const float *zmax, int K, const void *Z,
                                const float *scale, const float *shift, const float *mean, const float *rstd, int64_t M, int C,
                                float *partials, int nparts, float *pgo, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_CHECK_DT(dtype);
    MPB_REQUIRE(M > 0 && Z && scale && shift && mean && rstd && partials, "bad argument");
    MPB_REQUIRE((dA != nullptr) != (dOut != nullptr), "exactly one of dA (dense) / dOut (pooled) must be given");
    cudaStream_t st = (cudaStream_t)stream;
    if (dA) {
        MPB_REQUIRE(nparts == stat_parts(M, C), "nparts mismatch");
        MPB_DISPATCH_ACT(dtype, bwd_stats_dense_kernel<T><<<nparts, kEwThreads, stat_smem(C), st>>>((const T *)dA, (const T *)Z, scale, shift, mean, rstd, M,
                                                                                                  C, partials));
    } else {
        MPB_REQUIRE(argmax && K > 0 && M % K == 0 && nparts == stat_parts(M / K, C), "pooled: bad argmax/K/nparts");
        MPB_DISPATCH_ACT(dtype, bwd_stats_pooled_kernel<T><<<nparts, kEwThreads, stat_smem(C), st>>>(dOut, argmax, (const T *)Z, zmax, scale, shift, mean,
                                                                                                   rstd, M / K, K, C, partials, pgo));
    }
    return check_launch("bwd_stats kernel");
}

extern "C" int mpb_bn_bwd_finalize_f32(const float *partials, int nparts, int C, int C_valid, int64_t M, const float *gamma,
                                       const float *mean, const float *rstd, float *dgamma, float *dbeta, float *coef,
                                       float *clear, int64_t clear_count, float *negw_e, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(partials && mean && rstd && coef && C > 0 && nparts > 0 && M > 0, "bad argument");
    MPB_REQUIRE(C_valid >= 0 && C_valid <= C, "C_valid out of range");
    MPB_REQUIRE(clear_count >= 0 && (clear_count == 0 || clear), "clear buffer missing");
    MPB_REQUIRE(clear_count % 4 == 0 && (reinterpret_cast<uintptr_t>(clear) & 15) == 0, "clear buffer must be 16-byte granular");
    bwd_finalize_kernel<<<(C + 7) / 8, kFinThreads, 0, (cudaStream_t)stream>>>(partials, nparts, C, C_valid, (double)M, gamma, mean, rstd,
                                                                               dgamma, dbeta, coef, reinterpret_cast<float4 *>(clear),
                                                                               clear_count / 4, negw_e);
    return check_launch("bwd_finalize_kernel");
}

extern "C" int mpb_bn_bwd_apply(int dtype, const void *dA, const float *dOut, const int32_t *argmax, int K, const void *Z, const float *scale,
                                const float *shift, const float *mean, const float *rstd, const float *coef, int64_t M, int C,
                                void *dZ, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_CHECK_DT(dtype);
    MPB_REQUIRE(M > 0 && Z && scale && shift && mean && rstd && coef && dZ, "bad argument");
    MPB_REQUIRE((dA != nullptr) != (dOut != nullptr), "exactly one of dA (dense) / dOut (pooled) must be given");
    cudaStream_t st = (cudaStream_t)stream;
    if (dA) {
        MPB_DISPATCH_ACT(dtype, {
            static const int occ = resident_per_sm<T>((const void *)bwd_apply_dense_kernel<T>);
            bwd_apply_dense_kernel<T><<<row_blocks((M + kRowUnroll - 1) / kRowUnroll, C, occ), kEwThreads, 0, st>>>((const T *)dA, (const T *)Z, scale, shift,
                                                                                                               mean, rstd, coef, M, C, (T *)dZ);
        });
    } else {
        MPB_REQUIRE(argmax && K > 0 && M % K == 0, "pooled: bad argmax/K");
        MPB_DISPATCH_ACT(dtype, {
            static const int occ = resident_per_sm<T>((const void *)bwd_apply_pooled_kernel<T>);
            const int bx = row_blocks(M / K, C, occ), cap = occ * sm_count();
            int ks = (cap + bx - 1) / bx;                       // segments needed to fill one wave ...
            ks = ks > K / 8 ? K / 8 : ks;                       // ... of at least 8 rows each
            bwd_apply_pooled_kernel<T><<<dim3(bx, ks < 1 ? 1 : ks), kEwThreads, 0, st>>>(dOut, argmax, K, (const T *)Z, scale, shift, mean, rstd, coef,
                                                                                         M / K, C, (T *)dZ);
        });
    }
    return check_launch("bwd_apply kernel");
}

// Pooled form of mpb_bn_bwd_apply from the by-products of the statistics pass: pgo (mpb_bn_bwd_stats) and negw_e (mpb_bn_bwd_finalize_f32).
extern "C" int mpb_bn_bwd_apply_pooled(int dtype, const float *pgo, const int32_t *argmax, int K, const void *Z, const float *negw_e, int64_t M,
                                       int C, void *dZ, void *stream)
{
    using namespace mpb;
    MPB_CHECK_C(C);
    MPB_CHECK_DT(dtype);
    MPB_REQUIRE(M > 0 && K > 0 && M % K == 0 && pgo && argmax && Z && negw_e && dZ, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    MPB_DISPATCH_ACT(dtype, {
        static const int occ = resident_per_sm<T>((const void *)bwd_apply_pooled_lean_kernel<T>);
        const int bx = row_blocks(M / K, C, occ), cap = occ * sm_count();
        int ks = (cap + bx - 1) / bx;                       // segments needed to fill one wave ...
        ks = ks > K / 8 ? K / 8 : ks;                       // ... of at least 8 rows each
        bwd_apply_pooled_lean_kernel<T><<<dim3(bx, ks < 1 ? 1 : ks), kEwThreads, 0, st>>>(pgo, argmax, K, (const T *)Z, negw_e, M / K, C, (T *)dZ);
    });
    return check_launch("bwd_apply_pooled_lean_kernel");
}

namespace mpb {
static inline bool fits31(int64_t v) { return v >= 0 && v < ((int64_t)1 << 31); }
static inline NarrowRows narrow_rows(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *feats, int64_t fsb,
                                     int64_t fsn, int64_t fsc, const float *new_xyz, const int64_t *idx, int N, int S, int K, int D)
{
    NarrowRows g;
    g.xyz = xyz, g.xsb = (int32_t)xsb, g.xsn = (int32_t)xsn, g.xsc = (int32_t)xsc;
    g.feats = feats, g.fsb = (int32_t)fsb, g.fsn = (int32_t)fsn, g.fsc = (int32_t)fsc;
    g.new_xyz = new_xyz, g.idx = idx, g.N = N, g.S = S, g.K = K, g.D = D;
    return g;
}

template <class T, int CIN>
static void launch_narrow_fwd(bool share, const NarrowRows &g, const void *W, int ldw, int64_t M, int C, void *Z, float *partials, int nparts,
                              cudaStream_t st)
{
    const size_t smem = stat_smem(C);
    if (share)
        narrow_first_layer_kernel<T, CIN, true><<<nparts, kEwThreads, smem, st>>>(g, (const T *)W, ldw, M, C, (T *)Z, partials);
    else
        narrow_first_layer_kernel<T, CIN, false><<<nparts, kEwThreads, smem, st>>>(g, (const T *)W, ldw, M, C, (T *)Z, partials);
}
template <class T, int CIN, bool SHARE>
static void launch_narrow_bwd2(const NarrowRows &g, const void *dA, const void *Z, const float *scale, const float *shift, const float *mean,
                               const float *rstd, const float *coef, int64_t M, int C, float *dW, int ldw, cudaStream_t st)
{
    static const int occ = resident_per_sm<T>((const void *)narrow_first_layer_bwd_kernel<T, CIN, SHARE>);
    narrow_first_layer_bwd_kernel<T, CIN, SHARE><<<row_blocks((M + kNarrowBwdUnroll - 1) / kNarrowBwdUnroll, C, occ), kEwThreads,
                                                   (size_t)C * CIN * sizeof(float), st>>>(g, (const T *)dA, (const T *)Z, scale, shift, mean, rstd,
                                                                                          coef, M, C, dW, ldw);
}
template <class T, int CIN>
static void launch_narrow_bwd(bool share, const NarrowRows &g, const void *dA, const void *Z, const float *scale, const float *shift,
                              const float *mean, const float *rstd, const float *coef, int64_t M, int C, float *dW, int ldw, cudaStream_t st)
{
    if (share)
        launch_narrow_bwd2<T, CIN, true>(g, dA, Z, scale, shift, mean, rstd, coef, M, C, dW, ldw, st);
    else
        launch_narrow_bwd2<T, CIN, false>(g, dA, Z, scale, shift, mean, rstd, coef, M, C, dW, ldw, st);
}
}  // namespace mpb

#define MPB_NARROW_CHECKS()                                                                                    \
    MPB_CHECK_C(C);                                                                                            \
    MPB_CHECK_DT(dtype);                                                                                       \
    MPB_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0 && D >= 0 && D + 3 <= 8, "bad size (needs 3 + D <= 8)");       \
    MPB_REQUIRE(xyz && new_xyz && idx && (D == 0 || feats), "null pointer");                                   \
    MPB_REQUIRE(ldw >= D + 3, "ldw must cover the 3 + D input channels");                                      \
    MPB_REQUIRE(mpb::fits31(xsb) && mpb::fits31(xsn) && mpb::fits31(xsc) && mpb::fits31(fsb) && mpb::fits31(fsn) && \
                    mpb::fits31(fsc) && (int64_t)B * S * 3 < ((int64_t)1 << 31),                                \
                "strides must fit 31 bits");                                                                   \
    MPB_REQUIRE((int64_t)B * S * K < ((int64_t)1 << 31), "row count exceeds 32-bit indexing")

// W: the first layer's packed weight [C, ldw] in the activation storage type (bf16, or fp32 = the UNSPLIT weight: the
// CUDA-core FMAs of this kernel are exact fp32, there is no tensor-core rounding to compensate).
extern "C" int mpb_sa_first_layer(int dtype, const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *feats, int64_t fsb,
                                  int64_t fsn, int64_t fsc, const float *new_xyz, const int64_t *idx, int B, int N, int S, int K,
                                  int D, const void *W, int ldw, int C, void *Z, float *partials, int nparts, void *stream)
{
    using namespace mpb;
    MPB_NARROW_CHECKS();
    const int64_t M = (int64_t)B * S * K;
    MPB_REQUIRE(W && Z && partials && nparts == stat_parts(M, C), "bad argument");
    const NarrowRows g = narrow_rows(xyz, xsb, xsn, xsc, feats, fsb, fsn, fsc, new_xyz, idx, N, S, K, D);
    cudaStream_t st = (cudaStream_t)stream;
    const bool share = (32 % (C >> 3)) == 0;
    if (D + 3 <= 3)
        MPB_DISPATCH_ACT(dtype, (launch_narrow_fwd<T, 3>(share, g, W, ldw, M, C, Z, partials, nparts, st)));
    else if (D + 3 <= 6)
        MPB_DISPATCH_ACT(dtype, (launch_narrow_fwd<T, 6>(share, g, W, ldw, M, C, Z, partials, nparts, st)));
    else
        MPB_DISPATCH_ACT(dtype, (launch_narrow_fwd<T, 8>(share, g, W, ldw, M, C, Z, partials, nparts, st)));
    return check_launch("narrow_first_layer_kernel");
}

// dW [C, ldw] fp32 is ACCUMULATED into (caller zero-fills; mpb_bn_bwd_finalize_f32 does it as its side job).
extern "C" int mpb_sa_first_layer_bwd(int dtype, const void *dA, const void *Z, const float *scale, const float *shift, const float *mean,
                                      const float *rstd, const float *coef, const float *xyz, int64_t xsb, int64_t xsn,
                                      int64_t xsc, const float *feats, int64_t fsb, int64_t fsn, int64_t fsc,
                                      const float *new_xyz, const int64_t *idx, int B, int N, int S, int K, int D, int C,
                                      float *dW, int ldw, void *stream)
{
    using namespace mpb;
    MPB_NARROW_CHECKS();
    MPB_REQUIRE(dA && Z && scale && shift && mean && rstd && coef && dW, "bad argument");
    const int64_t M = (int64_t)B * S * K;
    const NarrowRows g = narrow_rows(xyz, xsb, xsn, xsc, feats, fsb, fsn, fsc, new_xyz, idx, N, S, K, D);
    cudaStream_t st = (cudaStream_t)stream;
    const bool share = (32 % (C >> 3)) == 0;
    if (D + 3 <= 3)
        MPB_DISPATCH_ACT(dtype, (launch_narrow_bwd<T, 3>(share, g, dA, Z, scale, shift, mean, rstd, coef, M, C, dW, ldw, st)));
    else if (D + 3 <= 6)
        MPB_DISPATCH_ACT(dtype, (launch_narrow_bwd<T, 6>(share, g, dA, Z, scale, shift, mean, rstd, coef, M, C, dW, ldw, st)));
    else
        MPB_DISPATCH_ACT(dtype, (launch_narrow_bwd<T, 8>(share, g, dA, Z, scale, shift, mean, rstd, coef, M, C, dW, ldw, st)));
    return check_launch("narrow_first_layer_bwd_kernel");
}
