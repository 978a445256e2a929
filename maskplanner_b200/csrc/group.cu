// Gather / grouping kernels for sm_100a.
// Reference: models/pointnet2_utils.py:45-62 (index_points), :133-141 (gather + centre + concat
// inside sample_and_group).  The reference issues four advanced-indexing gathers, a subtract and a
// cat, each materialised; mpb_group_points_f32 produces the concatenated [B,S,K,3+D] tensor (with
// optional zero K-padding for the GEMM that consumes it) in one pass.
//
// These are pure data movement (HBM/L2-bound): one thread per output element, channel index
// fastest so that both the gathered row read and the output write are coalesced.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mpb {

// Forward gathers: `lpr` (a power of two, <= 32) consecutive lanes share one output row -- the row's index
// arithmetic is done once per lane, the lanes then stride over the row's columns, so both the gathered read and
// the write are coalesced and no per-element 64-bit division is left.
__global__ void index_points_kernel(const float *__restrict__ pts, int64_t sb, int64_t sn, int64_t sc, int N, int C,
                                    const int64_t *__restrict__ idx, int64_t M, int64_t rows, int lpr_shift,
                                    float *__restrict__ out)
{
    const int lpr = 1 << lpr_shift;
    const int sub = threadIdx.x & (lpr - 1);
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> lpr_shift;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> lpr_shift; row < rows; row += stride) {
        const int64_t b = row / M;
        const int64_t i = idx[row];
        const bool ok = i >= 0 && i < N;
        const float *src = pts + b * sb + (ok ? i : 0) * sn;
        float *dst = out + row * C;
        for (int c = sub; c < C; c += lpr) dst[c] = ok ? src[(int64_t)c * sc] : 0.f;
    }
}

__global__ void index_points_bwd_kernel(const float *__restrict__ go, const int64_t *__restrict__ idx, int N, int C,
                                        int64_t M, int64_t total, float *__restrict__ gp)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const int64_t row = e / C;
        const int64_t b = row / M;
        const int64_t i = idx[row];
        if (i >= 0 && i < N) atomicAdd(gp + (b * N + i) * C + c, go[e]);
    }
}

__device__ __forceinline__ void store_as(float *p, float v) { *p = v; }
__device__ __forceinline__ void store_as(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ float load_as(const float *p) { return *p; }
__device__ __forceinline__ float load_as(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

template <typename OutT>
__global__ void group_points_kernel(const float *__restrict__ xyz, int64_t xsb, int64_t xsn, int64_t xsc,
                                    const float *__restrict__ feats, int64_t fsb, int64_t fsn, int64_t fsc,
                                    const float *__restrict__ new_xyz, const int64_t *__restrict__ idx, int N, int S,
                                    int K, int D, int ldo, int64_t rows, int lpr_shift, OutT *__restrict__ out)
{
    const int lpr = 1 << lpr_shift;
    const int sub = threadIdx.x & (lpr - 1);
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> lpr_shift;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> lpr_shift; row < rows; row += stride) {
        const int64_t bs = row / K;    // b*S + s ; row = (b*S + s)*K + k
        const int64_t b = bs / S;
        const int64_t i = idx[row];
        const bool ok = i >= 0 && i < N;
        const float *px = xyz + b * xsb + (ok ? i : 0) * xsn;
        const float *pf = feats ? feats + b * fsb + (ok ? i : 0) * fsn : nullptr;
        OutT *dst = out + row * ldo;
        for (int c = sub; c < ldo; c += lpr) {
            float v = 0.f;
            if (ok) {
                if (c < 3)
                    v = __fsub_rn(px[c * xsc], new_xyz[bs * 3 + c]);      // :134
                else if (c < 3 + D)
                    v = pf[(int64_t)(c - 3) * fsc];                       // :137-138
            }
            store_as(dst + c, v);
        }
    }
}

template <typename InT>
__global__ void group_points_bwd_kernel(const InT *__restrict__ go, int ldo, const int64_t *__restrict__ idx, int N,
                                        int S, int K, int D, int64_t total, float *__restrict__ gfeats,
                                        float *__restrict__ gxyz, float *__restrict__ gnew)
{
    const int W = 3 + D;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(e % W);
        const int64_t row = e / W;
        const int64_t bs = row / K;
        const int64_t b = bs / S;
        const int64_t i = idx[row];
        if (i < 0 || i >= N) continue;
        const float g = load_as(go + row * ldo + c);
        if (c < 3) {
            if (gxyz) atomicAdd(gxyz + (b * N + i) * 3 + c, g);
            if (gnew) atomicAdd(gnew + bs * 3 + c, -g);
        } else if (gfeats) {
            atomicAdd(gfeats + (b * N + i) * D + (c - 3), g);
        }
    }
}

// smallest power of two >= min(width, 32), as a shift
static inline int lanes_per_row_shift(int width)
{
    int s = 0;
    while ((1 << s) < width && s < 5) ++s;
    return s;
}

static inline unsigned grid_for(int64_t total, int threads)
{
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace mpb

extern "C" int mpb_index_points_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B, int N, int C,
                                    const int64_t *idx, int64_t M, float *out, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && C >= 0 && M >= 0, "negative size");
    const int64_t total = (int64_t)B * M * C;
    if (total == 0) return MPB_OK;
    MPB_REQUIRE(points && idx && out, "null pointer");
    const int shift = lanes_per_row_shift(C);
    index_points_kernel<<<grid_for(((int64_t)B * M) << shift, 256), 256, 0, (cudaStream_t)stream>>>(points, sb, sn, sc, N, C, idx, M,
                                                                                                  (int64_t)B * M, shift, out);
    return check_launch("index_points_kernel");
}

extern "C" int mpb_index_points_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int C, int64_t M,
                                        float *grad_points, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && C >= 0 && M >= 0, "negative size");
    const int64_t total = (int64_t)B * M * C;
    if (total == 0) return MPB_OK;
    MPB_REQUIRE(grad_out && idx && grad_points, "null pointer");
    index_points_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, idx, N, C, M, total, grad_points);
    return check_launch("index_points_bwd_kernel");
}

extern "C" int mpb_group_points_f32(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *feats,
                                    int64_t fsb, int64_t fsn, int64_t fsc, const float *new_xyz, const int64_t *idx,
                                    int B, int N, int S, int K, int D, int ldo, float *out, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && S >= 0 && K >= 0 && D >= 0, "negative size");
    MPB_REQUIRE(ldo >= 3 + D, "ldo < 3 + D");
    const int64_t total = (int64_t)B * S * K * ldo;
    if (total == 0) return MPB_OK;
    MPB_REQUIRE(xyz && new_xyz && idx && out, "null pointer");
    MPB_REQUIRE(D == 0 || feats, "feats is null but D > 0");
    const int shift = lanes_per_row_shift(ldo);
    const int64_t rows = (int64_t)B * S * K;
    group_points_kernel<float><<<grid_for(rows << shift, 256), 256, 0, (cudaStream_t)stream>>>(xyz, xsb, xsn, xsc, feats, fsb, fsn, fsc,
                                                                                             new_xyz, idx, N, S, K, D, ldo, rows, shift, out);
    return check_launch("group_points_kernel");
}

extern "C" int mpb_group_points_bwd_f32(const float *grad_out, int ldo, const int64_t *idx, int B, int N, int S, int K,
                                        int D, float *grad_feats, float *grad_xyz, float *grad_new_xyz, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && S >= 0 && K >= 0 && D >= 0, "negative size");
    MPB_REQUIRE(ldo >= 3 + D, "ldo < 3 + D");
    const int64_t total = (int64_t)B * S * K * (3 + D);
    if (total == 0) return MPB_OK;
    MPB_REQUIRE(grad_out && idx, "null pointer");
    group_points_bwd_kernel<float><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(grad_out, ldo, idx, N, S, K, D, total,
                                                                                   grad_feats, grad_xyz, grad_new_xyz);
    return check_launch("group_points_bwd_kernel");
}

// bf16 rows for the tensor-core shared MLP (sa_gemm.cu): the GEMM's A operand, written directly.
// Column order is FEATURES FIRST: [feats(0..D-1), centred xyz(D..D+2), zero padding up to ldo] so that the
// D feature channels of a gathered point are 16-byte aligned vector copies (the host permutes the first
// layer's weight columns to match).  One thread = one row x 8 columns: a single index computation and a
// single 16-byte store per 8 outputs.
namespace mpb {

// IT = uint32_t whenever the element count fits (always, in practice): the per-item row / batch decomposition is three
// integer divisions, and 64-bit ones (~100 instructions each) made the kernel instruction-bound at 1.2-1.6 TB/s.
template <typename IT>
__global__ void __launch_bounds__(256)
group_rows_bf16_kernel(const float *__restrict__ xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *__restrict__ feats,
                       int64_t fsb, int64_t fsn, int64_t fsc, const float *__restrict__ new_xyz, const int64_t *__restrict__ idx,
                       int N, int S, int K, int D, int ldo, int64_t total_vec, int vec_ok, __nv_bfloat16 *__restrict__ out,
                       FastDiv dnv, FastDiv dk, FastDiv ds)
{
    constexpr bool FAST = sizeof(IT) == 4;   // multiply-shift division (common.cuh) on the 32-bit path
    const IT nv = (IT)(ldo >> 3), total = (IT)total_vec, step = (IT)gridDim.x * blockDim.x;
    for (IT v = (IT)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += step) {
        const IT row = FAST ? (IT)dnv.div((uint32_t)v) : v / nv;
        const int c0 = (int)(v - row * nv) * 8;
        float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (c0 < D + 3) {
            const int64_t i = idx[row];
            if (i >= 0 && i < N) {
                const IT bs = FAST ? (IT)dk.div((uint32_t)row) : row / (IT)K;
                const int64_t b = (int64_t)(FAST ? (IT)ds.div((uint32_t)bs) : bs / (IT)S);
                if (vec_ok && c0 + 8 <= D) {
                    const float4 *src = reinterpret_cast<const float4 *>(feats + b * fsb + i * fsn + c0);
                    const float4 lo = src[0], hi = src[1];
                    f[0] = lo.x, f[1] = lo.y, f[2] = lo.z, f[3] = lo.w, f[4] = hi.x, f[5] = hi.y, f[6] = hi.z, f[7] = hi.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = c0 + e;
                        if (c < D)
                            f[e] = feats[b * fsb + i * fsn + (int64_t)c * fsc];
                        else if (c < D + 3)
                            f[e] = __fsub_rn(xyz[b * xsb + i * xsn + (c - D) * xsc], new_xyz[bs * 3 + (c - D)]);
                    }
                }
            }
        }
        uint4 pk;
        __nv_bfloat162 *pp = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) pp[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
        reinterpret_cast<uint4 *>(out)[v] = pk;
    }
}

// grad_feats[b, idx[row], c] += grad_rows[row, c] for c < D: one thread = one row x 4 channels, one 16-byte
// vector reduction (RED.128) per 4 channels when D % 4 == 0.
template <typename IT>
__global__ void __launch_bounds__(256)
group_rows_bwd_bf16_kernel(const __nv_bfloat16 *__restrict__ go, int ldo, const int64_t *__restrict__ idx, int N, int S, int K, int D,
                           int64_t total_vec, float *__restrict__ gfeats, FastDiv dnv, FastDiv dsk)
{
    constexpr bool FAST = sizeof(IT) == 4;
    const IT nv = (IT)((D + 3) >> 2), total = (IT)total_vec, step = (IT)gridDim.x * blockDim.x, sk = (IT)S * (IT)K;
    for (IT v = (IT)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += step) {
        const IT row = FAST ? (IT)dnv.div((uint32_t)v) : v / nv;
        const int c0 = (int)(v - row * nv) * 4;
        const int64_t i = idx[row];
        if (i < 0 || i >= N) continue;
        const int64_t b = (int64_t)(FAST ? (IT)dsk.div((uint32_t)row) : row / sk);
        const uint2 raw = *reinterpret_cast<const uint2 *>(go + (int64_t)row * ldo + c0);
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.x));
        const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&raw.y));
        float *dst = gfeats + (b * N + i) * D + c0;
        if ((D & 3) == 0) {
            atomicAdd(reinterpret_cast<float4 *>(dst), make_float4(a.x, a.y, c.x, c.y));
        } else {
            const float vals[4] = {a.x, a.y, c.x, c.y};
            for (int e = 0; e < 4 && c0 + e < D; ++e) atomicAdd(dst + e, vals[e]);
        }
    }
}

}  // namespace mpb

extern "C" int mpb_group_points_bf16(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *feats,
                                     int64_t fsb, int64_t fsn, int64_t fsc, const float *new_xyz, const int64_t *idx, int B,
                                     int N, int S, int K, int D, int ldo, void *out, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && S >= 0 && K >= 0 && D >= 0, "negative size");
    MPB_REQUIRE(ldo >= 3 + D && ldo % 8 == 0, "ldo must be a multiple of 8 and >= 3 + D");
    const int64_t total_vec = (int64_t)B * S * K * (ldo / 8);
    if (total_vec == 0) return MPB_OK;
    MPB_REQUIRE(xyz && new_xyz && idx && out, "null pointer");
    MPB_REQUIRE(D == 0 || feats, "feats is null but D > 0");
    const int vec_ok = D >= 8 && fsc == 1 && (fsb % 4 == 0) && (fsn % 4 == 0) && (((uintptr_t)feats & 15) == 0);
    const FastDiv dnv = make_fastdiv((uint32_t)(ldo / 8)), dk = make_fastdiv((uint32_t)(K > 0 ? K : 1)), ds = make_fastdiv((uint32_t)(S > 0 ? S : 1));
    // 32-bit item arithmetic needs total + one grid stride to stay below 2^32
    if (total_vec < (int64_t)3 << 30)
        group_rows_bf16_kernel<uint32_t><<<grid_for(total_vec, 256), 256, 0, (cudaStream_t)stream>>>(
            xyz, xsb, xsn, xsc, feats, fsb, fsn, fsc, new_xyz, idx, N, S, K, D, ldo, total_vec, vec_ok, (__nv_bfloat16 *)out, dnv, dk, ds);
    else
        group_rows_bf16_kernel<int64_t><<<grid_for(total_vec, 256), 256, 0, (cudaStream_t)stream>>>(
            xyz, xsb, xsn, xsc, feats, fsb, fsn, fsc, new_xyz, idx, N, S, K, D, ldo, total_vec, vec_ok, (__nv_bfloat16 *)out, dnv, dk, ds);
    return check_launch("group_rows_bf16_kernel");
}

extern "C" int mpb_group_points_bwd_bf16(const void *grad_out, int ldo, const int64_t *idx, int B, int N, int S, int K, int D,
                                         float *grad_feats, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && S >= 0 && K >= 0 && D >= 0, "negative size");
    MPB_REQUIRE(ldo >= 3 + D && ldo % 8 == 0, "ldo must be a multiple of 8 and >= 3 + D");
    const int64_t total_vec = (int64_t)B * S * K * ((D + 3) / 4);
    if (total_vec == 0) return MPB_OK;
    MPB_REQUIRE(grad_out && idx && grad_feats, "null pointer");
    MPB_REQUIRE(((uintptr_t)grad_feats & 15) == 0, "grad_feats must be 16-byte aligned");
    const uint64_t sk64 = (uint64_t)(S > 0 ? S : 1) * (uint64_t)(K > 0 ? K : 1);
    const FastDiv dnv = make_fastdiv((uint32_t)((D + 3) / 4)), dsk = make_fastdiv((uint32_t)(sk64 < (1ull << 31) ? sk64 : 1));
    if (total_vec < (int64_t)3 << 30 && sk64 < (1ull << 31))
        group_rows_bwd_bf16_kernel<uint32_t><<<grid_for(total_vec, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16 *)grad_out, ldo, idx, N, S, K, D, total_vec, grad_feats, dnv, dsk);
    else
        group_rows_bwd_bf16_kernel<int64_t><<<grid_for(total_vec, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16 *)grad_out, ldo, idx, N, S, K, D, total_vec, grad_feats, dnv, dsk);
    return check_launch("group_rows_bwd_bf16_kernel");
}
