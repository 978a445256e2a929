// Shared-MLP GEMMs of PointNetSetAbstraction on 5th-gen tensor cores (sm_100a).
// Reference: models/pointnet2_utils.py:210-212 -- `F.relu(bn(conv(new_points)))` with a 1x1 Conv2d over the grouped
// tensor [B, C, K, S] is a row-wise GEMM  Z[M, Cout] = A[M, Cin] * W[Cout, Cin]^T  with M = B*S*K rows, followed by a
// per-channel affine + ReLU that this file applies INSIDE the consumer GEMM's operand path instead of a separate pass.
//
//   gemm_tn_kernel : C[M,N] = f(A)[M,K] * B[N,K]^T       forward (B = W) and dgrad (A = dZ, B = W^T)
//   wgrad_kernel   : dW[N,K] = dZ[M,N]^T * f(A)[M,K]     contraction over the M rows, split across CTAs,
//                                                         per-split partial tiles summed in a fixed order (deterministic)
//
// Arithmetic (template parameter DT):
//   DT_BF16    bf16 operands (tcgen05 kind::f16), fp32 accumulation in TMEM, bf16 output          tolerance 1e-2
//   DT_TF32    fp32 operands read as TF32 (kind::tf32, one pass: what cuDNN does for the reference on a GPU)
//   DT_TF32X3  fp32 operands split hi + lo in shared memory, three TF32 products hi*hi + hi*lo + lo*hi,
//              fp32 accumulation: error ~2^-21 relative per product, the reference-precision mode  tolerance 1e-4
//
// Common structure: operands staged by TMA (SWIZZLE_128B) through a multi-stage mbarrier ring, tcgen05.mma
// (cta_group::1, UMMA 128 x N x {16 bf16 | 8 tf32}) issued by one elected thread, epilogue warps drain TMEM with
// tcgen05.ld.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 (+ 8..11 in gemm_tn) =
// epilogue, and -- new -- a TRANSFORM warpgroup that rewrites the landed A tile in place between the TMA's
// mbarrier and the MMA issue:
//     f(a) = relu(scale[k] * a + shift[k])      the previous layer's BatchNorm + ReLU (XFORM), so the normalised
//                                               activation tensor is never written to or read from HBM;
//     a -> (tf32(a), a - tf32(a))               the hi/lo split of DT_TF32X3.
// A third transform (XFORM 2, bf16, off by default) rebuilds dZ of the max-pooled layer from its stored pre-activation (sparse
// gather -> dense affine pass -> scatter, arg-max / p*dY slices delivered by TMA into the stage): mpb_sa_gemm_tn_pool /
// mpb_sa_gemm_wgrad_pool.  The two transform warpgroups of gemm_tn take alternate k-blocks but BOTH wait on full[stage] for every
// k-block: mbarrier waits test a phase parity, and a warpgroup that skipped a phase of a barrier can be told "done" one phase early
// (the deadlock of DESIGN.md section 9.5).  N = 2 x 128 with the tensor-core statistics runs as ONE paired tile (GemmTnArgs::pair).
// Epilogue options of gemm_tn: plain store (one TMA store per 128-row x 128-byte block through a swizzled staging
// tile); + per-column sum / sum of squares of the stored values (forward BatchNorm statistics); + per-column
// sum dY / sum dY*z with dY = C * [zscale*z + zshift > 0] against a TMA-loaded tile of the previous layer's
// pre-activation z (the BatchNorm-backward statistics of the layer below, fused into the dgrad GEMM).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace mpb {

using namespace tc;

enum { DT_BF16 = 0, DT_TF32 = 1, DT_TF32X3 = 2 };

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// Row-major [rows, cols] matrix of bf16 (esz = 2) or fp32 (esz = 4), row stride = ld elements;
// box = one 128-byte swizzle row (64 bf16 / 32 fp32 columns) x box_rows, SWIZZLE_128B -- or, atom32, the 32-byte-atom
// variant (32-byte chunks XOR row & 3) that MN-major 32-bit tensor-core operands require.
static int make_map(CUtensorMap *map, int esz, const void *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool atom32 = false)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return MPB_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * (cuuint64_t)esz};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(ptr),
                     gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) esz=%d rows=%lld cols=%lld ld=%lld box_rows=%d ptr=%p", (int)r, esz, (long long)rows,
                  (long long)cols, (long long)ld, box_rows, ptr);
        return MPB_ERR_CUDA;
    }
    return MPB_OK;
}

// Row-major [rows, cols] table of 32-bit words, no swizzle: box = box_cols x box_rows (the arg-max / pgo tables of the pooled transform).
static int make_map_plain32(CUtensorMap *map, const void *ptr, int64_t rows, int64_t cols, int box_cols, int box_rows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return MPB_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)cols * 4u};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (plain) failed (%d) rows=%lld cols=%lld box=%dx%d ptr=%p", (int)r, (long long)rows, (long long)cols, box_cols,
                  box_rows, ptr);
        return MPB_ERR_CUDA;
    }
    return MPB_OK;
}

constexpr int kTileM = 128;
constexpr int kPoolStageBytes = 4096;   // per stage: arg-max and pgo slices of the box's pooling groups (2 x <= 2 KB)
constexpr int kABytes = kTileM * 128;   // one 128-row x 128-byte operand / staging tile
constexpr uint32_t kTmemCols = 512;
constexpr int kMaxVec = 1024;           // longest per-channel vector (scale/shift) kept in shared memory

struct GemmSmemTail {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t ready[8];      // transform warps -> MMA issuer
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint64_t zfull[2][2];   // [epilogue warpgroup][z staging buffer]
    uint64_t sdone[2][2];   // [epilogue warpgroup][staging buffer]: statistic MMAs that read the buffer have completed
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u)
{
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
}

// Column sums of a 32 x 32 block held one ROW per lane (v[c] = this lane's value in column c): a transpose-reduce
// butterfly.  Each step halves the live registers -- a lane keeps the half of the columns selected by one bit of its
// lane index and receives the other lanes' contributions for that half -- so after 16+8+4+2+1 = 31 shuffles lane L
// holds the sum over all 32 rows of column L.  No shared memory, no atomics, fixed order (bitwise reproducible).
__device__ __forceinline__ float warp_column_sum_32x32(float (&v)[32], int lane)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? v[i] : v[i + off];
            const float keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi)
{
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));   // max(x, 0) folded into the rounding conversion
    return d;
}

// Column-owner form of the same transform: the thread owns ONE logical 16-byte chunk `c` (8 bf16 / 4 fp32 channels) of
// the rows r0, r0 + RSTEP, ... (NR of them) of a box of 128-byte swizzled rows, so its per-channel scale / shift sit in
// registers for the whole box (the row-owner form re-read them from shared memory for every chunk: four uniform
// 16-byte loads per data load, 3/4 of the kernel's shared-memory wavefronts in the ncu capture) and, RSTEP being a
// multiple of 8, the swizzled chunk position is a per-thread constant.  A quarter warp (8 lanes = the 8 chunks of one
// row) touches one whole 128-byte row per access: conflict-free.  Rows >= rows_valid (TMA zero fill past M) stay zero.
template <int DT, bool XFORM, bool ATOM32, int NR, int RSTEP, bool RELU = true>
__device__ __forceinline__ void transform_cols(uint8_t *box, uint8_t *box_lo, int c, int r0, const float *sc, const float *sh, int rows_valid)
{
    static_assert(RSTEP % 8 == 0, "the swizzle key must not change along the thread's rows");
    const int r7 = r0 & 7;
    const int off = ATOM32 ? (((((c >> 1) ^ (r7 & 3)) << 1) | (c & 1)) << 4) : ((c ^ r7) << 4);
    uint8_t *p0 = box + r0 * 128 + off;
    uint4 raw[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) raw[j] = *reinterpret_cast<const uint4 *>(p0 + j * RSTEP * 128);
    if (DT == DT_BF16) {
        float2 s2[4], h2[4];
        if (XFORM) {
            const float4 s0 = *reinterpret_cast<const float4 *>(sc + c * 8), s1 = *reinterpret_cast<const float4 *>(sc + c * 8 + 4);
            const float4 h0 = *reinterpret_cast<const float4 *>(sh + c * 8), h1 = *reinterpret_cast<const float4 *>(sh + c * 8 + 4);
            s2[0] = make_float2(s0.x, s0.y), s2[1] = make_float2(s0.z, s0.w), s2[2] = make_float2(s1.x, s1.y), s2[3] = make_float2(s1.z, s1.w);
            h2[0] = make_float2(h0.x, h0.y), h2[1] = make_float2(h0.z, h0.w), h2[2] = make_float2(h1.x, h1.y), h2[3] = make_float2(h1.z, h1.w);
        }
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const uint32_t w[4] = {raw[j].x, raw[j].y, raw[j].z, raw[j].w};
            uint32_t o[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float2 x = unpack_bf16(w[u]);
                if (XFORM) {
                    x = __ffma2_rn(x, s2[u], h2[u]);
                    o[u] = RELU ? pack_bf16_relu(x.x, x.y) : pack_bf16(x.x, x.y);
                } else {
                    o[u] = w[u];
                }
            }
            const bool valid = r0 + j * RSTEP < rows_valid;
            *reinterpret_cast<uint4 *>(p0 + j * RSTEP * 128) = valid ? make_uint4(o[0], o[1], o[2], o[3]) : make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (XFORM) s4 = *reinterpret_cast<const float4 *>(sc + c * 4), h4 = *reinterpret_cast<const float4 *>(sh + c * 4);
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            float f[4] = {__uint_as_float(raw[j].x), __uint_as_float(raw[j].y), __uint_as_float(raw[j].z), __uint_as_float(raw[j].w)};
            if (XFORM) {
                f[0] = fmaxf(fmaf(f[0], s4.x, h4.x), 0.f), f[1] = fmaxf(fmaf(f[1], s4.y, h4.y), 0.f);
                f[2] = fmaxf(fmaf(f[2], s4.z, h4.z), 0.f), f[3] = fmaxf(fmaf(f[3], s4.w, h4.w), 0.f);
            }
            if (!(r0 + j * RSTEP < rows_valid)) f[0] = f[1] = f[2] = f[3] = 0.f;
            if (DT == DT_TF32X3) {
                float hi[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) hi[i] = to_tf32(f[i]), lo[i] = f[i] - hi[i];   // exact difference; the MMA truncates lo to TF32
                *reinterpret_cast<float4 *>(p0 + j * RSTEP * 128) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4 *>(box_lo + r0 * 128 + off + j * RSTEP * 128) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            } else {
                *reinterpret_cast<float4 *>(p0 + j * RSTEP * 128) = make_float4(f[0], f[1], f[2], f[3]);
            }
        }
    }
}

// Pooled operand transform (bf16): the landed box holds the stored pre-activation z of the LAST shared-MLP layer and is
// rewritten in place as the BatchNorm/ReLU/max-pool backward of it,
//     dZ[m, ch] = pgo[g, ch] * [m - g*Kg == arg[g, ch]]  -  w[ch] * z[m, ch]  +  e[ch]          (g = m / Kg),
// i.e. what mpb_bn_bwd_apply (pooled form) would have written to HBM and the two consumer GEMMs read back: one read and one
// write of the widest activation tensor of the layer stack are gone (reference: the autograd of pointnet2_utils.py:210-214).
// The dense part is the plain column-owner affine transform (transform_cols without the ReLU); the sparse part -- ONE element
// per (group, channel) -- is handled around it: pool_gather computes the FINAL value of every special element of the box from
// the original z before the dense pass overwrites it, pool_scatter stores those values after it (two 128-thread named
// barriers in between).  A first version compared every element's row against arg[g, ch] inside the dense pass: 16 extra
// instructions per 16 bytes made the GEMMs transform-bound (dgrad 80 -> 139 us at M = 1M).
// Box geometry: `rows` rows of 128 bytes (64 channels ch_base .. ch_base+63), first row = global row R0; Kg divides `rows`
// or is a multiple of it (group boundaries never fall inside... a box holds whole groups or part of one).
// pool_gather: the final values of this thread's special elements of one box, from the original z in the landed box and the
// (arg-max, pgo) slices of the box's pooling groups that the TMA producer loaded into the same pipeline stage (tab_arg / tab_pgo:
// [groups][W] words, W = 64 * boxes covered; the first version fetched them with plain global loads from the transform warps and
// their latency under a saturated HBM -- several microseconds -- stalled the whole pipeline: 79 -> 126 us).
// rel0 = global row of (first group, row 0) minus the box's first row; g_valid = groups of the box that exist (rows < M).
template <int NP>
__device__ __forceinline__ void pool_gather(const uint8_t *box, int rows, int Kg, int ngroups, int g_valid, int rel0, const int32_t *tab_arg,
                                            const float *tab_pgo, int W, int col0, const float *negw, const float *e, int t, uint32_t (&off)[NP],
                                            uint32_t (&val)[NP])
{
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        off[i] = 0xffffffffu;
        const int q = t + 128 * i;
        const int gl = q >> 6, ch = q & 63;
        if (gl < ngroups && gl < g_valid) {
            const int r = rel0 + gl * Kg + tab_arg[gl * W + col0 + ch];      // the arg-max row of (group, ch), relative to the box
            if (r >= 0 && r < rows) {
                const uint32_t o = (uint32_t)(r * 128 + ((((ch >> 3) ^ (r & 7)) << 4) | ((ch & 7) << 1)));
                const float z = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(box + o));
                off[i] = o;
                val[i] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(fmaf(z, negw[ch], e[ch]) + tab_pgo[gl * W + col0 + ch]));
            }
        }
    }
}
template <int NP>
__device__ __forceinline__ void pool_scatter(uint8_t *box, const uint32_t (&off)[NP], const uint32_t (&val)[NP])
{
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (off[i] != 0xffffffffu) *reinterpret_cast<uint16_t *>(box + off[i]) = (uint16_t)val[i];
}

struct GemmTnArgs {
    int M, N, K, BN, stages, out_bufs;
    int pair;                          // N == 2*BN computed as ONE tile: both halves share the landed (and transformed) A stage, accumulator h and
                                       // epilogue warpgroup h take columns h*BN .. (bf16 + tensor-core statistics, whose TMEM layout caps BN at 128)
    const float *a_scale, *a_shift;    // XFORM 1: A' = relu(a_scale[k] * A + a_shift[k]); XFORM 2 (pooled): (-w[k], e[k])
    const float *z_scale, *z_shift;    // EPI 2: relu mask of the layer below
    const int32_t *pool_arg;           // XFORM 2: arg-max row of every (group, channel) [M / pool_k, K]
    const float *pool_pgo;             //          p * dY at that row [M / pool_k, K]
    int pool_k;                        //          rows per pooling group
    float *partials;                   // EPI 1/2: [nparts][2][N]; nparts = 8 * grid (CUDA-core statistics) or 2 * grid (tensor-core)
};

constexpr int kGemmTnThreads = 384;    // producer, MMA, allocator, spare + two epilogue warpgroups
constexpr int kGemmTnThreadsXf = 640;  // + two transform warpgroups (each thread rewrites half a row of the landed A tile)

// Statistics on the tensor core (bf16 kernels, TCS): the 128-row x 64-column tile the epilogue has just staged in shared
// memory (swizzled, exactly a wgrad-style MN-major operand) is multiplied with itself,
//     D[128 x 64] (+)= [ X | E ]^T [128 rows -> M = 64 + 64]  *  Y [128 rows x 64]      (8 MMAs of K = 16 rows)
// where E is a constant box whose column 0 is all ones.  Rows 0..63 of D accumulate the Gram block X^T Y -- its DIAGONAL is
// the per-column sum of x*y -- and row 64 accumulates 1^T Y, the per-column sum of y.  bf16 x bf16 products are exact in the
// fp32 accumulator, so these are the same sums the CUDA-core butterfly produced, for 8 issued instructions per block instead
// of ~250 shuffle/select/add instructions per warp (ncu: the statistics epilogue made the kernel issue-bound, 30 M warp
// instructions against 7 M for the plain store).  EPI 1: X = Y = stored output -> (sum z, sum z^2).  EPI 2: X = z tile of the
// layer below (TMA-loaded), Y = dY = C * [z_scale*z + z_shift > 0] -> (sum dY, sum dY*z).  The accumulators live in TMEM for
// the whole kernel (one set per epilogue warpgroup) and are read once at the end.
// EPI: 0 = store only; 1 = + forward BatchNorm statistics of the stored values; 2 = + BatchNorm-backward statistics.
// XFORM: 0 = A as stored; 1 = previous layer's BatchNorm + ReLU on A; 2 = pooled BatchNorm/ReLU/max backward on A (bf16).
template <int DT, int XFORM, int EPI>
__global__ void __launch_bounds__((XFORM || DT == DT_TF32X3) ? kGemmTnThreadsXf : kGemmTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmArg,
               const __grid_constant__ CUtensorMap tmPgo, const GemmTnArgs p)
{
    constexpr bool kXf = XFORM != 0 || DT == DT_TF32X3;
    constexpr bool TCS = DT == DT_BF16 && EPI != 0;       // statistics on the tensor core
    static_assert(XFORM != 2 || DT == DT_BF16, "the pooled operand transform is built for bf16 activations");
    constexpr int EPR = DT == DT_BF16 ? 64 : 32;          // elements per 128-byte row = columns per k-block and per output block
    constexpr int NA = DT == DT_TF32X3 ? 2 : 1;           // operand copies per stage (hi, lo)
    constexpr uint32_t kAccStride = TCS ? 128u : 256u;    // TMEM columns between the two accumulators (TCS: BN <= 128)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B tiles need 1024-B alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = p.M, N = p.N, K = p.K, BN = p.BN, stages = p.stages;
    const uint32_t b_bytes = (uint32_t)BN * 128u;
    const int halves = p.pair ? 2 : 1;                                               // B tiles (of BN rows) per stage
    const uint32_t pool_off = NA * (kABytes + halves * b_bytes);                     // XFORM 2: [arg slices | pgo slices] behind the operands
    const uint32_t stage_bytes = pool_off + (XFORM == 2 ? kPoolStageBytes : 0);      // [A | A_lo | B | B_lo | pool tables]
    const int pool_groups = XFORM == 2 ? (p.pool_k >= kTileM ? 1 : kTileM / p.pool_k) : 0;   // pooling groups per 128-row tile
    // carve-up (every tile 1024-byte aligned): ring | output staging [2 wg][out_bufs] | z tiles [2 wg][2] (EPI 2) |
    // dY tiles [2 wg] (TCS, EPI 2) | ones box (TCS) | CUDA-core statistic slots [8 warps][2][256] (!TCS) | vectors | barriers
    uint8_t *staging = smem + (size_t)stages * stage_bytes;
    uint8_t *zstage = staging + (size_t)2 * p.out_bufs * kABytes;
    uint8_t *dystage = zstage + (EPI == 2 ? 4 * kABytes : 0);
    uint8_t *ones = dystage + ((TCS && EPI == 2) ? 2 * kABytes : 0);
    float *stat_acc = reinterpret_cast<float *>(ones + (TCS ? kABytes : 0));
    float *vec = stat_acc + ((EPI != 0 && !TCS) ? 8 * 2 * 256 : 0);                           // [a_scale K][a_shift K][z_scale N][z_shift N]
    float *s_ascale = vec, *s_ashift = vec + (XFORM ? K : 0);
    float *s_zscale = s_ashift + (XFORM ? K : 0), *s_zshift = s_zscale + (EPI == 2 ? N : 0);
    GemmSmemTail *tail = reinterpret_cast<GemmSmemTail *>(s_zshift + (EPI == 2 ? N : 0));
    const int num_kb = K / EPR;
    const int tiles_m = (M + kTileM - 1) / kTileM, tiles_n = p.pair ? 1 : N / BN;
    const int total = tiles_m * tiles_n;

    if (XFORM)
        for (int i = threadIdx.x; i < K; i += blockDim.x) s_ascale[i] = p.a_scale[i], s_ashift[i] = p.a_shift[i];
    if (EPI == 2)
        for (int i = threadIdx.x; i < N; i += blockDim.x) s_zscale[i] = p.z_scale[i], s_zshift[i] = p.z_shift[i];
    if (TCS) {   // E: 128 rows x 64 bf16 columns, column 0 = 1 (chunk 0 of row r sits at physical chunk r & 7), everything else 0
        for (int i = threadIdx.x; i < kABytes / 16; i += blockDim.x) {
            const int r = i >> 3, pc = i & 7;
            reinterpret_cast<uint4 *>(ones)[i] = make_uint4(pc == (r & 7) ? 0x00003F80u : 0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
    }
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmA);
        prefetch_tensormap(&tmB);
        prefetch_tensormap(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&tail->full[s], 1);
            mbar_init(&tail->empty[s], 1);
            mbar_init(&tail->ready[s], 4);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tail->tfull[a], 1);
            mbar_init(&tail->tempty[a], 4);
            mbar_init(&tail->zfull[a][0], 1);
            mbar_init(&tail->zfull[a][1], 1);
            mbar_init(&tail->sdone[a][0], 1);
            mbar_init(&tail->sdone[a][1], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(&tail->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * kTileM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&tail->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&tail->full[stage], kABytes + NA * halves * b_bytes + (XFORM == 2 ? (uint32_t)pool_groups * 512u : 0u));
                    uint8_t *sa = smem + (size_t)stage * stage_bytes;
                    if (XFORM == 2) {   // the pooling tables of this tile's groups, channels kb*64 .. +63 (rows past the table: zero fill)
                        tma_load_2d(sa + pool_off, &tmArg, kb * EPR, m0 / p.pool_k, &tail->full[stage]);
                        tma_load_2d(sa + pool_off + pool_groups * 256, &tmPgo, kb * EPR, m0 / p.pool_k, &tail->full[stage]);
                    }
                    tma_load_2d(sa, &tmA, kb * EPR, m0, &tail->full[stage]);
                    tma_load_2d(sa + NA * kABytes, &tmB, kb * EPR, n0, &tail->full[stage]);
                    if (halves == 2) tma_load_2d(sa + NA * kABytes + b_bytes, &tmB, kb * EPR, BN, &tail->full[stage]);   // pair: bf16 only (NA = 1)
                    if (DT == DT_TF32X3) tma_load_2d(sa + NA * kABytes + b_bytes, &tmBlo, kb * EPR, n0, &tail->full[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = DT == DT_BF16 ? make_idesc_bf16((uint32_t)BN, false, false) : make_idesc_tf32((uint32_t)BN, false, false);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                mbar_wait(&tail->tempty[acc], acc_phase ^ 1);
                if (halves == 2) mbar_wait(&tail->tempty[1], acc_phase ^ 1);       // pair: both accumulators belong to this tile
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccStride;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&tail->full[stage], phase);              // B (and the raw A) have landed
                    if (kXf) mbar_wait(&tail->ready[stage], phase);    // A has been rewritten by the transform warps
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(sa + NA * kABytes, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {   // one MMA consumes 32 bytes of every 128-byte row (16 bf16 / 8 tf32)
                        const uint64_t ko = (uint64_t)(k * 2);
                        if (DT == DT_BF16) {
                            umma_bf16(d_tmem, adesc + ko, bdesc + ko, idesc, (kb | k) ? 1u : 0u);
                        } else if (DT == DT_TF32) {
                            umma_tf32(d_tmem, adesc + ko, bdesc + ko, idesc, (kb | k) ? 1u : 0u);
                        } else {
                            const uint64_t alo = make_smem_desc_sw128(sa + kABytes, 16, 1024);
                            const uint64_t blo = make_smem_desc_sw128(sa + NA * kABytes + b_bytes, 16, 1024);
                            umma_tf32(d_tmem, alo + ko, bdesc + ko, idesc, (kb | k) ? 1u : 0u);   // small terms first
                            umma_tf32(d_tmem, adesc + ko, blo + ko, idesc, 1u);
                            umma_tf32(d_tmem, adesc + ko, bdesc + ko, idesc, 1u);
                        }
                    }
                    if (DT == DT_BF16 && halves == 2) {   // second half of the columns: same A stage, second B tile, accumulator 1
                        const uint64_t bdesc1 = make_smem_desc_sw128(sa + kABytes + b_bytes, 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem + kAccStride, adesc + (uint64_t)(k * 2), bdesc1 + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&tail->empty[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
                umma_commit(&tail->tfull[acc]);
                if (halves == 2) {
                    umma_commit(&tail->tfull[1]);
                    acc_phase ^= 1;                                                 // acc stays 0: every tile uses both buffers
                } else if ((acc ^= 1) == 0) {
                    acc_phase ^= 1;
                }
            }
        }
    } else if (kXf && warp >= 12) {
        // two transform warpgroups work on ALTERNATE k-blocks (two stages in flight: the wake-up, shared-memory and proxy-fence
        // latencies of one overlap the other's); inside a warpgroup thread t owns chunk t & 7 of the rows (t >> 3) + 16 j
        const int wg = (warp - 12) >> 2;
        const int t = (threadIdx.x - 384) & 127;
        const int c = t & 7, r0 = t >> 3;
        if (XFORM == 2) {
            // pooled transform: special elements first (from the original z), the dense affine pass, then the special values back
            int stage = 0, it = 0;
            uint32_t phase = 0;
            const int Kg = p.pool_k;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * kTileM;
                const int g0 = m0 / Kg;
                const int rel0 = g0 * Kg - m0;                                   // 0 unless a group spans several tiles
                const int g_valid = (M - g0 * Kg + Kg - 1) / Kg;                 // groups of this tile that exist
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    mbar_wait(&tail->full[stage], phase);              // BOTH warpgroups observe every phase (see below)
                    if ((it & 1) == wg) {
                        uint8_t *sa = smem + (size_t)stage * stage_bytes;
                        uint32_t soff[4], sval[4];
                        pool_gather<4>(sa, kTileM, Kg, pool_groups, g_valid, rel0, reinterpret_cast<const int32_t *>(sa + pool_off),
                                       reinterpret_cast<const float *>(sa + pool_off + pool_groups * 256), EPR, 0, s_ascale + kb * EPR,
                                       s_ashift + kb * EPR, t, soff, sval);
                        named_bar_sync(3 + wg, 128);       // every special element has been read from the original z
                        transform_cols<DT, true, false, 8, 16, false>(sa, sa, c, r0, s_ascale + kb * EPR, s_ashift + kb * EPR, M - m0);
                        named_bar_sync(3 + wg, 128);
                        pool_scatter<4>(sa, soff, sval);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tail->ready[stage]);
                    }
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
            }
        } else {
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * kTileM;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    // BOTH warpgroups wait for every stage, also the ones they do not transform.  mbarrier waits test a phase PARITY: a
                    // waiter that has not observed phase n of a barrier and asks for phase n+1 gets "done" while phase n is still
                    // pending.  With an odd stage count a stage changes hands between the warpgroups every revolution, and when the
                    // TMA loads of two consecutive stages completed out of order (seen under NCCL traffic with the 3-stage 3xTF32 ring)
                    // the idle warpgroup ran one revolution ahead, rewrote a stage that had not landed and arrived on `ready` early:
                    // a deadlock ~once per thousand steps (found with the MPB_MBAR_DEBUG build).
                    mbar_wait(&tail->full[stage], phase);
                    if ((it & 1) == wg) {
                        uint8_t *sa = smem + (size_t)stage * stage_bytes;
                        transform_cols<DT, XFORM == 1, false, 8, 16>(sa, sa + kABytes, c, r0, s_ascale + kb * EPR, s_ashift + kb * EPR, M - m0);
                        fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async-proxy reads
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tail->ready[stage]);
                    }
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // Two epilogue warpgroups: warps 4-7 drain accumulator 0 (even tiles of this CTA), warps 8-11 accumulator 1
        // (odd tiles), so TMEM->register->global of one tile overlaps the next tile's drain as well as its MMAs.
        const int ew = warp & 3;           // the TMEM lane quarter this warp may read (warp % 4)
        const int acc = (warp - 4) >> 2;   // accumulator buffer owned by this warpgroup
        uint32_t acc_phase = 0;
        int nblk = 0;                                   // 128-byte column blocks processed so far by this warpgroup
        const bool issuer = ew == 0 && lane == 0;       // the one thread of the warpgroup that owns its bulk-store groups / z loads / statistic MMAs
        const int r_in_tile = ew * 32 + lane;
        const int r7 = r_in_tile & 7;
        float *my_stat = stat_acc + (size_t)(warp - 4) * 2 * 256;
        if (EPI != 0 && !TCS)
            for (int c = lane; c < 2 * BN; c += 32) my_stat[c] = 0.f;
        uint8_t *my_staging = staging + (size_t)acc * p.out_bufs * kABytes;
        uint8_t *my_z = zstage + (size_t)acc * 2 * kABytes;
        uint8_t *my_dy = dystage + (size_t)acc * kABytes;
        const uint32_t stat_tmem = tmem_base + 256u + (uint32_t)acc * 128u;     // TCS: [128 lanes x 64 columns] per column block
        const uint32_t sidesc = make_idesc_bf16(64u, true, true);
        uint32_t stat_used = 0;                                                  // bit j: column block j's accumulator has been started
        // z-tile prefetch (EPI 2): CUDA-core statistics keep two blocks in flight; the tensor-core variant refills a z buffer
        // only once the statistic MMA that read it has completed, i.e. one block ahead
        int z_tile = blockIdx.x + acc * (int)gridDim.x, z_c0 = 0, z_issued = 0;
        auto issue_z = [&]() {
            if (z_tile >= total) return;
            const int zm0 = (z_tile / tiles_n) * kTileM, zn0 = (z_tile % tiles_n) * BN;
            uint64_t *bar = &tail->zfull[acc][z_issued & 1];
            mbar_arrive_expect_tx(bar, kABytes);
            tma_load_2d(my_z + (size_t)(z_issued & 1) * kABytes, &tmZ, zn0 + z_c0, zm0, bar);
            ++z_issued;
            z_c0 += EPR;
            if (z_c0 >= BN) z_c0 = 0, z_tile += 2 * (int)gridDim.x;
        };
        if (EPI == 2 && issuer) {
            issue_z();
            if (!TCS) issue_z();
        }
        int t = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++t) {
            if (halves == 1 && (t & 1) != acc) continue;            // pair: every tile, this warpgroup's half of the columns
            const int m0 = (tile / tiles_n) * kTileM, n0 = halves == 2 ? acc * BN : (tile % tiles_n) * BN;
            mbar_wait(&tail->tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)acc * kAccStride + ((uint32_t)(ew * 32) << 16);
            for (int c0 = 0; c0 < BN; c0 += EPR, ++nblk) {
                const int ob = p.out_bufs == 2 ? (nblk & 1) : 0;
                uint8_t *buf = my_staging + (size_t)ob * kABytes;
                if (issuer) {
                    // the TMA store that last used this staging buffer has drained its reads ...
                    if (p.out_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
                    if (TCS) {
                        // ... and so has the statistic MMA that read it (EPI 1) / that read the dY tile and the z buffer about to
                        // be refilled (EPI 2: one dY tile, so the previous block's MMA must be complete)
                        if (EPI == 1) {
                            const int uses = p.out_bufs == 2 ? (nblk >> 1) : nblk;       // earlier uses of this buffer
                            if (uses > 0) mbar_wait(&tail->sdone[acc][ob], (uint32_t)(uses - 1) & 1u);
                        } else {
                            if (nblk > 0) mbar_wait(&tail->sdone[acc][0], (uint32_t)(nblk - 1) & 1u);
                            issue_z();                                                     // z tile of block nblk + 1
                        }
                    }
                }
                const uint8_t *zrow = my_z + (size_t)(nblk & 1) * kABytes + r_in_tile * 128;
                if (EPI == 2) mbar_wait(&tail->zfull[acc][nblk & 1], (uint32_t)(nblk >> 1) & 1u);
                named_bar_sync(1 + acc, 128);
                uint8_t *rowp = buf + r_in_tile * 128;
                uint8_t *dyrow = my_dy + r_in_tile * 128;
#pragma unroll
                for (int h = 0; h < EPR / 32; ++h) {                 // 32 accumulator columns per TMEM load
                    if (c0 + 32 * h < BN) {
                        uint32_t r[32];
                        tmem_ld_32x32(taddr + (uint32_t)(c0 + 32 * h), r);
                        float zv[32], zq[32];
                        if (DT == DT_BF16) {
                            uint32_t pk[16];
#pragma unroll
                            for (int v = 0; v < 16; ++v) pk[v] = pack_bf16(__uint_as_float(r[2 * v]), __uint_as_float(r[2 * v + 1]));
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const int chunk = (4 * h + v) ^ r7;   // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
                                *reinterpret_cast<uint4 *>(rowp + chunk * 16) = make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
                            }
                            if (TCS && EPI == 2) {
                                // dY tile for the statistic MMA: the stored (bf16) gradient where the layer below was active, else 0
                                const float *zs = s_zscale + n0 + c0 + 32 * h, *zh = s_zshift + n0 + c0 + 32 * h;
#pragma unroll
                                for (int v = 0; v < 4; ++v) {
                                    const int chunk = (4 * h + v) ^ r7;
                                    const uint4 zr = *reinterpret_cast<const uint4 *>(zrow + (chunk << 4));
                                    const uint32_t zw[4] = {zr.x, zr.y, zr.z, zr.w};
                                    const float4 s0 = *reinterpret_cast<const float4 *>(zs + 8 * v), s1 = *reinterpret_cast<const float4 *>(zs + 8 * v + 4);
                                    const float4 h0 = *reinterpret_cast<const float4 *>(zh + 8 * v), h1 = *reinterpret_cast<const float4 *>(zh + 8 * v + 4);
                                    const float sc8[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                                    const float sh8[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                                    uint32_t dw[4];
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const float2 z2 = unpack_bf16(zw[u]);
                                        const uint32_t c2 = pk[4 * v + u];
                                        const uint32_t lo = fmaf(z2.x, sc8[2 * u], sh8[2 * u]) > 0.f ? (c2 & 0x0000FFFFu) : 0u;
                                        const uint32_t hi = fmaf(z2.y, sc8[2 * u + 1], sh8[2 * u + 1]) > 0.f ? (c2 & 0xFFFF0000u) : 0u;
                                        dw[u] = lo | hi;
                                    }
                                    *reinterpret_cast<uint4 *>(dyrow + (chunk << 4)) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                                }
                            }
                            if (EPI != 0 && !TCS) {   // statistics of exactly what was stored
#pragma unroll
                                for (int v = 0; v < 16; ++v) {
                                    const float2 f = unpack_bf16(pk[v]);
                                    zv[2 * v] = f.x, zv[2 * v + 1] = f.y;
                                }
                            }
                        } else {
#pragma unroll
                            for (int v = 0; v < 8; ++v) {
                                const int chunk = v ^ r7;
                                *reinterpret_cast<uint4 *>(rowp + chunk * 16) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                            }
                            if (EPI != 0) {
#pragma unroll
                                for (int v = 0; v < 32; ++v) zv[v] = __uint_as_float(r[v]);
                            }
                        }
                        if (EPI == 1 && !TCS) {
#pragma unroll
                            for (int v = 0; v < 32; ++v) zq[v] = zv[v] * zv[v];
                        }
                        if (EPI == 2 && !TCS) {   // dY = C * [relu mask of the layer below]; second moment against the raw pre-activation z
                            const float *zs = s_zscale + n0 + c0 + 32 * h, *zh = s_zshift + n0 + c0 + 32 * h;
#pragma unroll
                            for (int v = 0; v < 8; ++v) {
                                const float4 z4 = *reinterpret_cast<const float4 *>(zrow + ((v ^ r7) << 4));
                                const float4 s4 = *reinterpret_cast<const float4 *>(zs + 4 * v), h4 = *reinterpret_cast<const float4 *>(zh + 4 * v);
                                const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int j = 4 * v + u;
                                    const float d0 = fmaf(zz[u], ss[u], hh[u]) > 0.f ? zv[j] : 0.f;
                                    zv[j] = d0;
                                    zq[j] = d0 * zz[u];
                                }
                            }
                        }
                        if (EPI != 0 && !TCS) {
                            const float cs = warp_column_sum_32x32(zv, lane), cq = warp_column_sum_32x32(zq, lane);
                            my_stat[c0 + 32 * h + lane] += cs;
                            my_stat[BN + c0 + 32 * h + lane] += cq;
                        }
                    }
                }
                if (c0 + EPR >= BN) {                              // accumulator fully drained: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tail->tempty[acc]);
                }
                fence_proxy_async_smem();
                named_bar_sync(1 + acc, 128);                      // staging (and dY) tile complete, z tile fully read
                if (issuer) {
                    tma_store_2d(&tmC, buf, n0 + c0, m0);
                    bulk_commit();
                    if (EPI == 2 && !TCS) issue_z();               // refill the z buffer this block just released
                    if (TCS) {
                        // D[acc, block j] (+)= [X | E]^T * Y over the tile's 128 rows
                        const int j = c0 / EPR;
                        const uint32_t x_addr = smem_u32(EPI == 1 ? buf : my_z + (size_t)(nblk & 1) * kABytes);
                        const uint32_t y_addr = smem_u32(EPI == 1 ? buf : my_dy);
                        const uint64_t adesc = make_smem_desc_sw128(x_addr, smem_u32(ones) - x_addr, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(y_addr, kABytes, 1024);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < 8; ++k)   // 16 rows per MMA = two 8-row groups = 2048 B
                            umma_bf16(stat_tmem + (uint32_t)j * 64u, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), sidesc,
                                      (k > 0 || ((stat_used >> j) & 1u)) ? 1u : 0u);
                        stat_used |= 1u << j;
                        umma_commit(&tail->sdone[acc][EPI == 1 ? ob : 0]);
                    }
                }
            }
            acc_phase ^= 1;
        }
        if (issuer) bulk_wait<0>();          // all stores complete before the CTA (and its shared memory) goes away
        if (EPI != 0 && !TCS) {
            // one partial row per epilogue warp; a CTA always owns the same column tile (gridDim.x % tiles_n == 0)
            const int n0 = ((int)blockIdx.x % tiles_n) * BN;
            float *dst = p.partials + ((size_t)blockIdx.x * 8 + (warp - 4)) * 2 * N;
            for (int c = lane; c < N; c += 32) {
                const bool mine = c >= n0 && c < n0 + BN;
                dst[c] = mine ? my_stat[c - n0] : 0.f;
                dst[N + c] = mine ? my_stat[BN + c - n0] : 0.f;
            }
        }
        if (TCS) {
            // one partial row per warpgroup: wait for its last statistic MMA, then read row 64 (sums) and the Gram diagonal
            const int n0 = halves == 2 ? acc * BN : ((int)blockIdx.x % tiles_n) * BN;
            float *dst = p.partials + ((size_t)blockIdx.x * 2 + acc) * 2 * N;
            for (int c = r_in_tile; c < 2 * N; c += 128) dst[c] = 0.f;       // columns of other tiles / no tile processed at all
            named_bar_sync(1 + acc, 128);
            if (nblk > 0) {
                if (EPI == 1) {
                    const int last = nblk - 1, ob = p.out_bufs == 2 ? (last & 1) : 0;
                    const int uses = p.out_bufs == 2 ? (last >> 1) : last;
                    mbar_wait(&tail->sdone[acc][ob], (uint32_t)uses & 1u);
                    if (p.out_bufs == 2 && nblk > 1) {
                        const int prev = nblk - 2;
                        mbar_wait(&tail->sdone[acc][prev & 1], (uint32_t)(prev >> 1) & 1u);
                    }
                } else {
                    mbar_wait(&tail->sdone[acc][0], (uint32_t)(nblk - 1) & 1u);
                }
                tc_fence_after();
                const int nb = BN / EPR > 0 ? (BN + EPR - 1) / EPR : 1;
                for (int j = 0; j < nb; ++j) {
                    const uint32_t base = stat_tmem + (uint32_t)j * 64u + ((uint32_t)(ew * 32) << 16);
                    if (ew < 2) {            // rows 32*ew + lane of the Gram block: the diagonal element sits in column 32*ew + lane
                        uint32_t r[32];
                        tmem_ld_32x32(base + (uint32_t)(32 * ew), r);
                        float d = 0.f;
#pragma unroll
                        for (int k = 0; k < 32; ++k) d = lane == k ? __uint_as_float(r[k]) : d;
                        const int c = j * EPR + 32 * ew + lane;
                        if (c < BN) dst[N + n0 + c] = d;
                    } else if (ew == 2) {    // row 64 (lane 0 of this warp): the 64 column sums
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            uint32_t r[32];
                            tmem_ld_32x32(base + (uint32_t)(32 * hh), r);
                            if (lane == 0) {
#pragma unroll
                                for (int k = 0; k < 32; ++k) {
                                    const int c = j * EPR + 32 * hh + k;
                                    if (c < BN) dst[n0 + c] = __uint_as_float(r[k]);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// ---- weight gradient ------------------------------------------------------------------------------------------------
// dW_partial[ms][n0 + i, k0 + j] = sum_{m in split ms} dZ[m, n0 + i] * f(A)[m, k0 + j]
// Both operands are read exactly as they sit in HBM (row-major [M, C]) and fed to the tensor core as MN-major
// operands (contraction index = row), so no transposed copy of dZ or A is ever made.
// blockIdx.x = ((n_tile * k_tiles) + k_tile) * m_splits + m_split
struct WgradArgs {
    int M, N, K, k_tiles, m_splits, rows_per_split, stages;
    const float *a_scale, *a_shift;   // XFORM on A
    float *partials;                  // [m_splits][N_pad128][K] fp32
    int n_pad;
    const int32_t *pool_arg;          // ZPOOL: arg-max row per (group, channel) [M / pool_k, N]
    const float *pool_pgo;            //        p * dY at that row [M / pool_k, N]
    const float *pool_negw, *pool_e;  //        [N] each
    int pool_k;
    int ab;                           // A boxes a stage holds: min(K, KT) / EPR (the stage is sized for the real K, not for KT)
};

constexpr int kWgradThreads = 256;
constexpr int kWgradThreadsXf = 384;

// ZPOOL (bf16): the dZ operand is rebuilt from the stored pre-activation of the max-pooled layer (transform_pool).
template <int DT, bool XFORM, bool ZPOOL = false>
__global__ void __launch_bounds__((XFORM || ZPOOL || DT == DT_TF32X3) ? kWgradThreadsXf : kWgradThreads, DT == DT_BF16 ? 2 : 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmArg,
             const __grid_constant__ CUtensorMap tmPgo, const WgradArgs p)
{
    constexpr bool kXf = XFORM || ZPOOL || DT == DT_TF32X3;
    static_assert(!ZPOOL || DT == DT_BF16, "the pooled operand transform is built for bf16 activations");
    constexpr int EPR = DT == DT_BF16 ? 64 : 32;          // channels per 128-byte box row
    constexpr int RB = DT == DT_BF16 ? 64 : 32;           // contraction rows per stage
    constexpr int kBox = RB * 128;                        // one box: RB rows x 128 bytes
    constexpr int KT = DT == DT_BF16 ? 256 : 128;         // A channels (dW columns) per CTA
    constexpr int ZB = 128 / EPR;                         // dZ boxes per stage (128 channels)
    const int AB = p.ab;                                  // A boxes per stage
    constexpr int NA = DT == DT_TF32X3 ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = p.M, N = p.N, K = p.K, stages = p.stages;
    const int ms = blockIdx.x % p.m_splits;
    const int kt = (blockIdx.x / p.m_splits) % p.k_tiles;
    const int nt = blockIdx.x / (p.m_splits * p.k_tiles);
    const int n0 = nt * 128, k0 = kt * KT;
    const int NU = min(KT, K - k0);                       // multiple of EPR
    const int a_boxes = min(ZB, (N - n0 + EPR - 1) / EPR);   // boxes of dZ that hold real data
    const int b_boxes = NU / EPR;
    // stage layout: [dZ boxes (ZB)] [A boxes (AB)] and, for the split, the same again for the low parts
    const uint32_t half_bytes = (uint32_t)(ZB + AB) * kBox;
    const uint32_t pool_off = NA * half_bytes;                                       // ZPOOL: [arg slices | pgo slices], 128 channels wide
    const uint32_t stage_bytes = pool_off + (ZPOOL ? kPoolStageBytes : 0);
    const int pool_groups = ZPOOL ? (p.pool_k >= RB ? 1 : RB / p.pool_k) : 0;        // pooling groups per row block
    float *s_ascale = reinterpret_cast<float *>(smem + (size_t)stages * stage_bytes);
    float *s_ashift = s_ascale + (XFORM ? KT : 0);
    float *s_pw = s_ashift + (XFORM ? KT : 0), *s_pe = s_pw + (ZPOOL ? 128 : 0);
    GemmSmemTail *tail = reinterpret_cast<GemmSmemTail *>(s_pe + (ZPOOL ? 128 : 0));
    const int m_begin = ms * p.rows_per_split, m_end = min(M, m_begin + p.rows_per_split);
    const int num_rb = m_end > m_begin ? (m_end - m_begin + RB - 1) / RB : 0;

    if (XFORM)
        for (int i = threadIdx.x; i < NU; i += blockDim.x) s_ascale[i] = p.a_scale[k0 + i], s_ashift[i] = p.a_shift[k0 + i];
    if (ZPOOL)
        for (int i = threadIdx.x; i < 128; i += blockDim.x) {
            const bool in = n0 + i < N;
            s_pw[i] = in ? p.pool_negw[n0 + i] : 0.f, s_pe[i] = in ? p.pool_e[n0 + i] : 0.f;
        }
    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmZ);
        prefetch_tensormap(&tmA);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&tail->full[s], 1);
            mbar_init(&tail->empty[s], 1);
            mbar_init(&tail->ready[s], 4);
        }
        mbar_init(&tail->tfull[0], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<256>(&tail->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (num_rb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int rb = 0; rb < num_rb; ++rb) {
                    mbar_wait(&tail->empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&tail->full[stage], (uint32_t)(a_boxes + b_boxes) * kBox + (ZPOOL ? (uint32_t)pool_groups * 1024u : 0u));
                    uint8_t *s = smem + (size_t)stage * stage_bytes;
                    const int r0 = m_begin + rb * RB;   // rows_per_split is a multiple of RB: a box never straddles two splits
                    if (ZPOOL) {   // pooling tables of the row block's groups, channels n0 .. n0+127 (outside the table: zero fill)
                        tma_load_2d(s + pool_off, &tmArg, n0, r0 / p.pool_k, &tail->full[stage]);
                        tma_load_2d(s + pool_off + pool_groups * 512, &tmPgo, n0, r0 / p.pool_k, &tail->full[stage]);
                    }
                    for (int b = 0; b < a_boxes; ++b) tma_load_2d(s + b * kBox, &tmZ, n0 + b * EPR, r0, &tail->full[stage]);
                    for (int b = 0; b < b_boxes; ++b) tma_load_2d(s + (ZB + b) * kBox, &tmA, k0 + b * EPR, r0, &tail->full[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = DT == DT_BF16 ? make_idesc_bf16((uint32_t)NU, true, true) : make_idesc_tf32((uint32_t)NU, true, true);
                int stage = 0;
                uint32_t phase = 0;
                for (int rb = 0; rb < num_rb; ++rb) {
                    mbar_wait(&tail->full[stage], phase);
                    if (kXf) mbar_wait(&tail->ready[stage], phase);
                    tc_fence_after();
                    const uint32_t s = smem_u32(smem + (size_t)stage * stage_bytes);
                    // MN-major SWIZZLE_128B: LBO = distance between EPR-channel column blocks (one box),
                    // SBO = distance between 8-row groups along the contraction (1024 B)
                    // (32-bit operands: the 32-byte-atom swizzle, whose pattern repeats every 4 rows -> SBO = 512 B)
                    constexpr uint32_t kLt = DT == DT_BF16 ? 2u : 1u, kSbo = DT == DT_BF16 ? 1024u : 512u;
                    const uint64_t adesc = make_smem_desc(s, kBox, kSbo, kLt);
                    const uint64_t bdesc = make_smem_desc(s + ZB * kBox, kBox, kSbo, kLt);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // one MMA: 16 contraction rows (bf16) = 2048 B, 8 rows (tf32) = 1024 B
                        const uint64_t ko = (uint64_t)(k * (DT == DT_BF16 ? 128 : 64));
                        if (DT == DT_BF16) {
                            umma_bf16(tmem_base, adesc + ko, bdesc + ko, idesc, (rb | k) ? 1u : 0u);
                        } else if (DT == DT_TF32) {
                            umma_tf32(tmem_base, adesc + ko, bdesc + ko, idesc, (rb | k) ? 1u : 0u);
                        } else {
                            const uint64_t alo = make_smem_desc(s + half_bytes, kBox, kSbo, kLt);
                            const uint64_t blo = make_smem_desc(s + half_bytes + ZB * kBox, kBox, kSbo, kLt);
                            umma_tf32(tmem_base, alo + ko, bdesc + ko, idesc, (rb | k) ? 1u : 0u);
                            umma_tf32(tmem_base, adesc + ko, blo + ko, idesc, 1u);
                            umma_tf32(tmem_base, adesc + ko, bdesc + ko, idesc, 1u);
                        }
                    }
                    umma_commit(&tail->empty[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
                umma_commit(&tail->tfull[0]);
            }
        } else if (kXf && warp >= 8) {
            // transform warpgroup: thread t owns chunk t & 7 of the rows (t >> 3) + 16 j of every box that needs rewriting
            const int t = threadIdx.x - 256;
            const int c = t & 7, r0 = t >> 3;
            int stage = 0;
            uint32_t phase = 0;
            const int first_box = DT == DT_TF32X3 ? 0 : ZB;            // bf16 / single-pass tf32: only the A boxes change
            for (int rb = 0; rb < num_rb; ++rb) {
                uint8_t *s = smem + (size_t)stage * stage_bytes;
                // rows past M were zero-filled by the TMA; they must stay zero through relu(shift)
                const int rows_valid = M - (m_begin + rb * RB);
                mbar_wait(&tail->full[stage], phase);
                if (ZPOOL) {
                    // dZ boxes: special elements first (from the original z), the dense affine pass, then the special values back
                    uint32_t soff[2][2], sval[2][2];
                    const int R0 = m_begin + rb * RB, Kg = p.pool_k;
                    const int g0 = R0 / Kg, rel0 = g0 * Kg - R0, g_valid = (M - g0 * Kg + Kg - 1) / Kg;
                    const int32_t *tab_arg = reinterpret_cast<const int32_t *>(s + pool_off);
                    const float *tab_pgo = reinterpret_cast<const float *>(s + pool_off + pool_groups * 512);
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        pool_gather<2>(s + (size_t)b * kBox, RB, Kg, b < a_boxes ? pool_groups : 0, g_valid, rel0, tab_arg, tab_pgo, 128, b * EPR, s_pw + b * EPR,
                                       s_pe + b * EPR, t, soff[b], sval[b]);
                    named_bar_sync(1, 128);
#pragma unroll
                    for (int b = 0; b < 2; ++b)
                        if (b < a_boxes) {
                            uint8_t *bp = s + (size_t)b * kBox;
                            transform_cols<DT, true, false, RB / 16, 16, false>(bp, bp, c, r0, s_pw + b * EPR, s_pe + b * EPR, rows_valid);
                        }
                    named_bar_sync(1, 128);
#pragma unroll
                    for (int b = 0; b < 2; ++b) pool_scatter<2>(s + (size_t)b * kBox, soff[b], sval[b]);
                }
                for (int box = first_box; box < ZB + AB; ++box) {
                    if (box < ZB ? (box >= a_boxes) : (box - ZB >= b_boxes)) continue;
                    uint8_t *bp = s + (size_t)box * kBox;
                    if (XFORM && box >= ZB)
                        transform_cols<DT, true, DT != DT_BF16, RB / 16, 16>(bp, bp + half_bytes, c, r0, s_ascale + (box - ZB) * EPR,
                                                                             s_ashift + (box - ZB) * EPR, rows_valid);
                    else if (!ZPOOL || box >= ZB)
                        transform_cols<DT, false, DT != DT_BF16, RB / 16, 16>(bp, bp + half_bytes, c, r0, s_ascale, s_ashift, rows_valid);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tail->ready[stage]);
                if (++stage == stages) stage = 0, phase ^= 1;
            }
        }
    }
    if (warp >= 4 && warp < 8) {
        const int ew = warp - 4;
        const int row = n0 + ew * 32 + lane;
        float *dst_row = p.partials + ((size_t)ms * p.n_pad + row) * K + k0;
        const bool real_row = row < N;            // rows >= N: channel padding of the 128-row tile, never read by the reduction
        if (num_rb > 0) {
            mbar_wait(&tail->tfull[0], 0);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
            for (int c0 = 0; c0 < NU; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + (uint32_t)c0, r);      // warp-collective: every lane takes part, only real rows store
                if (real_row) {
#pragma unroll
                    for (int v = 0; v < 32; v += 4)
                        *reinterpret_cast<float4 *>(dst_row + c0 + v) = make_float4(__uint_as_float(r[v]), __uint_as_float(r[v + 1]),
                                                                                    __uint_as_float(r[v + 2]), __uint_as_float(r[v + 3]));
                }
            }
        } else if (real_row) {
            for (int c0 = 0; c0 < NU; c0 += 4) *reinterpret_cast<float4 *>(dst_row + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<256>(tmem_base);
}

// dW[r, c] = sum over splits (fixed order) of the partial tiles, cropped to the real [cout, cin] and with the packed
// column order undone (xyz_last: the three coordinate channels were moved behind the features by the weight packing).
// A block = G warps x 32 consecutive outputs: warp g sums the splits i = g, g + G, ... (coalesced 128-byte rows of the
// partial slabs, four loads in flight), the G per-warp sums meet in shared memory and are added in warp order --
// the same order on every run, so the result is bitwise reproducible.  (The first version walked all splits from
// one thread per output: 74 dependent rounds of L2 latency on 16 CTAs.)
constexpr int kReduceMaxWarps = 32;
__global__ void __launch_bounds__(32 * kReduceMaxWarps)
wgrad_reduce_kernel(const float *__restrict__ partials, int m_splits, int n_pad, int K, int cout, int cin, int xyz_last, int accumulate,
                    float *__restrict__ dW)
{
    __shared__ float red[kReduceMaxWarps][32];
    const int G = blockDim.x >> 5, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (e < cout * cin) {
        const int r = e / cin, c = e - r * cin;
        const int pc = (xyz_last && cin > 3) ? (c < 3 ? cin - 3 + c : c - 3) : c;
        const float *src = partials + (size_t)r * K + pc;
        const size_t stride = (size_t)n_pad * K;
        int i = g;
        for (; i + 3 * G < m_splits; i += 4 * G) {
            const float a0 = src[(size_t)i * stride], a1 = src[(size_t)(i + G) * stride];
            const float a2 = src[(size_t)(i + 2 * G) * stride], a3 = src[(size_t)(i + 3 * G) * stride];
            s0 += a0, s1 += a1, s2 += a2, s3 += a3;
        }
        for (; i < m_splits; i += G) s0 += src[(size_t)i * stride];
    }
    red[g][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (g == 0 && e < cout * cin) {
        float t = 0.f;
        for (int w = 0; w < G; ++w) t += red[w][lane];
        dW[e] = accumulate ? dW[e] + t : t;
    }
}

// Columns per tile: the whole N when it fits one UMMA (N <= cap, multiple of 32); otherwise the largest multiple of
// the 128-byte block width dividing N.
static int pick_bn(int N, int cap, int blk)
{
    if (N <= cap) return N % 32 == 0 ? N : 0;
    for (int bn = cap; bn >= blk; bn -= blk)
        if (N % bn == 0) return bn;
    return 0;
}

struct TnPlan {
    int BN, stages, out_bufs, grid, threads, parts_per_cta, pair;
    size_t smem;
};

static bool os_pair_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MPB_GEMM_PAIR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

static bool plan_gemm_tn(int dt, int xform, int epi, int M, int N, int K, TnPlan *pl)
{
    const int epr = dt == DT_BF16 ? 64 : 32;
    const int na = dt == DT_TF32X3 ? 2 : 1;
    if (M <= 0 || N <= 0 || K <= 0 || K % epr || N % 32 || (xform && K > kMaxVec) || (epi == 2 && N > kMaxVec)) return false;
    const bool tcs = dt == DT_BF16 && epi != 0;            // statistics on the tensor core: accumulators share TMEM, so BN <= 128
    const int bn = pick_bn(N, (dt == DT_TF32X3 || tcs) ? 128 : 256, epr);
    if (bn <= 0 || (bn != N && bn % epr)) return false;
    if (dt != DT_BF16 && bn % 32) return false;
    // N = 2 x 128 with the tensor-core statistics (forward, last layer of sa2): ONE tile per 128 rows -- the A stage is loaded and
    // transformed once for both halves, accumulator / epilogue warpgroup h own columns h*128 .. (two separate N tiles re-read and
    // re-transformed A and ran at 3.6 TB/s: 105 us for 402 MB)
    const int pair = (tcs && epi == 1 && N == 2 * bn && bn == 128 && os_pair_enabled()) ? 1 : 0;
    const int stage_bytes = na * (kABytes + (pair ? 2 : 1) * bn * 128) + (xform == 2 ? kPoolStageBytes : 0);
    const int fixed0 = (epi == 2 ? 4 * kABytes : 0) + ((tcs && epi == 2) ? 2 * kABytes : 0) + (tcs ? kABytes : 0) +
                       ((epi != 0 && !tcs) ? 8 * 2 * 256 * 4 : 0) + (xform ? 2 * K * 4 : 0) + (epi == 2 ? 2 * N * 4 : 0) +
                       (int)sizeof(GemmSmemTail) + 1024;
    int out_bufs = 2;
    int stages = (227 * 1024 - fixed0 - 2 * out_bufs * kABytes) / stage_bytes;
    if (stages < 3) {
        out_bufs = 1;
        stages = (227 * 1024 - fixed0 - 2 * out_bufs * kABytes) / stage_bytes;
    }
    if (stages < 2) return false;
    stages = stages > 8 ? 8 : stages;
    const int tiles_n = pair ? 1 : N / bn;
    const int tiles = ((M + kTileM - 1) / kTileM) * tiles_n;
    int grid = tiles < sm_count() ? tiles : sm_count();
    grid = grid / tiles_n * tiles_n;       // a CTA always owns the same column tile (statistics partial rows)
    if (grid < tiles_n) return false;
    pl->BN = bn, pl->stages = stages, pl->out_bufs = out_bufs, pl->grid = grid, pl->pair = pair;
    pl->parts_per_cta = tcs ? 2 : 8;
    pl->threads = (xform || dt == DT_TF32X3) ? kGemmTnThreadsXf : kGemmTnThreads;
    pl->smem = (size_t)stages * stage_bytes + 2 * out_bufs * kABytes + fixed0;
    return true;
}

template <int DT, int XFORM, int EPI>
static int launch_tn(const TnPlan &pl, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmBlo, const CUtensorMap &tmC,
                     const CUtensorMap &tmZ, const CUtensorMap &tmArg, const CUtensorMap &tmPgo, const GemmTnArgs &args, cudaStream_t st)
{
    auto kern = gemm_tn_kernel<DT, XFORM, EPI>;
    MPB_ENSURE_DYN_SMEM(kern, 227 * 1024);
    kern<<<pl.grid, pl.threads, pl.smem, st>>>(tmA, tmB, tmBlo, tmC, tmZ, tmArg, tmPgo, args);
    return check_launch("gemm_tn_kernel");
}

}  // namespace mpb

// Number of [2][N] partial rows a fused-statistics launch writes (0: this shape cannot fuse the statistics).
extern "C" int mpb_sa_gemm_stat_partials(int dtype, int M, int N, int K, int xform, int epi)
{
    mpb::TnPlan pl;
    if (epi == 0 || !mpb::plan_gemm_tn(dtype, xform, epi, M, N, K, &pl)) return 0;
    return pl.parts_per_cta * pl.grid;
}

namespace mpb {
static int gemm_tn_impl(int dtype, const void *A, const void *B, const void *B_lo, void *C, int M, int N, int K, int xform,
                        const float *a_scale, const float *a_shift, const int32_t *pool_arg, const float *pool_pgo, int pool_k, int epi,
                        float *partials, int nparts, const void *Z, const float *z_scale, const float *z_shift, void *stream)
{
    MPB_REQUIRE(dtype >= DT_BF16 && dtype <= DT_TF32X3, "dtype must be 0 (bf16), 1 (tf32) or 2 (tf32x3)");
    MPB_REQUIRE(M >= 0 && N > 0 && K > 0, "bad size");
    if (M == 0) return MPB_OK;
    MPB_REQUIRE(A && B && C, "null pointer");
    MPB_REQUIRE(dtype != DT_TF32X3 || B_lo, "tf32x3 needs the low part of B");
    MPB_REQUIRE(epi >= 0 && epi <= 2, "epi must be 0, 1 or 2");
    MPB_REQUIRE(epi == 0 || partials, "statistics need a partials buffer");
    MPB_REQUIRE(epi != 2 || (Z && z_scale && z_shift), "epi 2 needs Z, z_scale, z_shift");
    MPB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0, "operands must be 16-byte aligned");
    TnPlan pl;
    MPB_REQUIRE(plan_gemm_tn(dtype, xform, epi, M, N, K, &pl), "unsupported shape (K multiple of 64 bf16 / 32 tf32, N multiple of 32, tiles must fit)");
    MPB_REQUIRE(epi == 0 || nparts == pl.parts_per_cta * pl.grid, "nparts mismatch (ask mpb_sa_gemm_stat_partials)");
    const int esz = dtype == DT_BF16 ? 2 : 4;
    CUtensorMap tmA, tmB, tmBlo, tmC, tmZ;
    int rc = make_map(&tmA, esz, A, M, K, K, kTileM);
    if (rc) return rc;
    rc = make_map(&tmB, esz, B, N, K, K, pl.BN);
    if (rc) return rc;
    tmBlo = tmB;
    if (dtype == DT_TF32X3) {
        rc = make_map(&tmBlo, esz, B_lo, N, K, K, pl.BN);
        if (rc) return rc;
    }
    rc = make_map(&tmC, esz, C, M, N, N, kTileM);
    if (rc) return rc;
    tmZ = tmC;
    if (epi == 2) {
        rc = make_map(&tmZ, esz, Z, M, N, N, kTileM);
        if (rc) return rc;
    }
    CUtensorMap tmArg = tmC, tmPgo = tmC;
    if (xform == 2) {
        const int groups = pool_k >= kTileM ? 1 : kTileM / pool_k;
        rc = make_map_plain32(&tmArg, pool_arg, M / pool_k, K, 64, groups);
        if (rc) return rc;
        rc = make_map_plain32(&tmPgo, pool_pgo, M / pool_k, K, 64, groups);
        if (rc) return rc;
    }
    GemmTnArgs args;
    args.M = M, args.N = N, args.K = K, args.BN = pl.BN, args.stages = pl.stages, args.out_bufs = pl.out_bufs, args.pair = pl.pair;
    args.a_scale = a_scale, args.a_shift = a_shift, args.z_scale = z_scale, args.z_shift = z_shift, args.partials = partials;
    args.pool_arg = pool_arg, args.pool_pgo = pool_pgo, args.pool_k = pool_k;

    cudaStream_t st = (cudaStream_t)stream;
#define MPB_TN_CASE(DT, XF, EP) \
    if (dtype == DT && xform == XF && epi == EP) return launch_tn<DT, XF, EP>(pl, tmA, tmB, tmBlo, tmC, tmZ, tmArg, tmPgo, args, st)
    MPB_TN_CASE(DT_BF16, 0, 0);
    MPB_TN_CASE(DT_BF16, 0, 1);
    MPB_TN_CASE(DT_BF16, 0, 2);
    MPB_TN_CASE(DT_BF16, 1, 0);
    MPB_TN_CASE(DT_BF16, 1, 1);
    MPB_TN_CASE(DT_BF16, 2, 0);
    MPB_TN_CASE(DT_BF16, 2, 2);
    MPB_TN_CASE(DT_TF32, 0, 0);
    MPB_TN_CASE(DT_TF32, 0, 1);
    MPB_TN_CASE(DT_TF32, 0, 2);
    MPB_TN_CASE(DT_TF32, 1, 0);
    MPB_TN_CASE(DT_TF32, 1, 1);
    MPB_TN_CASE(DT_TF32X3, 0, 0);
    MPB_TN_CASE(DT_TF32X3, 0, 1);
    MPB_TN_CASE(DT_TF32X3, 0, 2);
    MPB_TN_CASE(DT_TF32X3, 1, 0);
    MPB_TN_CASE(DT_TF32X3, 1, 1);
#undef MPB_TN_CASE
    set_error("mpb_sa_gemm_tn: combination dtype=%d xform=%d epi=%d not built", dtype, xform, epi);
    return MPB_ERR_UNSUPPORTED;
}
}  // namespace mpb

extern "C" int mpb_sa_gemm_tn(int dtype, const void *A, const void *B, const void *B_lo, void *C, int M, int N, int K,
                              const float *a_scale, const float *a_shift, int epi, float *partials, int nparts, const void *Z,
                              const float *z_scale, const float *z_shift, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE((a_scale != nullptr) == (a_shift != nullptr), "a_scale / a_shift must come together");
    return gemm_tn_impl(dtype, A, B, B_lo, C, M, N, K, a_scale ? 1 : 0, a_scale, a_shift, nullptr, nullptr, 0, epi, partials, nparts, Z, z_scale,
                        z_shift, stream);
}

// C[M,N] = dZ[M,K] * B[N,K]^T where dZ is rebuilt on the fly from the stored pre-activation Zl [M,K] (bf16) of a max-pooled
// layer: dZ = pgo * [row is the arg-max row of its (group, channel)] + negw_e[0] * Zl + negw_e[1]   (see mpb_bn_bwd_stats /
// mpb_bn_bwd_finalize_f32).  pool_k rows per group (16, 32, 64 or a multiple of 128), argmax / pgo [M / pool_k, K].
extern "C" int mpb_sa_gemm_tn_pool(int dtype, const void *Zl, const void *B, void *C, int M, int N, int K, int pool_k,
                                   const int32_t *argmax, const float *pgo, const float *negw_e, int epi, float *partials, int nparts,
                                   const void *Z, const float *z_scale, const float *z_shift, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(dtype == DT_BF16, "the pooled operand transform needs bf16 activations");
    MPB_REQUIRE(argmax && pgo && negw_e, "null pointer");
    MPB_REQUIRE(pool_k >= 16 && (128 % pool_k == 0 || pool_k % 128 == 0) && M % pool_k == 0,
                "pool_k must be 16, 32, 64 or a multiple of 128, and divide M");
    MPB_REQUIRE(epi == 0 || epi == 2, "epi must be 0 or 2");
    return gemm_tn_impl(dtype, Zl, B, nullptr, C, M, N, K, 2, negw_e, negw_e + K, argmax, pgo, pool_k, epi, partials, nparts, Z, z_scale, z_shift,
                        stream);
}

namespace mpb {
struct WgPlan {
    int n_tiles, k_tiles, m_splits, rows_per_split, stages, n_pad, threads, grid, ab;
    size_t smem;
};
static bool plan_wgrad(int dt, bool xform, int M, int N, int K, WgPlan *pl, bool zpool = false)
{
    const int epr = dt == DT_BF16 ? 64 : 32, rb = dt == DT_BF16 ? 64 : 32, kt = dt == DT_BF16 ? 256 : 128;
    const int na = dt == DT_TF32X3 ? 2 : 1;
    if (M <= 0 || N <= 0 || K <= 0 || N % 8 || K % epr) return false;
    pl->n_tiles = (N + 127) / 128, pl->k_tiles = (K + kt - 1) / kt;
    const int row_blocks = (M + rb - 1) / rb;
    const int per_sm = dt == DT_BF16 ? 2 : 1;
    int m_splits = (per_sm * sm_count()) / (pl->n_tiles * pl->k_tiles);
    m_splits = m_splits < 1 ? 1 : (m_splits > row_blocks ? row_blocks : m_splits);
    pl->rows_per_split = ((row_blocks + m_splits - 1) / m_splits) * rb;
    pl->m_splits = (M + pl->rows_per_split - 1) / pl->rows_per_split;
    const int box = rb * 128;
    pl->ab = (K < kt ? K : kt) / epr;                       // A boxes a stage really needs (was sized for kt: 2 stages where 4 fit)
    const int stage_bytes = na * (128 / epr + pl->ab) * box + (zpool ? kPoolStageBytes : 0);
    const int budget = dt == DT_BF16 ? 104 * 1024 : 200 * 1024;
    int stages = budget / stage_bytes;
    stages = stages < 2 ? 2 : (stages > 6 ? 6 : stages);
    pl->stages = stages;
    pl->n_pad = pl->n_tiles * 128;
    pl->threads = (xform || zpool || dt == DT_TF32X3) ? kWgradThreadsXf : kWgradThreads;
    pl->grid = pl->n_tiles * pl->k_tiles * pl->m_splits;
    pl->smem = (size_t)stages * stage_bytes + (xform ? 2 * kt * 4 : 0) + (zpool ? 2 * 128 * 4 : 0) + sizeof(GemmSmemTail) + 1024;
    return pl->smem <= 227 * 1024;
}
template <int DT, bool XFORM, bool ZPOOL = false>
static int launch_wg(const WgPlan &pl, const CUtensorMap &tmZ, const CUtensorMap &tmA, const CUtensorMap &tmArg, const CUtensorMap &tmPgo,
                     const WgradArgs &args, cudaStream_t st)
{
    auto kern = wgrad_kernel<DT, XFORM, ZPOOL>;
    MPB_ENSURE_DYN_SMEM(kern, 227 * 1024);
    kern<<<pl.grid, pl.threads, pl.smem, st>>>(tmZ, tmA, tmArg, tmPgo, args);
    return check_launch("wgrad_kernel");
}
static int launch_wgrad_reduce(const WgPlan &pl, const float *workspace, int K, int cout, int cin, int xyz_last, int accumulate, float *dW,
                               cudaStream_t st)
{
    int G = 1;                                           // warps per block: ~4 splits per thread, at most 32 warps
    while (G < kReduceMaxWarps && G * 4 < pl.m_splits) G <<= 1;
    wgrad_reduce_kernel<<<(cout * cin + 31) / 32, 32 * G, 0, st>>>(workspace, pl.m_splits, pl.n_pad, K, cout, cin, xyz_last, accumulate, dW);
    return check_launch("wgrad_reduce_kernel");
}
}  // namespace mpb

// Bytes of fp32 workspace the weight-gradient call needs for its per-split partial tiles.
extern "C" int64_t mpb_sa_gemm_wgrad_workspace(int dtype, int M, int N, int K, int xform)
{
    mpb::WgPlan pl;
    if (!mpb::plan_wgrad(dtype, xform != 0, M, N, K, &pl)) return -1;
    return (int64_t)pl.m_splits * pl.n_pad * K * 4;
}

namespace mpb {
static int wgrad_impl(int dtype, const void *dZ, const void *A, int M, int N, int K, const float *a_scale, const float *a_shift,
                      const int32_t *pool_arg, const float *pool_pgo, const float *pool_negw_e, int pool_k, float *workspace, int cout, int cin,
                      int xyz_last, int accumulate, float *dW, void *stream)
{
    MPB_REQUIRE(dtype >= DT_BF16 && dtype <= DT_TF32X3, "dtype must be 0 (bf16), 1 (tf32) or 2 (tf32x3)");
    MPB_REQUIRE(M > 0 && N > 0 && K > 0, "bad size");
    MPB_REQUIRE(dZ && A && dW && workspace, "null pointer");
    MPB_REQUIRE((a_scale != nullptr) == (a_shift != nullptr), "a_scale / a_shift must come together");
    MPB_REQUIRE(cout > 0 && cout <= N && cin > 0 && cin <= K, "cout / cin must fit the padded operand widths");
    MPB_REQUIRE(((uintptr_t)dZ & 15) == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "operands must be 16-byte aligned");
    const bool xform = a_scale != nullptr, zpool = pool_arg != nullptr;
    WgPlan pl;
    MPB_REQUIRE(plan_wgrad(dtype, xform, M, N, K, &pl, zpool), "unsupported shape (N multiple of 8, K multiple of 64 bf16 / 32 tf32)");
    const int esz = dtype == DT_BF16 ? 2 : 4;
    const int rb = dtype == DT_BF16 ? 64 : 32;
    CUtensorMap tmZ, tmA;
    int rc = make_map(&tmZ, esz, dZ, M, N, N, rb, dtype != DT_BF16);
    if (rc) return rc;
    rc = make_map(&tmA, esz, A, M, K, K, rb, dtype != DT_BF16);
    if (rc) return rc;
    WgradArgs args;
    args.M = M, args.N = N, args.K = K, args.k_tiles = pl.k_tiles, args.m_splits = pl.m_splits, args.rows_per_split = pl.rows_per_split;
    args.stages = pl.stages, args.a_scale = a_scale, args.a_shift = a_shift, args.partials = workspace, args.n_pad = pl.n_pad;
    args.pool_arg = pool_arg, args.pool_pgo = pool_pgo, args.pool_negw = pool_negw_e, args.pool_e = pool_negw_e ? pool_negw_e + N : nullptr;
    args.pool_k = pool_k, args.ab = pl.ab;
    CUtensorMap tmArg = tmZ, tmPgo = tmZ;
    if (zpool) {
        const int groups = pool_k >= rb ? 1 : rb / pool_k;
        rc = make_map_plain32(&tmArg, pool_arg, M / pool_k, N, 128, groups);
        if (rc) return rc;
        rc = make_map_plain32(&tmPgo, pool_pgo, M / pool_k, N, 128, groups);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = MPB_ERR_UNSUPPORTED;
#define MPB_WG_CASE(DT, XF) \
    if (dtype == DT && xform == XF && !zpool) rc = launch_wg<DT, XF>(pl, tmZ, tmA, tmArg, tmPgo, args, st)
    MPB_WG_CASE(DT_BF16, false);
    MPB_WG_CASE(DT_BF16, true);
    MPB_WG_CASE(DT_TF32, false);
    MPB_WG_CASE(DT_TF32, true);
    MPB_WG_CASE(DT_TF32X3, false);
    MPB_WG_CASE(DT_TF32X3, true);
#undef MPB_WG_CASE
    if (zpool && dtype == DT_BF16 && xform) rc = launch_wg<DT_BF16, true, true>(pl, tmZ, tmA, tmArg, tmPgo, args, st);
    if (zpool && dtype == DT_BF16 && !xform) rc = launch_wg<DT_BF16, false, true>(pl, tmZ, tmA, tmArg, tmPgo, args, st);
    if (rc == MPB_ERR_UNSUPPORTED) set_error("mpb_sa_gemm_wgrad: combination dtype=%d xform=%d pool=%d not built", dtype, (int)xform, (int)zpool);
    if (rc) return rc;
    if (accumulate < 0) return MPB_OK;                   // partial tiles only: the caller reduces them with mpb_sa_gemm_wgrad_reduce
    return launch_wgrad_reduce(pl, workspace, K, cout, cin, xyz_last, accumulate, dW, st);
}
}  // namespace mpb

extern "C" int mpb_sa_gemm_wgrad(int dtype, const void *dZ, const void *A, int M, int N, int K, const float *a_scale,
                                 const float *a_shift, float *workspace, int cout, int cin, int xyz_last, int accumulate, float *dW,
                                 void *stream)
{
    return mpb::wgrad_impl(dtype, dZ, A, M, N, K, a_scale, a_shift, nullptr, nullptr, nullptr, 0, workspace, cout, cin, xyz_last, accumulate, dW,
                           stream);
}

// dW = dZ^T * f(A) with dZ rebuilt on the fly from the stored pre-activation Zl [M,N] (bf16) of a max-pooled layer, exactly as in
// mpb_sa_gemm_tn_pool (a 64-row operand box holds whole pooling groups or part of one: pool_k = 16, 32 or a multiple of 64).
extern "C" int mpb_sa_gemm_wgrad_pool(int dtype, const void *Zl, const void *A, int M, int N, int K, const float *a_scale,
                                      const float *a_shift, int pool_k, const int32_t *argmax, const float *pgo, const float *negw_e,
                                      float *workspace, int cout, int cin, int xyz_last, int accumulate, float *dW, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(dtype == DT_BF16, "the pooled operand transform needs bf16 activations");
    MPB_REQUIRE(argmax && pgo && negw_e, "null pointer");
    MPB_REQUIRE(pool_k >= 16 && (64 % pool_k == 0 || pool_k % 64 == 0) && M % pool_k == 0,
                "pool_k must be 16, 32 or a multiple of 64, and divide M");
    MPB_REQUIRE(N % 64 == 0, "N must be a multiple of 64 (padded channels)");
    return wgrad_impl(dtype, Zl, A, M, N, K, a_scale, a_shift, argmax, pgo, negw_e, pool_k, workspace, cout, cin, xyz_last, accumulate, dW, stream);
}

// Second half of mpb_sa_gemm_wgrad(accumulate = -1): the fixed-order sum of the per-split partial tiles.  Separate entry point
// so that the caller can issue it on another stream (it only feeds the optimizer, the next GEMM need not wait for it).
extern "C" int mpb_sa_gemm_wgrad_reduce(int dtype, int M, int N, int K, int xform, const float *workspace, int cout, int cin,
                                        int xyz_last, int accumulate, float *dW, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(dtype >= DT_BF16 && dtype <= DT_TF32X3, "dtype must be 0 (bf16), 1 (tf32) or 2 (tf32x3)");
    MPB_REQUIRE(M > 0 && N > 0 && K > 0 && workspace && dW, "bad argument");
    MPB_REQUIRE(cout > 0 && cout <= N && cin > 0 && cin <= K, "cout / cin must fit the padded operand widths");
    WgPlan pl;
    MPB_REQUIRE(plan_wgrad(dtype, xform != 0, M, N, K, &pl), "unsupported shape");
    return launch_wgrad_reduce(pl, workspace, K, cout, cin, xyz_last, accumulate != 0, dW, (cudaStream_t)stream);
}

// Developer aid (MPB_MBAR_DEBUG builds): where the first timed-out mbarrier wait happened; all zeros otherwise.
// out8 = {source line in sa_gemm.cu, blockIdx.x, threadIdx.x, gridDim.x, parity, barrier shared address, blockDim.x, 0}
extern "C" int mpb_debug_mbar_state(int *out8, int reset)
{
    using namespace mpb;
    MPB_REQUIRE(out8, "null pointer");
    for (int i = 0; i < 8; ++i) out8[i] = 0;
#ifdef MPB_MBAR_DEBUG
    MPB_CUDA(cudaMemcpyFromSymbol(out8, tc::g_mbar_dbg, sizeof(int) * 8));
    if (reset) {
        int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        MPB_CUDA(cudaMemcpyToSymbol(tc::g_mbar_dbg, z, sizeof(z)));
        MPB_CUDA(cudaMemcpyToSymbol(tc::g_mbar_abort, z, sizeof(int)));
    }
#else
    (void)reset;
#endif
    return MPB_OK;
}
