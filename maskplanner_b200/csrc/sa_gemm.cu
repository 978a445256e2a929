// Shared-MLP GEMMs of PointNetSetAbstraction on 5th-gen tensor cores (sm_100a).
// Reference: models/pointnet2_utils.py:210-212 -- `conv(new_points)` with a 1x1 Conv2d over the grouped
// tensor [B, C, K, S] is a row-wise GEMM  Z[M, Cout] = A[M, Cin] * W[Cout, Cin]^T  with M = B*S*K rows.
//
//   gemm_tn_kernel   : C[M,N] = A[M,K] * B[N,K]^T      forward (B = W) and dgrad (A = dZ, B = W^T)
//   wgrad_kernel     : dW[N,K] += dZ[M,N]^T * A[M,K]   contraction over the M rows, split across CTAs
//
// Both: bf16 operands, fp32 accumulation in TMEM, operands staged by TMA (SWIZZLE_128B) through a
// multi-stage mbarrier ring, tcgen05.mma (cta_group::1, UMMA 128 x N x 16) issued by one elected thread,
// epilogue warps drain TMEM with tcgen05.ld.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM
// allocator, 4..7 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31).
//
// gemm_tn: operands are K-major ([rows][64 bf16 = 128 B] swizzle atoms, SBO = 1024 B); persistent over
// (m-tile, n-tile) pairs with a double-buffered TMEM accumulator so the epilogue of tile i overlaps the
// MMAs of tile i+1.
// wgrad:   both operands are read exactly as they sit in HBM (row-major [M, C]) and fed to the tensor
// core as MN-major operands (contraction index = row), so no transposed copy of dZ or A is ever made.
#include "common.cuh"
#include "tc_common.cuh"

namespace mpb {

using namespace tc;

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

// bf16 row-major [rows, cols] (row stride = ld elements); box = 64 columns (128 B) x box_rows, SWIZZLE_128B.
static int make_map_bf16(CUtensorMap *map, const void *ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return MPB_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d ptr=%p", (int)r, (long long)rows,
                  (long long)cols, (long long)ld, box_rows, ptr);
        return MPB_ERR_CUDA;
    }
    return MPB_OK;
}

constexpr int kGemmThreads = 256;     // wgrad: producer, MMA, allocator, spare + one epilogue warpgroup
constexpr int kGemmTnThreads = 384;   // gemm_tn: the same + a second epilogue warpgroup
constexpr int kTileM = 128;
constexpr int kTileK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kABytes = kTileM * 128; // 16 KB per A stage
constexpr uint32_t kTmemCols = 512;

struct GemmSmemTail {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
}

// Column sums of a 32 x 32 block held one ROW per lane (v[c] = this lane's value in column c): a transpose-reduce
// butterfly.  Each step halves the live registers -- a lane keeps the half of the columns selected by one bit of its
// lane index and receives the other lanes' contributions for that half -- so after 16+8+4+2+1 = 31 shuffles lane L
// holds the sum over all 32 rows of column L.  No shared memory, no atomics.
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

// STATS: additionally accumulate per-column sum / sum of squares of the (bf16-rounded) outputs -- the training-mode
// BatchNorm statistics of the layer (reference :210-212) -- in the epilogue, so Z is not read again for them.
// Each epilogue warp keeps its own [2][BN] accumulator in shared memory (lane L owns column 32j+L: no conflicts, no
// atomics) and writes it as one partial row at the end: partials[(blockIdx.x*8 + epilogue warp)][2][N].
template <bool OUT_F32, bool STATS>
__global__ void __launch_bounds__(kGemmTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, void *__restrict__ Cout, int M, int N, int K, int BN, int ldc, int stages,
               float *__restrict__ stat_partials)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B tiles need 1024-B alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t stage_bytes = kABytes + (uint32_t)BN * 128u;
    // [stages x (A | B)] [bf16 path: 2 warpgroups x 2 output staging tiles of 128 rows x 128 B] [barriers]
    uint8_t *staging = smem + (size_t)stages * stage_bytes;
    float *stat_acc = reinterpret_cast<float *>(staging + (OUT_F32 ? 0 : 4 * kABytes));          // STATS: [8 warps][2][BN]
    GemmSmemTail *tail = reinterpret_cast<GemmSmemTail *>(reinterpret_cast<uint8_t *>(stat_acc) + (STATS ? 8 * 2 * 256 * 4 : 0));
    const int num_kb = K / kTileK;
    const int tiles_m = (M + kTileM - 1) / kTileM, tiles_n = N / BN;
    const int total = tiles_m * tiles_n;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmA);
        prefetch_tensormap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&tail->full[s], 1);
            mbar_init(&tail->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tail->tfull[a], 1);
            mbar_init(&tail->tempty[a], 4);
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
                    mbar_arrive_expect_tx(&tail->full[stage], stage_bytes);
                    uint8_t *sa = smem + (size_t)stage * stage_bytes;
                    tma_load_2d(sa, &tmA, kb * kTileK, m0, &tail->full[stage]);
                    tma_load_2d(sa + kABytes, &tmB, kb * kTileK, n0, &tail->full[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16((uint32_t)BN, false, false);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                mbar_wait(&tail->tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&tail->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(sa + kABytes, 16, 1024);
#pragma unroll
                    for (int k = 0; k < kTileK / 16; ++k)   // UMMA_K = 16 bf16 = 32 B inside the 128-B swizzle row
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&tail->empty[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
                umma_commit(&tail->tfull[acc]);
                if ((acc ^= 1) == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // Two epilogue warpgroups: warps 4-7 drain accumulator 0 (even tiles of this CTA), warps 8-11 accumulator 1
        // (odd tiles), so TMEM->register->global of one tile overlaps the next tile's drain as well as its MMAs.
        const int ew = warp & 3;           // the TMEM lane quarter this warp may read (warp % 4)
        const int acc = (warp - 4) >> 2;   // accumulator buffer owned by this warpgroup
        uint32_t acc_phase = 0;
        int t = 0;
        int nstore = 0;                                 // 64-column blocks stored so far by this warpgroup (staging ring of 2)
        const bool issuer = ew == 0 && lane == 0;       // the one thread of the warpgroup that owns its bulk-store groups
        const int r_in_tile = ew * 32 + lane;
        float *my_stat = stat_acc + (size_t)(warp - 4) * 2 * BN;
        if (STATS)
            for (int c = lane; c < 2 * BN; c += 32) my_stat[c] = 0.f;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++t) {
            if ((t & 1) != acc) continue;
            const int m0 = (tile / tiles_n) * kTileM, n0 = (tile % tiles_n) * BN;
            mbar_wait(&tail->tfull[acc], acc_phase);
            tc_fence_after();
            const int row = m0 + r_in_tile;
            const uint32_t taddr = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(ew * 32) << 16);
            if (OUT_F32) {
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + (uint32_t)c0, r);
                    if (row < M) {
                        float4 *dst = reinterpret_cast<float4 *>(static_cast<float *>(Cout) + (size_t)row * ldc + n0 + c0);
#pragma unroll
                        for (int v = 0; v < 8; ++v)
                            dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                                 __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tail->tempty[acc]);
            } else {
                // bf16: TMEM -> registers -> swizzled staging tile in shared memory -> ONE TMA store per 128 x 64 block
                // (full 128-byte lines to L2 instead of 32 half-sector writes per warp instruction)
                for (int c0 = 0; c0 < BN; c0 += 64, ++nstore) {
                    uint8_t *buf = staging + (size_t)(acc * 2 + (nstore & 1)) * kABytes;
                    if (issuer) bulk_wait_read<1>();                 // the store that used this buffer two blocks ago has drained
                    named_bar_sync(1 + acc, 128);
                    uint8_t *rowp = buf + r_in_tile * 128;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (c0 + 32 * h < BN) {
                            uint32_t r[32];
                            tmem_ld_32x32(taddr + (uint32_t)(c0 + 32 * h), r);
                            uint32_t pk[16];
#pragma unroll
                            for (int v = 0; v < 16; ++v) pk[v] = pack_bf16(__uint_as_float(r[2 * v]), __uint_as_float(r[2 * v + 1]));
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                const int chunk = (4 * h + v) ^ (r_in_tile & 7);   // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
                                *reinterpret_cast<uint4 *>(rowp + chunk * 16) = make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]);
                            }
                            if (STATS) {   // statistics of exactly what was stored (rows past M are zero: TMA zero-fills A)
                                float zv[32], zq[32];
#pragma unroll
                                for (int v = 0; v < 16; ++v) {
                                    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&pk[v]));
                                    zv[2 * v] = f.x, zv[2 * v + 1] = f.y;
                                    zq[2 * v] = f.x * f.x, zq[2 * v + 1] = f.y * f.y;
                                }
                                const float cs = warp_column_sum_32x32(zv, lane), cq = warp_column_sum_32x32(zq, lane);
                                my_stat[c0 + 32 * h + lane] += cs;
                                my_stat[BN + c0 + 32 * h + lane] += cq;
                            }
                        }
                    }
                    if (c0 + 64 >= BN) {                               // accumulator fully drained: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tail->tempty[acc]);
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1 + acc, 128);
                    if (issuer) {
                        tma_store_2d(&tmC, buf, n0 + c0, m0);
                        bulk_commit();
                    }
                }
            }
            acc_phase ^= 1;
        }
        if (!OUT_F32 && issuer) bulk_wait<0>();          // all stores complete before the CTA (and its shared memory) goes away
        if (STATS) {
            float *dst = stat_partials + ((size_t)blockIdx.x * 8 + (warp - 4)) * 2 * N;      // tiles_n == 1 here: BN == N
            for (int c = lane; c < 2 * BN; c += 32) dst[c] = my_stat[c];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<kTmemCols>(tmem_base);
}

// dW[n0 + i, k0 + j] += sum_{m in this CTA's row range} dZ[m, n0 + i] * A[m, k0 + j]
// blockIdx.x = ((n_tile * k_tiles) + k_tile) * m_splits + m_split
__global__ void __launch_bounds__(kGemmThreads, 2)
wgrad_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmA, float *__restrict__ dW, int M,
             int N, int K, int ldw, int k_tiles, int m_splits, int rows_per_split, int stages)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ms = blockIdx.x % m_splits;
    const int kt = (blockIdx.x / m_splits) % k_tiles;
    const int nt = blockIdx.x / (m_splits * k_tiles);
    const int n0 = nt * 128, k0 = kt * 256;
    const int NU = min(256, K - k0);                 // multiple of 64
    const int a_boxes = min(2, (N - n0 + 63) / 64);   // 64-channel boxes of dZ that hold real data
    const int b_boxes = NU / 64;
    constexpr int kBox = 64 * 128;                    // 64 rows x 128 B
    const uint32_t stage_bytes = (uint32_t)(2 + b_boxes) * kBox;
    GemmSmemTail *tail = reinterpret_cast<GemmSmemTail *>(smem + (size_t)stages * stage_bytes);
    const int m_begin = ms * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    const int num_rb = m_end > m_begin ? (m_end - m_begin + 63) / 64 : 0;

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tmZ);
        prefetch_tensormap(&tmA);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&tail->full[s], 1);
            mbar_init(&tail->empty[s], 1);
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
                    mbar_arrive_expect_tx(&tail->full[stage], (uint32_t)(a_boxes + b_boxes) * kBox);
                    uint8_t *s = smem + (size_t)stage * stage_bytes;
                    const int r0 = m_begin + rb * 64;  // rows past m_end but < M belong to the next split: mask below
                    for (int b = 0; b < a_boxes; ++b) tma_load_2d(s + b * kBox, &tmZ, n0 + b * 64, r0, &tail->full[stage]);
                    for (int b = 0; b < b_boxes; ++b) tma_load_2d(s + (2 + b) * kBox, &tmA, k0 + b * 64, r0, &tail->full[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = make_idesc_bf16((uint32_t)NU, true, true);
                int stage = 0;
                uint32_t phase = 0;
                for (int rb = 0; rb < num_rb; ++rb) {
                    mbar_wait(&tail->full[stage], phase);
                    tc_fence_after();
                    const uint32_t s = smem_u32(smem + (size_t)stage * stage_bytes);
                    // MN-major SWIZZLE_128B: LBO = distance between 64-channel column blocks (one box),
                    // SBO = distance between 8-row groups along the contraction (1024 B)
                    const uint64_t adesc = make_smem_desc_sw128(s, kBox, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(s + 2 * kBox, kBox, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)   // 16 contraction rows per MMA = two 8-row groups = 2048 B
                        umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (rb | k) ? 1u : 0u);
                    umma_commit(&tail->empty[stage]);
                    if (++stage == stages) stage = 0, phase ^= 1;
                }
                umma_commit(&tail->tfull[0]);
            }
        } else if (warp >= 4) {
            const int ew = warp - 4;
            mbar_wait(&tail->tfull[0], 0);
            tc_fence_after();
            const int row = n0 + ew * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16);
            for (int c0 = 0; c0 < NU; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(taddr + (uint32_t)c0, r);
                if (row < N) {
                    float *dst = dW + (size_t)row * ldw + k0 + c0;
#pragma unroll
                    for (int v = 0; v < 32; v += 4)   // 16-byte vector reductions (RED.E.ADD.F32x4)
                        atomicAdd(reinterpret_cast<float4 *>(dst + v),
                                  make_float4(__uint_as_float(r[v]), __uint_as_float(r[v + 1]), __uint_as_float(r[v + 2]),
                                              __uint_as_float(r[v + 3])));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<256>(tmem_base);
}

// Columns per tile: the whole N when it fits one UMMA (N <= 256, multiple of 32); otherwise the largest multiple of
// 64 dividing N (the bf16 epilogue stores whole 64-column blocks, which must not straddle two tiles).
static int pick_bn(int N)
{
    if (N <= 256) return N % 32 == 0 ? N : 0;
    for (int bn = 256; bn >= 64; bn -= 64)
        if (N % bn == 0) return bn;
    return 0;
}

}  // namespace mpb

namespace mpb {
static int gemm_tn_launch(const void *A, const void *B, void *C, int M, int N, int K, int out_fp32, float *stat_partials, int nparts,
                          void *stream);
}

extern "C" int mpb_gemm_bf16_tn(const void *A, const void *B, void *C, int M, int N, int K, int out_fp32, void *stream)
{
    return mpb::gemm_tn_launch(A, B, C, M, N, K, out_fp32, nullptr, 0, stream);
}

// Number of [2][N] partial rows mpb_gemm_bf16_tn_stats writes (0: this shape cannot fuse the statistics).
extern "C" int mpb_gemm_tn_stat_partials(int M, int N, int K)
{
    if (M <= 0 || N <= 0 || N > 256 || N % 32 || K % 64) return 0;
    const int tiles = (M + mpb::kTileM - 1) / mpb::kTileM;
    return 8 * (tiles < mpb::sm_count() ? tiles : mpb::sm_count());
}

extern "C" int mpb_gemm_bf16_tn_stats(const void *A, const void *B, void *C, int M, int N, int K, float *partials, int nparts,
                                      void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(partials && nparts > 0 && nparts == mpb_gemm_tn_stat_partials(M, N, K), "partials / nparts mismatch");
    return gemm_tn_launch(A, B, C, M, N, K, 0, partials, nparts, stream);
}

static int mpb::gemm_tn_launch(const void *A, const void *B, void *C, int M, int N, int K, int out_fp32, float *stat_partials,
                               int nparts, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(M >= 0 && N > 0 && K > 0, "bad size");
    if (M == 0) return MPB_OK;
    MPB_REQUIRE(A && B && C, "null pointer");
    MPB_REQUIRE(K % 64 == 0, "K must be a multiple of 64 (pad the operands)");
    const int BN = pick_bn(N);
    MPB_REQUIRE(BN > 0, "N must be a multiple of 32");
    MPB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0, "operands must be 16-byte aligned");
    MPB_REQUIRE(BN == N || BN % 64 == 0, "N > 256 must be a multiple of 64");
    CUtensorMap tmA, tmB, tmC;
    int rc = make_map_bf16(&tmA, A, M, K, K, kTileM);
    if (rc) return rc;
    rc = make_map_bf16(&tmB, B, N, K, K, BN);
    if (rc) return rc;
    const int stage_bytes = kABytes + BN * 128;
    const int staging_bytes = (out_fp32 ? 0 : 4 * kABytes)    // bf16 output: 2 warpgroups x 2 staging tiles for the TMA stores
                              + (stat_partials ? 8 * 2 * 256 * 4 : 0);   // fused statistics: 8 warps x [2][256] floats
    int stages = (224 * 1024 - staging_bytes) / stage_bytes;
    stages = stages > 8 ? 8 : stages;
    const size_t smem = (size_t)stages * stage_bytes + staging_bytes + sizeof(GemmSmemTail) + 1024;
    const int tiles = ((M + kTileM - 1) / kTileM) * (N / BN);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    cudaStream_t st = (cudaStream_t)stream;
    if (out_fp32) {
        auto kern = gemm_tn_kernel<true, false>;
        MPB_ENSURE_DYN_SMEM(kern, 227 * 1024);
        kern<<<grid, kGemmTnThreads, smem, st>>>(tmA, tmB, tmA, C, M, N, K, BN, N, stages, nullptr);
    } else {
        rc = make_map_bf16(&tmC, C, M, N, N, kTileM);          // output tiles: 128 rows x 64 columns, SWIZZLE_128B
        if (rc) return rc;
        if (stat_partials) {
            MPB_REQUIRE(BN == N && nparts == 8 * grid, "fused statistics need a single column tile");
            auto kern = gemm_tn_kernel<false, true>;
            MPB_ENSURE_DYN_SMEM(kern, 227 * 1024);
            kern<<<grid, kGemmTnThreads, smem, st>>>(tmA, tmB, tmC, C, M, N, K, BN, N, stages, stat_partials);
        } else {
            auto kern = gemm_tn_kernel<false, false>;
            MPB_ENSURE_DYN_SMEM(kern, 227 * 1024);
            kern<<<grid, kGemmTnThreads, smem, st>>>(tmA, tmB, tmC, C, M, N, K, BN, N, stages, nullptr);
        }
    }
    return check_launch("gemm_tn_kernel");
}

extern "C" int mpb_gemm_bf16_wgrad(const void *dZ, const void *A, float *dW, int M, int N, int K, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(M >= 0 && N > 0 && K > 0, "bad size");
    if (M == 0) return MPB_OK;
    MPB_REQUIRE(dZ && A && dW, "null pointer");
    MPB_REQUIRE(N % 8 == 0 && K % 64 == 0, "N must be a multiple of 8 and K a multiple of 64");
    MPB_REQUIRE(((uintptr_t)dZ & 15) == 0 && ((uintptr_t)A & 15) == 0, "operands must be 16-byte aligned");
    CUtensorMap tmZ, tmA;
    int rc = make_map_bf16(&tmZ, dZ, M, N, N, 64);
    if (rc) return rc;
    rc = make_map_bf16(&tmA, A, M, K, K, 64);
    if (rc) return rc;
    const int n_tiles = (N + 127) / 128, k_tiles = (K + 255) / 256;
    const int row_blocks = (M + 63) / 64;
    int m_splits = (2 * sm_count()) / (n_tiles * k_tiles);
    m_splits = m_splits < 1 ? 1 : (m_splits > row_blocks ? row_blocks : m_splits);
    const int rows_per_split = ((row_blocks + m_splits - 1) / m_splits) * 64;
    m_splits = (M + rows_per_split - 1) / rows_per_split;
    // stage = 64 contraction rows x (128 dZ channels + up to 256 A channels); sized so that two CTAs share an SM
    // (one CTA's TMEM drain + reductions overlap the other's loads)
    const int stage_bytes = (2 + (K < 256 ? K : 256) / 64) * 64 * 128;
    int stages = (100 * 1024) / stage_bytes;
    stages = stages < 2 ? 2 : (stages > 4 ? 4 : stages);
    const size_t smem = (size_t)stages * stage_bytes + sizeof(GemmSmemTail) + 1024;
    MPB_ENSURE_DYN_SMEM(wgrad_kernel, 227 * 1024);
    wgrad_kernel<<<n_tiles * k_tiles * m_splits, kGemmThreads, smem, (cudaStream_t)stream>>>(tmZ, tmA, dW, M, N, K, K, k_tiles, m_splits,
                                                                                            rows_per_split, stages);
    return check_launch("wgrad_kernel");
}
