// Chamfer nearest-neighbour search (forward + backward) for sm_100a.
// Reference: pytorch3d_chamfer.py:257-258 -> pytorch3d.ops.knn.knn_points (third-party, K = 1),
//            :138-149 (padded-length scan), :205-206 (K = 2 branches).
//
// Tiled all-pairs min/arg-min that never materialises the P1 x P2 distance matrix: a CTA owns
// THREADS*R query points held in REGISTERS (R per thread), streams the target set through a
// shared-memory tile (every thread reads the same target row => broadcast LDS.128, reused by the
// R register-resident queries) and keeps a running (min, arg-min) per query.  Both chamfer
// directions run in ONE launch (blockIdx.z), and either can be skipped.  HBM traffic is the
// compulsory (4D+12)*N*(P1+P2) bytes; the kernel is bound by the fma pipe: D packed fp32x2
// instruction pairs (FADD2 + FFMA2) per TWO (query, target) pairs plus one FMNMX3.
//
// Arithmetic contract: d = sum_k (q_k - t_k)^2 accumulated with FMA over k = 0..D-1 in order
// (what nvcc emits for pytorch3d's CUDA loop); strict '<' while scanning targets in index order
// => lowest index wins ties; only the first len_t targets are candidates; query rows >= len_q
// are written as (0, 0).
#include "common.cuh"

namespace mpb {

constexpr int kChThreads = 128;
constexpr int kChChunkPairs = 4;  // index bookkeeping granularity: 4 target pairs = 8 targets

// Shared-memory tile of NEGATED targets, two targets per row so that one 64-bit register pair holds
// (-t[2p][d], -t[2p+1][d]): row p = { d0.lo d0.hi d1.lo d1.hi ... } padded to a float4 multiple.
template <int D>
struct ChTile {
    static constexpr int ROW = (2 * D + 3) / 4 * 4;                                  // floats per target pair
    static constexpr int PAIRS = (4096 / ROW) / kChChunkPairs * kChChunkPairs;       // ~16 KB tile
    static constexpr int TILE = 2 * PAIRS;                                           // targets per tile
    // Coordinates whose differences are taken with two scalar FADDs instead of one FADD2.  Measured on
    // B200: every split other than 0 is slower (29.8 vs 27.8 ms at 64k x 64k x 32, D = 3), i.e. FADD shares
    // the fma pipe's issue budget, so everything stays packed.
    static constexpr int SCALAR_DIMS = 0;
};

// VW consecutive floats of a target row, negated (targets at infinity when the row does not exist).
template <int VW>
__device__ __forceinline__ void load_neg_row(const float *__restrict__ src, bool exists, bool vec_ok, float (&o)[VW])
{
    if (!exists) {
#pragma unroll
        for (int k = 0; k < VW; ++k) o[k] = INFINITY;
    } else if (VW == 4 && vec_ok) {
        const float4 v = *reinterpret_cast<const float4 *>(src);
        o[0] = -v.x, o[1 % VW] = -v.y, o[2 % VW] = -v.z, o[3 % VW] = -v.w;
    } else if (VW == 2 && vec_ok) {
        const float2 v = *reinterpret_cast<const float2 *>(src);
        o[0] = -v.x, o[1 % VW] = -v.y;
    } else {
#pragma unroll
        for (int k = 0; k < VW; ++k) o[k] = -src[k];
    }
}

// One squared distance with the contract's rounding: df = q - t, acc = fma(df, df, acc), k ascending.
template <int D>
__device__ __forceinline__ float chamfer_dist_scalar(const float (&q)[D], const float *__restrict__ t)
{
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const float df = q[d] - t[d];
        acc = __fmaf_rn(df, df, acc);
    }
    return acc;
}

// The pair loop runs on packed fp32x2 math (FADD2 / FFMA2: two targets per instruction, each lane
// IEEE-identical to the scalar form above since q + (-t) == q - t) and tracks only the running
// MINIMUM (one FMNMX per pair on the alu pipe instead of compare + two selects); the arg-min is kept
// at chunk granularity ("the first chunk of 8 targets in which the minimum strictly dropped to its
// final value") and resolved after the scan by recomputing that one chunk: the first target whose
// distance equals the minimum, i.e. exactly what a strict '<' scan in index order selects.
template <int D, int R>
__global__ void __launch_bounds__(kChThreads)
chamfer_nn_kernel(const float *__restrict__ x, const float *__restrict__ y, int P1, int P2,
                  const int64_t *__restrict__ x_len, const int64_t *__restrict__ y_len, float *__restrict__ dist_x,
                  int64_t *__restrict__ idx_x, float *__restrict__ dist_y, int64_t *__restrict__ idx_y, int dir0)
{
    constexpr int ROW = ChTile<D>::ROW, PAIRS = ChTile<D>::PAIRS, TILE = ChTile<D>::TILE;
    constexpr int CH = kChChunkPairs;
    constexpr int VW = (D % 4 == 0) ? 4 : (D % 2 == 0) ? 2 : 1, G = D / VW;
    constexpr int NS = ChTile<D>::SCALAR_DIMS;
    __shared__ __align__(16) float tile[PAIRS * ROW];
    const int dir = dir0 + blockIdx.z;
    const int n = blockIdx.y;
    const float *Q = dir == 0 ? x : y;
    const float *T = dir == 0 ? y : x;
    const int Pq = dir == 0 ? P1 : P2, Pt = dir == 0 ? P2 : P1;
    const int64_t *qlen = dir == 0 ? x_len : y_len, *tlen = dir == 0 ? y_len : x_len;
    float *od = dir == 0 ? dist_x : dist_y;
    int64_t *oi = dir == 0 ? idx_x : idx_y;
    const int q0 = blockIdx.x * (kChThreads * R);
    if (q0 >= Pq) return;
    long long lq = qlen ? qlen[n] : Pq, lt = tlen ? tlen[n] : Pt;
    lq = lq < 0 ? 0 : (lq > Pq ? Pq : lq);
    lt = lt < 0 ? 0 : (lt > Pt ? Pt : lt);

    float q[R][D], best[R], prev[R];
    int bc[R];  // chunk (of 2*CH targets, counted from target 0) holding the arg-min
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = q0 + r * kChThreads + threadIdx.x;
        best[r] = prev[r] = INFINITY;
        bc[r] = 0;
        const float *qp = Q + ((int64_t)n * Pq + (i < Pq ? i : 0)) * D;
#pragma unroll
        for (int d = 0; d < D; ++d) q[r][d] = qp[d];
    }
    const float *Tn = T + (int64_t)n * Pt * D;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(Tn) & (VW * 4 - 1)) == 0;  // row pitch D*4 keeps it
    if (q0 < lq) {
        for (int t0 = 0; t0 < (int)lt; t0 += TILE) {
            const int tn = min(TILE, (int)lt - t0);
            const int nchunks = (tn + 2 * CH - 1) / (2 * CH);
            __syncthreads();
            // one (target pair, column group) per thread: VW columns of rows 2p and 2p+1, negated and
            // interleaved; rows past the end become targets at infinity (distance +inf, never the minimum)
            for (int e = threadIdx.x; e < nchunks * CH * G; e += kChThreads) {
                const int p = e / G, g = e - p * G;
                float a[VW], b[VW];
                load_neg_row<VW>(Tn + ((int64_t)t0 + 2 * p) * D + g * VW, 2 * p < tn, vec_ok, a);
                load_neg_row<VW>(Tn + ((int64_t)t0 + 2 * p + 1) * D + g * VW, 2 * p + 1 < tn, vec_ok, b);
                float2 *dst = reinterpret_cast<float2 *>(tile + p * ROW + 2 * g * VW);
#pragma unroll
                for (int k = 0; k < VW; ++k) dst[k] = make_float2(a[k], b[k]);
            }
            __syncthreads();
            const int chunk0 = t0 / (2 * CH);
            for (int c = 0; c < nchunks; ++c) {
#pragma unroll
                for (int p = 0; p < CH; ++p) {
                    float2 t2[ROW / 2];
                    const float4 *rowp = reinterpret_cast<const float4 *>(tile + (c * CH + p) * ROW);
#pragma unroll
                    for (int v = 0; v < ROW / 4; ++v) {
                        const float4 f = rowp[v];
                        t2[2 * v] = make_float2(f.x, f.y);
                        t2[2 * v + 1] = make_float2(f.z, f.w);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float2 acc;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            float2 df;
                            if (d < D - NS) {
                                df = __fadd2_rn(make_float2(q[r][d], q[r][d]), t2[d]);
                            } else {
                                df.x = __fadd_rn(q[r][d], t2[d].x);
                                df.y = __fadd_rn(q[r][d], t2[d].y);
                            }
                            acc = d == 0 ? __fmul2_rn(df, df) : __ffma2_rn(df, df, acc);
                        }
                        best[r] = fminf(fminf(best[r], acc.x), acc.y);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (best[r] < prev[r]) {
                        prev[r] = best[r];
                        bc[r] = chunk0 + c;
                    }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = q0 + r * kChThreads + threadIdx.x;
        if (i < Pq) {
            const bool valid = i < lq && lt > 0;
            int bi = 0;
            if (valid && best[r] < INFINITY) {
                const int js = bc[r] * (2 * CH), je = min(js + 2 * CH, (int)lt);
                bi = js;
                for (int j = je - 1; j >= js; --j)  // descending: the last hit kept is the lowest index
                    if (chamfer_dist_scalar<D>(q[r], Tn + (int64_t)j * D) == best[r]) bi = j;
            }
            od[(int64_t)n * Pq + i] = valid ? best[r] : 0.f;
            oi[(int64_t)n * Pq + i] = valid ? bi : 0;
        }
    }
}

// Runtime-D, general-K search (K <= KMAX): one thread per query, query rows and target tile in
// shared memory.  Serves the K = 2 branches of the wrapper and any D without a specialisation.
template <int KMAX>
__global__ void __launch_bounds__(kChThreads)
knn_generic_kernel(const float *__restrict__ p1, const float *__restrict__ p2, int P1, int P2, int D,
                   const int64_t *__restrict__ len1, const int64_t *__restrict__ len2, int K,
                   float *__restrict__ dists, int64_t *__restrict__ idx, int tile_rows)
{
    extern __shared__ float sm[];
    const int QS = D + 1;  // odd-ish stride: threads hit different banks when walking their own row
    float *sq = sm;                          // [kChThreads][QS]
    float *st = sm + kChThreads * QS;        // [tile_rows][D]
    const int n = blockIdx.y;
    const int i = blockIdx.x * kChThreads + threadIdx.x;
    long long l1 = len1 ? len1[n] : P1, l2 = len2 ? len2[n] : P2;
    l1 = l1 < 0 ? 0 : (l1 > P1 ? P1 : l1);
    l2 = l2 < 0 ? 0 : (l2 > P2 ? P2 : l2);
    const int rows = min(kChThreads, P1 - blockIdx.x * kChThreads);
    for (int e = threadIdx.x; e < rows * D; e += kChThreads) {
        const int r = e / D, c = e - r * D;
        sq[r * QS + c] = p1[((int64_t)n * P1 + blockIdx.x * kChThreads + r) * D + c];
    }
    float bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) bd[k] = INFINITY, bi[k] = 0;
    float thr = INFINITY;
    int have = 0;
    const float *T = p2 + (int64_t)n * P2 * D;
    const float *myq = sq + threadIdx.x * QS;
    for (int t0 = 0; t0 < (int)l2; t0 += tile_rows) {
        const int tn = min(tile_rows, (int)l2 - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < tn * D; e += kChThreads) st[e] = T[(int64_t)t0 * D + e];
        __syncthreads();
        if (i < l1) {
            for (int j = 0; j < tn; ++j) {
                float acc = 0.f;
                const float *t = st + j * D;
                for (int d = 0; d < D; ++d) {
                    const float df = myq[d] - t[d];
                    acc = __fmaf_rn(df, df, acc);
                }
                if (acc < thr) {
                    float cd = acc;
                    int ci = t0 + j;
                    bool shifting = false;
                    have = have < K ? have + 1 : have;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) {
                        if (k < K) {
                            const bool sw = shifting || cd < bd[k];
                            const float td = bd[k];
                            const int ti = bi[k];
                            bd[k] = sw ? cd : td;
                            bi[k] = sw ? ci : ti;
                            cd = sw ? td : cd;
                            ci = sw ? ti : ci;
                            shifting = sw;
                            if (k == K - 1) thr = bd[k];
                        }
                    }
                }
            }
        }
    }
    if (i < P1) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                const bool ok = i < l1 && k < have;
                dists[((int64_t)n * P1 + i) * K + k] = ok ? bd[k] : 0.f;
                idx[((int64_t)n * P1 + i) * K + k] = ok ? bi[k] : 0;
            }
    }
}

// ---- backward -----------------------------------------------------------------------------------
// pass 1 (plain stores, full coverage): grad_q[n,i,:] = 2*g[n,i]*(q[n,i]-t[n,idx[n,i]]) for i < len_q
// pass 2 (RED.ADD):                     grad_t[n,idx[n,i],:] -= the same
__global__ void chamfer_bwd_own_kernel(const float *__restrict__ q, const float *__restrict__ t, int Pq, int Pt, int D,
                                       const int64_t *__restrict__ qlen, const int64_t *__restrict__ tlen,
                                       const int64_t *__restrict__ idx,
                                       const float *__restrict__ g, int64_t total, float *__restrict__ gq)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(e % D);
        const int64_t row = e / D;  // n*Pq + i
        const int64_t n = row / Pq, i = row - n * Pq;
        float v = 0.f;
        if (g && idx) {
            const long long l = qlen ? qlen[n] : Pq, lt = tlen ? tlen[n] : Pt;
            if (i < l && lt > 0) {  // pytorch3d's backward: p1_idx < lengths1 && k < lengths2
                const int64_t j = idx[row];
                v = 2.0f * g[row] * (q[e] - t[(n * Pt + j) * D + d]);
            }
        }
        gq[e] = v;
    }
}

__global__ void chamfer_bwd_scatter_kernel(const float *__restrict__ q, const float *__restrict__ t, int Pq, int Pt,
                                           int D, const int64_t *__restrict__ qlen,
                                           const int64_t *__restrict__ tlen, const int64_t *__restrict__ idx,
                                           const float *__restrict__ g, int64_t total, float *__restrict__ gt)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(e % D);
        const int64_t row = e / D;
        const int64_t n = row / Pq, i = row - n * Pq;
        const long long l = qlen ? qlen[n] : Pq, lt = tlen ? tlen[n] : Pt;
        if (i < l && lt > 0) {
            const int64_t j = idx[row];
            const float v = 2.0f * g[row] * (q[e] - t[(n * Pt + j) * D + d]);
            atomicAdd(gt + (n * Pt + j) * D + d, -v);
        }
    }
}

// General-K backward (both parts with atomics into zero-filled buffers).
__global__ void knn_bwd_kernel(const float *__restrict__ p1, const float *__restrict__ p2, int P1, int P2, int D, int K,
                               const int64_t *__restrict__ len1, const int64_t *__restrict__ len2,
                               const int64_t *__restrict__ idx, const float *__restrict__ g, int64_t total,
                               float *__restrict__ g1, float *__restrict__ g2)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(e % D);
        const int64_t rk = e / D;  // (n*P1 + i)*K + k
        const int64_t row = rk / K;
        const int64_t n = row / P1, i = row - n * P1;
        const long long l = len1 ? len1[n] : P1, l2 = len2 ? len2[n] : P2;
        if (i < l && (rk - row * K) < l2) {
            const int64_t j = idx[rk];
            const float v = 2.0f * g[rk] * (p1[row * D + d] - p2[(n * P2 + j) * D + d]);
            atomicAdd(g1 + row * D + d, v);
            atomicAdd(g2 + (n * P2 + j) * D + d, -v);
        }
    }
}

__global__ void padded_lengths_kernel(const float *__restrict__ y, int P2, int D, float sentinel,
                                      int64_t *__restrict__ first, int32_t *__restrict__ any_flag)
{
    const int n = blockIdx.x;
    int best = P2;
    for (int j = threadIdx.x; j < P2; j += blockDim.x)
        if (y[((int64_t)n * P2 + j) * D] == sentinel) {
            best = j;  // ascending scan per thread: the first hit is this thread's minimum
            break;
        }
    best = redux_min_s32(best);
    __shared__ int w[32];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < (blockDim.x >> 5) ? w[threadIdx.x] : P2;
        v = redux_min_s32(v);
        if (threadIdx.x == 0) {
            first[n] = v;
            if (v < P2) atomicOr(any_flag, 1);
        }
    }
}

static inline unsigned ew_grid(int64_t total, int threads)
{
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

template <int D, int RBIG>
static int launch_nn(const float *x, const float *y, int N, int P1, int P2, const int64_t *xl, const int64_t *yl,
                     float *dx, int64_t *ix, float *dy, int64_t *iy, cudaStream_t st)
{
    const int dir0 = dx ? 0 : 1, ndir = (dx ? 1 : 0) + (dy ? 1 : 0);
    const int pmax = (dx && dy) ? (P1 > P2 ? P1 : P2) : (dx ? P1 : P2);
    // use the register-blocked variant only when it still fills the machine (>= 2 waves)
    const int64_t big_blocks = (int64_t)N * ndir * ((pmax + kChThreads * RBIG - 1) / (kChThreads * RBIG));
    if (RBIG > 1 && big_blocks >= 2 * (int64_t)sm_count()) {
        dim3 grid((pmax + kChThreads * RBIG - 1) / (kChThreads * RBIG), N, ndir);
        chamfer_nn_kernel<D, RBIG><<<grid, kChThreads, 0, st>>>(x, y, P1, P2, xl, yl, dx, ix, dy, iy, dir0);
    } else {
        dim3 grid((pmax + kChThreads - 1) / kChThreads, N, ndir);
        chamfer_nn_kernel<D, 1><<<grid, kChThreads, 0, st>>>(x, y, P1, P2, xl, yl, dx, ix, dy, iy, dir0);
    }
    return check_launch("chamfer_nn_kernel");
}

static int launch_generic(const float *p1, const float *p2, int N, int P1, int P2, int D, const int64_t *l1,
                          const int64_t *l2, int K, float *dists, int64_t *idx, cudaStream_t st)
{
    const int tile_rows = 4096 / D < 32 ? 32 : 4096 / D;
    const size_t smem = ((size_t)kChThreads * (D + 1) + (size_t)tile_rows * D) * sizeof(float);
    dim3 grid((P1 + kChThreads - 1) / kChThreads, N);
    auto kern = knn_generic_kernel<8>;
    if (smem > 48 * 1024) MPB_ENSURE_DYN_SMEM(kern, smem);
    kern<<<grid, kChThreads, smem, st>>>(p1, p2, P1, P2, D, l1, l2, K, dists, idx, tile_rows);
    return check_launch("knn_generic_kernel");
}

}  // namespace mpb

extern "C" int mpb_chamfer_nn_f32(const float *x, const float *y, int N, int P1, int P2, int D, const int64_t *x_len,
                                  const int64_t *y_len, float *dist_x, int64_t *idx_x, float *dist_y, int64_t *idx_y,
                                  void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 1, "bad size");
    MPB_REQUIRE(D <= 64, "D > 64 unsupported");
    MPB_REQUIRE((dist_x == nullptr) == (idx_x == nullptr) && (dist_y == nullptr) == (idx_y == nullptr),
                "dist/idx outputs must be given in pairs");
    if (N == 0 || (!dist_x && !dist_y)) return MPB_OK;
    MPB_REQUIRE(N <= 65535, "N exceeds grid.y");
    MPB_REQUIRE(x && y, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (P1 == 0 || P2 == 0) {  // nothing to match against: all-zero outputs, like zero-initialised knn results
        if (dist_x && P1) {
            MPB_CUDA(cudaMemsetAsync(dist_x, 0, sizeof(float) * (size_t)N * P1, st));
            MPB_CUDA(cudaMemsetAsync(idx_x, 0, sizeof(int64_t) * (size_t)N * P1, st));
        }
        if (dist_y && P2) {
            MPB_CUDA(cudaMemsetAsync(dist_y, 0, sizeof(float) * (size_t)N * P2, st));
            MPB_CUDA(cudaMemsetAsync(idx_y, 0, sizeof(int64_t) * (size_t)N * P2, st));
        }
        return MPB_OK;
    }
    switch (D) {
        case 3: return launch_nn<3, 4>(x, y, N, P1, P2, x_len, y_len, dist_x, idx_x, dist_y, idx_y, st);
        case 6: return launch_nn<6, 4>(x, y, N, P1, P2, x_len, y_len, dist_x, idx_x, dist_y, idx_y, st);
        case 24: return launch_nn<24, 2>(x, y, N, P1, P2, x_len, y_len, dist_x, idx_x, dist_y, idx_y, st);
        default: break;
    }
    if (dist_x) {
        int rc = launch_generic(x, y, N, P1, P2, D, x_len, y_len, 1, dist_x, idx_x, st);
        if (rc) return rc;
    }
    if (dist_y) return launch_generic(y, x, N, P2, P1, D, y_len, x_len, 1, dist_y, idx_y, st);
    return MPB_OK;
}

extern "C" int mpb_chamfer_nn_bwd_f32(const float *x, const float *y, int N, int P1, int P2, int D,
                                      const int64_t *x_len, const int64_t *y_len, const int64_t *idx_x,
                                      const int64_t *idx_y, const float *gdx, const float *gdy, float *grad_x,
                                      float *grad_y, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 1, "bad size");
    MPB_REQUIRE(x && y && grad_x && grad_y, "null pointer");
    MPB_REQUIRE((gdx == nullptr) == (idx_x == nullptr) && (gdy == nullptr) == (idx_y == nullptr),
                "grad/idx inputs must be given in pairs");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t tx = (int64_t)N * P1 * D, ty = (int64_t)N * P2 * D;
    if (tx) chamfer_bwd_own_kernel<<<ew_grid(tx, 256), 256, 0, st>>>(x, y, P1, P2, D, x_len, y_len, idx_x, gdx, tx, grad_x);
    if (ty) chamfer_bwd_own_kernel<<<ew_grid(ty, 256), 256, 0, st>>>(y, x, P2, P1, D, y_len, x_len, idx_y, gdy, ty, grad_y);
    if (tx && ty && gdx) chamfer_bwd_scatter_kernel<<<ew_grid(tx, 256), 256, 0, st>>>(x, y, P1, P2, D, x_len, y_len, idx_x, gdx, tx, grad_y);
    if (tx && ty && gdy) chamfer_bwd_scatter_kernel<<<ew_grid(ty, 256), 256, 0, st>>>(y, x, P2, P1, D, y_len, x_len, idx_y, gdy, ty, grad_x);
    return check_launch("chamfer_bwd kernels");
}

extern "C" int mpb_knn_points_f32(const float *p1, const float *p2, int N, int P1, int P2, int D, const int64_t *len1,
                                  const int64_t *len2, int K, float *dists, int64_t *idx, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 1 && K >= 1, "bad size");
    MPB_REQUIRE(K <= 8, "K > 8 unsupported");
    MPB_REQUIRE(D <= 64, "D > 64 unsupported");
    if (N == 0 || P1 == 0) return MPB_OK;
    MPB_REQUIRE(N <= 65535, "N exceeds grid.y");
    MPB_REQUIRE(p1 && p2 && dists && idx, "null pointer");
    return launch_generic(p1, p2, N, P1, P2, D, len1, len2, K, dists, idx, (cudaStream_t)stream);
}

extern "C" int mpb_knn_points_bwd_f32(const float *p1, const float *p2, int N, int P1, int P2, int D,
                                      const int64_t *len1, const int64_t *len2, const int64_t *idx, int K,
                                      const float *grad_dists, float *grad_p1, float *grad_p2, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(N >= 0 && P1 >= 0 && P2 >= 0 && D >= 1 && K >= 1, "bad size");
    MPB_REQUIRE(p1 && p2 && idx && grad_dists && grad_p1 && grad_p2, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if ((int64_t)N * P1 * D) MPB_CUDA(cudaMemsetAsync(grad_p1, 0, sizeof(float) * (size_t)N * P1 * D, st));
    if ((int64_t)N * P2 * D) MPB_CUDA(cudaMemsetAsync(grad_p2, 0, sizeof(float) * (size_t)N * P2 * D, st));
    const int64_t total = (int64_t)N * P1 * K * D;
    if (total == 0 || P2 == 0) return MPB_OK;
    knn_bwd_kernel<<<ew_grid(total, 256), 256, 0, st>>>(p1, p2, P1, P2, D, K, len1, len2, idx, grad_dists, total, grad_p1, grad_p2);
    return check_launch("knn_bwd_kernel");
}

extern "C" int mpb_padded_lengths_f32(const float *y, int N, int P2, int D, float sentinel, int64_t *first,
                                      int32_t *any_flag, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(N >= 0 && P2 >= 0 && D >= 1, "bad size");
    if (N == 0) return MPB_OK;
    MPB_REQUIRE(y && first && any_flag, "null pointer");
    padded_lengths_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(y, P2, D, sentinel, first, any_flag);
    return check_launch("padded_lengths_kernel");
}
