// Ball query, kNN grouping and square_distance for sm_100a.
// Reference: models/pointnet2_utils.py:21-42 (square_distance), :89-109 (query_ball_point).
//
// The reference materialises a [B,S,N] fp32 distance tensor plus a [B,S,N] int64 index tensor and
// fully sorts the latter; only the first `nsample` in-ball indices in ascending order survive.
// Here one WARP owns one query and walks the cloud in index order through a shared-memory tile
// of (x, y, z, |p|^2) float4s, 32 points per step, appending hits (ballot + popc keep them in index
// order) until `nsample` are found -- O(S*N) fp32 work, O(S*nsample) bytes written, nothing else
// touches HBM.  A block stops as soon as all of its queries are full.
//
// Bit-exactness contract: d = ((-2*dot) + |q|^2) + |p|^2, dot = fma(qz,pz, fma(qy,py, qx*px)),
// |v|^2 = ((vx*vx)+(vy*vy))+(vz*vz) with every op rounded; in-ball iff !(d > r2) (NaN counts as
// inside, like the reference's masked assignment at :104).
#include "common.cuh"

namespace mpb {

__device__ __forceinline__ float norm3_rn(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ float sqdist_expanded(float qx, float qy, float qz, float qn, const float4 p)
{
    const float dot = __fmaf_rn(qz, p.z, __fmaf_rn(qy, p.y, __fmul_rn(qx, p.x)));
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), qn), p.w);
}

constexpr int kBqTile = 1024;  // points per shared-memory tile (16 KB)

// One WARP per query: the 32 lanes test 32 consecutive points per step (conflict-free LDS.128), a ballot
// gives the in-ball mask and each hit lane stores its index at `cnt + popc(mask below me)`, which keeps
// the output in ascending index order.  B*S warps (32 768 at the micro-benchmark shape) fill the machine;
// a thread-per-query mapping has the same instruction count but only B*S/32 warps and is latency-bound.
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
ball_query_kernel(const float *__restrict__ xyz, int64_t xsb, int64_t xsn, int64_t xsc,
                  const float *__restrict__ q, int64_t qsb, int64_t qsn, int64_t qsc, int N, int S, float r2,
                  int nsample, int64_t *__restrict__ out)
{
    __shared__ float4 tile[kBqTile];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const bool live = s < S;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float *qp = q + (int64_t)b * qsb + (int64_t)s * qsn;
        qx = qp[0];
        qy = qp[qsc];
        qz = qp[2 * qsc];
    }
    const float qn = norm3_rn(qx, qy, qz);
    int64_t *o = out + ((int64_t)b * S + (live ? s : 0)) * nsample;
    int cnt = live ? 0 : nsample;   // warp-uniform
    int first = N;
    const unsigned below = (1u << lane) - 1u;
    const float *pb = xyz + (int64_t)b * xsb;
    for (int t0 = 0; t0 < N; t0 += kBqTile) {
        const int tn = min(kBqTile, N - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += WARPS * 32) {
            const float *p = pb + (int64_t)(t0 + i) * xsn;
            const float x = p[0], y = p[xsc], z = p[2 * xsc];
            tile[i] = make_float4(x, y, z, norm3_rn(x, y, z));
        }
        __syncthreads();
        for (int i0 = 0; i0 < tn && cnt < nsample; i0 += 64) {
            // two independent 32-point groups per iteration (two LDS + FMA chains in flight)
            const int ia = i0 + lane, ib = i0 + 32 + lane;
            const bool in_a = ia < tn && !(sqdist_expanded(qx, qy, qz, qn, tile[ia < tn ? ia : 0]) > r2);
            const bool in_b = ib < tn && !(sqdist_expanded(qx, qy, qz, qn, tile[ib < tn ? ib : 0]) > r2);
            const unsigned ma = __ballot_sync(0xffffffffu, in_a), mb = __ballot_sync(0xffffffffu, in_b);
            if (ma) {
                if (cnt == 0) first = t0 + i0 + __ffs(ma) - 1;
                const int pos = cnt + __popc(ma & below);
                if (in_a && pos < nsample) o[pos] = t0 + ia;
                cnt += __popc(ma);
            }
            if (mb && cnt < nsample) {
                if (cnt == 0) first = t0 + i0 + 32 + __ffs(mb) - 1;
                const int pos = cnt + __popc(mb & below);
                if (in_b && pos < nsample) o[pos] = t0 + ib;
                cnt += __popc(mb);
            }
        }
        if (__syncthreads_and(cnt >= nsample)) break;
    }
    if (live)
        for (int k = (cnt < nsample ? cnt : nsample) + lane; k < nsample; k += 32) o[k] = first;  // pad with the first hit (N if empty)
}

// kNN grouping: thread per query, the K smallest expanded-form distances kept sorted (ascending) in
// registers.  Insertion is a fully unrolled bubble-through pass (no dynamic register indexing):
// the candidate is carried down the list, takes the first slot it is strictly smaller than, and
// from there on every slot shifts by one -- equal distances therefore stay in index order and the
// element that falls off the end is the one with the highest index (lowest index wins ties).
template <int THREADS, int KMAX>
__global__ void __launch_bounds__(THREADS)
knn_group_kernel(const float *__restrict__ xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *__restrict__ q,
                 int64_t qsb, int64_t qsn, int64_t qsc, int N, int S, int K, int64_t *__restrict__ out,
                 float *__restrict__ out_d)
{
    __shared__ float4 tile[kBqTile];
    const int b = blockIdx.y;
    const int s = blockIdx.x * THREADS + threadIdx.x;
    const bool live = s < S;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
        const float *qp = q + (int64_t)b * qsb + (int64_t)s * qsn;
        qx = qp[0];
        qy = qp[qsc];
        qz = qp[2 * qsc];
    }
    const float qn = norm3_rn(qx, qy, qz);
    float bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) bd[k] = INFINITY, bi[k] = 0;
    float thr = INFINITY;  // current K-th smallest (INF until K finite candidates were seen)
    int have = 0;
    const float *pb = xyz + (int64_t)b * xsb;
    for (int t0 = 0; t0 < N; t0 += kBqTile) {
        const int tn = min(kBqTile, N - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += THREADS) {
            const float *p = pb + (int64_t)(t0 + i) * xsn;
            const float x = p[0], y = p[xsc], z = p[2 * xsc];
            tile[i] = make_float4(x, y, z, norm3_rn(x, y, z));
        }
        __syncthreads();
        if (live) {
            for (int i = 0; i < tn; ++i) {
                const float d = sqdist_expanded(qx, qy, qz, qn, tile[i]);
                if (d < thr) {
                    float cd = d;
                    int ci = t0 + i;
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
    if (live) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
            if (k < K) {
                const bool ok = k < have;
                out[((int64_t)b * S + s) * K + k] = ok ? bi[k] : 0;
                if (out_d) out_d[((int64_t)b * S + s) * K + k] = ok ? bd[k] : 0.f;
            }
    }
}

__global__ void square_distance_kernel(const float *__restrict__ src, const float *__restrict__ dst, int N, int M,
                                       float *__restrict__ out)
{
    const int b = blockIdx.z;
    const int i = blockIdx.y;
    const float *sp = src + ((int64_t)b * N + i) * 3;
    const float qx = sp[0], qy = sp[1], qz = sp[2];
    const float qn = norm3_rn(qx, qy, qz);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const float *dp = dst + ((int64_t)b * M + j) * 3;
        const float x = dp[0], y = dp[1], z = dp[2];
        out[((int64_t)b * N + i) * M + j] = sqdist_expanded(qx, qy, qz, qn, make_float4(x, y, z, norm3_rn(x, y, z)));
    }
}

}  // namespace mpb

extern "C" int mpb_ball_query_f32(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *new_xyz,
                                  int64_t qsb, int64_t qsn, int64_t qsc, int B, int N, int S, float r2, int nsample,
                                  int64_t *out_idx, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && S >= 0 && nsample >= 0, "negative size");
    if (B == 0 || S == 0 || nsample == 0) return MPB_OK;
    MPB_REQUIRE(xyz && new_xyz && out_idx, "null pointer");
    MPB_REQUIRE(B <= 65535, "B exceeds grid.y");
    cudaStream_t st = (cudaStream_t)stream;
    // 16 queries (warps) per block share one point tile; fall back to 4 when that would leave SMs idle
    if ((int64_t)B * ((S + 15) / 16) >= 2 * sm_count()) {
        dim3 grid((S + 15) / 16, B);
        ball_query_kernel<16><<<grid, 512, 0, st>>>(xyz, xsb, xsn, xsc, new_xyz, qsb, qsn, qsc, N, S, r2, nsample, out_idx);
    } else {
        dim3 grid((S + 3) / 4, B);
        ball_query_kernel<4><<<grid, 128, 0, st>>>(xyz, xsb, xsn, xsc, new_xyz, qsb, qsn, qsc, N, S, r2, nsample, out_idx);
    }
    return check_launch("ball_query_kernel");
}

extern "C" int mpb_knn_group_f32(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc, const float *new_xyz,
                                 int64_t qsb, int64_t qsn, int64_t qsc, int B, int N, int S, int k, int64_t *out_idx,
                                 float *out_dist, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && S >= 0 && k >= 0, "negative size");
    if (B == 0 || S == 0 || k == 0) return MPB_OK;
    MPB_REQUIRE(xyz && new_xyz && out_idx, "null pointer");
    MPB_REQUIRE(k <= 64, "k > 64 unsupported");
    MPB_REQUIRE(B <= 65535, "B exceeds grid.y");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((S + 63) / 64, B);
    if (k <= 16)
        knn_group_kernel<64, 16><<<grid, 64, 0, st>>>(xyz, xsb, xsn, xsc, new_xyz, qsb, qsn, qsc, N, S, k, out_idx, out_dist);
    else if (k <= 32)
        knn_group_kernel<64, 32><<<grid, 64, 0, st>>>(xyz, xsb, xsn, xsc, new_xyz, qsb, qsn, qsc, N, S, k, out_idx, out_dist);
    else
        knn_group_kernel<64, 64><<<grid, 64, 0, st>>>(xyz, xsb, xsn, xsc, new_xyz, qsb, qsn, qsc, N, S, k, out_idx, out_dist);
    return check_launch("knn_group_kernel");
}

extern "C" int mpb_square_distance_f32(const float *src, const float *dst, int B, int N, int M, float *out, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && M >= 0, "negative size");
    if (B == 0 || N == 0 || M == 0) return MPB_OK;
    MPB_REQUIRE(src && dst && out, "null pointer");
    MPB_REQUIRE(B <= 65535 && N <= 65535, "B or N exceeds grid limits");
    dim3 grid((unsigned)min((M + 255) / 256, 64), N, B);
    square_distance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, N, M, out);
    return check_launch("square_distance_kernel");
}
