// Farthest point sampling for sm_100a.                Reference: models/pointnet2_utils.py:65-86
//
// Design (B200-first, not a translation of the reference's 54-ATen-ops-per-sample loop):
//   * one persistent CTA per cloud -- or one thread-block CLUSTER per cloud when the cloud does
//     not fit one SM's registers (N > 8192): coordinates and the running min-distance of every
//     point stay in REGISTERS for the whole kernel (PPT points per thread), so the per-sample
//     HBM traffic is zero after the initial load;
//   * per sample: 8 individually rounded fp32 ops per point (the reference's arithmetic, see
//     below), a thread-local first-max, then a two-level REDUX.MAX / REDUX.MIN reduction over
//     (distance bits, index) with ONE __syncthreads (double-buffered warp slots);
//   * clusters exchange their CTA-local winner (value, index, x, y, z) through distributed shared
//     memory with one barrier.cluster per sample (double-buffered slots).
//
// Bit-exactness contract (checked against oracle/ and the reference's own torch code):
//   dist = ((dx*dx) + (dy*dy)) + (dz*dz), every product/sum rounded (torch.sum over 3 elements,
//   :82) -> __fmul_rn/__fadd_rn so ptxas cannot contract to FMA;  running distance starts at
//   1e10 and is replaced only when dist < it (:76, :83-84);  argmax = lowest index among equal
//   maxima (:85).  Distances are >= +0, so their bit patterns order like signed integers.
#include <stdlib.h>

#include "common.cuh"

namespace mpb {

constexpr int kFpsMaxPPT = 8;
constexpr int kFpsMaxThreads = 1024;

template <int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_resident_kernel(const float *__restrict__ xyz, int64_t sb, int64_t sn, int64_t sc, int N,
                    const int64_t *__restrict__ seed, int npoint, int64_t *__restrict__ out)
{
    extern __shared__ float smem_f[];
    const int T = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const unsigned CS = cluster_nctarank(), rank = cluster_ctarank();
    const int b = blockIdx.x / CS;
    const int slice = T * PPT;  // points held by this CTA
    float *sx = smem_f, *sy = smem_f + slice, *sz = smem_f + 2 * slice;
    __shared__ int2 wslot[2][32];
    __shared__ int4 cslot_a[2][16];  // (dist bits, index, x bits, y bits) from each cluster rank
    __shared__ int cslot_b[2][16];   // z bits

    const float *base = xyz + (int64_t)b * sb;
    float px[PPT], py[PPT], pz[PPT], md[PPT];
    const int stride_j = (int)CS * T;
    const int first = (int)rank * T + tid;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = j * stride_j + first;
        if (i < N) {
            const float *p = base + (int64_t)i * sn;
            px[j] = p[0];
            py[j] = p[sc];
            pz[j] = p[2 * sc];
            md[j] = 1e10f;
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            md[j] = -1.0f;  // padding lane: negative => can never win the arg-max
        }
        sx[j * T + tid] = px[j];
        sy[j * T + tid] = py[j];
        sz[j * T + tid] = pz[j];
    }
    long long s0 = seed[b];
    int cur = (int)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
    float cx, cy, cz;
    {
        const float *p = base + (int64_t)cur * sn;
        cx = p[0];
        cy = p[sc];
        cz = p[2 * sc];
    }
    if (CS > 1) cluster_arrive_release(), cluster_wait_acquire();  // peers' smem is live before any DSMEM store
    __syncthreads();

    int64_t *o = out + (int64_t)b * npoint;
    for (int s = 0; s < npoint; ++s) {
        if (rank == 0 && tid == 0) o[s] = cur;
        if (s == npoint - 1) break;
        const int buf = s & 1;
        float best = -2.0f;
        int bi = 0x7fffffff;
        // Packed fp32x2 update, two points per instruction: FADD2 differences (p + (-c) == p - c), FMUL2
        // products, scalar FADD sums (common.cuh: ptxas would contract packed sums of products), so every
        // lane carries the bits of the scalar form.  The odd point of an odd PPT goes scalar.
        const float ncx = -cx, ncy = -cy, ncz = -cz;
#pragma unroll
        for (int j = 0; j + 1 < PPT; j += 2) {
            const float2 dx = __fadd2_rn(make_float2(px[j], px[j + 1]), make_float2(ncx, ncx));
            const float2 dy = __fadd2_rn(make_float2(py[j], py[j + 1]), make_float2(ncy, ncy));
            const float2 dz = __fadd2_rn(make_float2(pz[j], pz[j + 1]), make_float2(ncz, ncz));
            const float2 d = add2_products_rn(add2_products_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
            md[j] = fminf(d.x, md[j]);          // == (d < md) ? d : md  (d is never -0, NaN keeps md)
            md[j + 1] = fminf(d.y, md[j + 1]);
        }
        if (PPT & 1) {
            constexpr int j = PPT - 1;
            const float dx = __fadd_rn(px[j], ncx), dy = __fadd_rn(py[j], ncy), dz = __fadd_rn(pz[j], ncz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            md[j] = fminf(d, md[j]);
        }
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            if (md[j] > best) {  // strict: the thread's points are visited in ascending index order
                best = md[j];
                bi = j * stride_j + first;
            }
        }
        // warp level: max distance bits, then lowest index among the lanes holding that max
        const int vb = __float_as_int(best);
        const int wmax = redux_max_s32(vb);
        const unsigned wi = redux_min_u32(vb == wmax ? (unsigned)bi : 0x7fffffffu);
        if (lane == 0) wslot[buf][warp] = make_int2(wmax, (int)wi);
        __syncthreads();
        int2 e = lane < nwarps ? wslot[buf][lane] : make_int2((int)0x80000000, 0x7fffffff);
        int gmax = redux_max_s32(e.x);
        unsigned gi = redux_min_u32(e.x == gmax ? (unsigned)e.y : 0x7fffffffu);
        if (CS == 1) {
            cur = (int)gi;
            cx = sx[cur];
            cy = sy[cur];
            cz = sz[cur];
        } else {
            // publish this CTA's winner to every rank (thread r writes to rank r), then reduce
            if (tid < (int)CS) {
                float wx = 0.f, wy = 0.f, wz = 0.f;
                if (gmax >= 0) {
                    const int li = ((int)gi / stride_j) * T + ((int)gi % T);
                    wx = sx[li];
                    wy = sy[li];
                    wz = sz[li];
                }
                st_shared_cluster_v4(map_shared_rank(&cslot_a[buf][rank], tid), gmax, (int)gi,
                                     __float_as_int(wx), __float_as_int(wy));
                st_shared_cluster_s32(map_shared_rank(&cslot_b[buf][rank], tid), __float_as_int(wz));
            }
            cluster_arrive_release();
            cluster_wait_acquire();
            int4 ca = lane < (int)CS ? cslot_a[buf][lane] : make_int4((int)0x80000000, 0x7fffffff, 0, 0);
            int cb = lane < (int)CS ? cslot_b[buf][lane] : 0;
            const int cmax = redux_max_s32(ca.x);
            const unsigned ci = redux_min_u32(ca.x == cmax ? (unsigned)ca.y : 0x7fffffffu);
            const unsigned who = __ballot_sync(0xffffffffu, ca.x == cmax && (unsigned)ca.y == ci);
            const int src = __ffs(who) - 1;
            cur = (int)ci;
            cx = __int_as_float(__shfl_sync(0xffffffffu, ca.z, src));
            cy = __int_as_float(__shfl_sync(0xffffffffu, ca.w, src));
            cz = __int_as_float(__shfl_sync(0xffffffffu, cb, src));
        }
    }
    if (CS > 1) cluster_arrive_release(), cluster_wait_acquire();  // nobody exits while peers may still store
}

// Fallback for clouds beyond the register/cluster capacity (N > 16 * 8192): running distances in a
// global workspace, coordinates re-read through L2 every sample.  Correct, not fast.
__global__ void __launch_bounds__(1024, 1)
fps_streaming_kernel(const float *__restrict__ xyz, int64_t sb, int64_t sn, int64_t sc, int N,
                     const int64_t *__restrict__ seed, int npoint, int64_t *__restrict__ out,
                     float *__restrict__ work)
{
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = blockDim.x, nwarps = T >> 5;
    __shared__ int2 wslot[2][32];
    const float *base = xyz + (int64_t)b * sb;
    float *md = work + (int64_t)b * N;
    for (int i = tid; i < N; i += T) md[i] = 1e10f;
    long long s0 = seed[b];
    int cur = (int)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
    int64_t *o = out + (int64_t)b * npoint;
    for (int s = 0; s < npoint; ++s) {
        if (tid == 0) o[s] = cur;
        if (s == npoint - 1) break;
        const int buf = s & 1;
        const float cx = base[(int64_t)cur * sn], cy = base[(int64_t)cur * sn + sc], cz = base[(int64_t)cur * sn + 2 * sc];
        float best = -2.0f;
        int bi = 0x7fffffff;
        for (int i = tid; i < N; i += T) {
            const float *p = base + (int64_t)i * sn;
            const float dx = __fsub_rn(p[0], cx), dy = __fsub_rn(p[sc], cy), dz = __fsub_rn(p[2 * sc], cz);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            float m = md[i];
            if (d < m) md[i] = m = d;
            if (m > best) best = m, bi = i;
        }
        const int vb = __float_as_int(best);
        const int wmax = redux_max_s32(vb);
        const unsigned wi = redux_min_u32(vb == wmax ? (unsigned)bi : 0x7fffffffu);
        if (lane == 0) wslot[buf][warp] = make_int2(wmax, (int)wi);
        __syncthreads();
        int2 e = lane < nwarps ? wslot[buf][lane] : make_int2((int)0x80000000, 0x7fffffff);
        const int gmax = redux_max_s32(e.x);
        cur = (int)redux_min_u32(e.x == gmax ? (unsigned)e.y : 0x7fffffffu);
    }
}

struct FpsPlan {
    int cluster;  // 0 => streaming fallback
    int threads;
    int ppt;
};

static FpsPlan fps_plan(int N)
{
    FpsPlan p;
    // 512-thread CTAs with up to 12 register-resident points per thread: same fp32 issue work as 1024 x 6 but half the
    // warps in the per-sample barrier / two-level arg-max (measured 0.265 vs 0.291 ms for B=64, 5120 -> 512).
    // MPB_FPS_THREADS=1024 restores the wide CTA, =512 forces the narrow one for any N <= 6144.
    static const int forced = [] {
        const char *e = getenv("MPB_FPS_THREADS");
        return e ? atoi(e) : 0;
    }();
    if (N <= 512 * 12 && (forced == 512 || (forced != 1024 && N > 4096))) {
        p.cluster = 1;
        p.threads = 512;
        p.ppt = (N + 511) / 512;
    } else if (N <= kFpsMaxPPT * kFpsMaxThreads) {
        p.cluster = 1;
        int t = ((N + 3) / 4 + 127) / 128 * 128;  // aim for ~4 points per thread
        p.threads = t < 128 ? 128 : (t > kFpsMaxThreads ? kFpsMaxThreads : t);
        p.ppt = (N + p.threads - 1) / p.threads;
    } else if (N <= 16 * kFpsMaxPPT * kFpsMaxThreads) {
        int cs = 2;
        while (cs * kFpsMaxPPT * kFpsMaxThreads < N) cs *= 2;
        p.cluster = cs;
        p.threads = kFpsMaxThreads;
        p.ppt = (N + cs * kFpsMaxThreads - 1) / (cs * kFpsMaxThreads);
    } else {
        p.cluster = 0;
        p.threads = 1024;
        p.ppt = 0;
    }
    return p;
}

template <int PPT, int MAXT>
static int launch_resident(const FpsPlan &p, const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                           const int64_t *seed, int npoint, int64_t *out, cudaStream_t st)
{
    auto kern = fps_resident_kernel<PPT, MAXT>;
    const size_t smem = (size_t)3 * p.threads * PPT * sizeof(float);
    MPB_ENSURE_DYN_SMEM(kern, smem);
    static bool nonportable_ok = false;
    if (p.cluster > 8 && !nonportable_ok) {
        MPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        nonportable_ok = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * p.cluster));
    cfg.blockDim = dim3((unsigned)p.threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)p.cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    MPB_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, sb, sn, sc, N, seed, npoint, out));
    return MPB_OK;
}

}  // namespace mpb

extern "C" int64_t mpb_fps_workspace_bytes(int B, int N)
{
    if (B <= 0 || N <= 0) return 0;
    return mpb::fps_plan(N).cluster == 0 ? (int64_t)B * N * (int64_t)sizeof(float) : 0;
}

extern "C" int mpb_fps_f32(const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N, const int64_t *seed_idx,
                           int npoint, int64_t *out_idx, void *workspace, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && N >= 0 && npoint >= 0, "negative size");
    if (B == 0 || npoint == 0) return MPB_OK;
    MPB_REQUIRE(N > 0, "cannot sample from an empty cloud");
    MPB_REQUIRE(xyz && seed_idx && out_idx, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const FpsPlan p = fps_plan(N);
    if (p.cluster == 0) {
        if (!workspace) {
            set_error("mpb_fps_f32: N=%d needs a workspace of %lld bytes", N, (long long)mpb_fps_workspace_bytes(B, N));
            return MPB_ERR_WORKSPACE;
        }
        fps_streaming_kernel<<<B, 1024, 0, st>>>(xyz, sb, sn, sc, N, seed_idx, npoint, out_idx, (float *)workspace);
        return check_launch("fps_streaming_kernel");
    }
    switch (p.ppt) {
#define MPB_CASE(P) \
    case P:         \
        return launch_resident<P, (P <= 8 ? 1024 : 512)>(p, xyz, sb, sn, sc, B, N, seed_idx, npoint, out_idx, st);
        MPB_CASE(1)
        MPB_CASE(2)
        MPB_CASE(3)
        MPB_CASE(4)
        MPB_CASE(5)
        MPB_CASE(6)
        MPB_CASE(7)
        MPB_CASE(8)
        MPB_CASE(9)
        MPB_CASE(10)
        MPB_CASE(11)
        MPB_CASE(12)
#undef MPB_CASE
    }
    set_error("mpb_fps_f32: internal plan error (ppt=%d)", p.ppt);
    return MPB_ERR_UNSUPPORTED;
}
