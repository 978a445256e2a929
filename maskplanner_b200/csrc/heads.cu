// Regression heads of the MaskPlanner regressor -- the small kernels around the head GEMMs.
// Reference: models/pointnet2_cls_ssg.py:270-295 (modules), :309-341 (forward):
//     x = drop(relu(bn1(fc1(feat))));  h = drop(relu(bn2(fc2(x))));  seg = fc3(h);  nrm = normalize(tanh(fc_normals(h))) * w
//     m = drop(relu(sm_bn1(sm_fc1(feat)))); m = drop(relu(sm_bn2(sm_fc2(m)))); masks = sm_fc3(m); scores = mask_conf_out(m)
//
// With M = batch size (<= 64 rows) these layers are pure weight streaming: >90 % of the model's parameters, each read once
// per pass.  The GEMMs run on the shared-MLP tcgen05 kernels (sa_gemm.cu) with the WEIGHT as the 128-row operand and the
// batch as the N = 32..64 columns, so every activation of the heads lives FEATURE-MAJOR, [features, Bp] with Bp = the batch
// padded to a multiple of 32:
//     forward    Yt [Nout, Bp] = W [Nout, Kin] * X [Bp, Kin]^T          mpb_sa_gemm_tn   (A = W straight from HBM, fp32 read as TF32)
//     dW         dW [Nout, Kin] = dYt [Nout, Bp] * Xt [Kin, Bp]^T       mpb_sa_gemm_tn   (K = Bp; written straight into .grad)
//     dX         dXt [Kin, Bp] = W^T * dYt                               mpb_sa_gemm_wgrad (both operands MN-major, split over Nout)
// In that layout BatchNorm1d's batch statistics, ReLU and dropout are ROW-LOCAL (one warp owns one feature and its Bp batch
// entries), which is what the kernels below do: bias + BatchNorm1d (training: batch statistics, running-stat update; eval:
// running statistics) + ReLU + dropout forward and backward, the two layout changes at the ends of the chain, and the pose
// output (tanh + per-3-vector normalisation + interleave, :332-339).  Dropout masks come from a counter-based hash keyed by
// (seed, step counter, layer, feature, sample): nothing is stored, the backward pass regenerates the mask, and a CUDA-graph
// replay advances the device-side step counter itself.
#include "common.cuh"

namespace mpb {

constexpr int kHeadWarps = 32;          // features per CTA: one warp each
constexpr int kHeadMaxBp = 128;         // batch columns (padded) a warp holds: up to 4 per lane

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float tf32_hi(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// Counter-based uniform in [0, 1): a 64-bit mix (splitmix64 finaliser) of (seed, step, layer, feature, sample).
__device__ __forceinline__ float head_uniform(uint64_t seed, uint64_t step, uint32_t layer, uint32_t f, uint32_t b)
{
    uint64_t x = seed ^ (step * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)layer << 56) ^ ((uint64_t)f << 24) ^ (uint64_t)b;
    x ^= x >> 30, x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27, x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (float)(uint32_t)(x >> 40) * (1.0f / 16777216.0f);
}

__global__ void rng_advance_kernel(int64_t *counter) { *counter += 1; }

struct HeadActArgs {
    int F, B, Bp;
    const float *bias, *gamma, *beta;
    float *running_mean, *running_var;
    float momentum, eps, drop_p;
    int training, layer;
    uint64_t seed;
    const int64_t *step;
};

// Yt [F, Bp] (the GEMM's output, bias not yet added) -> the activation in both layouts: Xt [F, Bp] (feature-major, operand
// of the dW GEMM and of the next act_bwd) and X [Bp, F] (batch-major, B operand of the next forward GEMM); with SPLIT also
// their low parts for the 3xTF32 GEMMs (hi = tf32(a) in X / Xt, lo = a - hi in X_lo / Xt_lo).  mean / rstd [F] are saved for
// the backward pass.  Columns b >= B (batch padding) are written as zeros.
template <bool SPLIT>
__global__ void __launch_bounds__(32 * kHeadWarps)
head_act_fwd_kernel(const float *__restrict__ Yt, HeadActArgs a, float *__restrict__ Xt, float *__restrict__ Xt_lo, float *__restrict__ X,
                    float *__restrict__ X_lo, float *__restrict__ mean_out, float *__restrict__ rstd_out)
{
    __shared__ float tile[kHeadWarps][kHeadMaxBp + 1];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f0 = blockIdx.x * kHeadWarps, f = f0 + w;
    const int nb = a.Bp >> 5;
    float v[kHeadMaxBp / 32];
    if (f < a.F) {
        const float bias = a.bias ? a.bias[f] : 0.f;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kHeadMaxBp / 32; ++i) {
            const int b = lane + 32 * i;
            v[i] = (i < nb && b < a.B) ? Yt[(size_t)f * a.Bp + b] + bias : 0.f;
            s += v[i];
        }
        float mean, rstd;
        if (a.training) {     // batch statistics (biased variance for the normalisation, unbiased for the running estimate)
            mean = warp_sum(s) / (float)a.B;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < kHeadMaxBp / 32; ++i) {
                const int b = lane + 32 * i;
                const float d = (i < nb && b < a.B) ? v[i] - mean : 0.f;
                q = fmaf(d, d, q);
            }
            const float var = warp_sum(q) / (float)a.B;
            rstd = rsqrtf(var + a.eps);
            if (lane == 0) {
                if (a.running_mean) a.running_mean[f] = (1.f - a.momentum) * a.running_mean[f] + a.momentum * mean;
                if (a.running_var) a.running_var[f] = (1.f - a.momentum) * a.running_var[f] + a.momentum * (a.B > 1 ? var * (float)a.B / (float)(a.B - 1) : var);
            }
        } else {
            mean = a.running_mean[f];
            rstd = rsqrtf(a.running_var[f] + a.eps);
        }
        if (lane == 0) mean_out[f] = mean, rstd_out[f] = rstd;
        const float g = a.gamma ? a.gamma[f] : 1.f, be = a.beta ? a.beta[f] : 0.f;
        const float keep_scale = (a.training && a.drop_p > 0.f) ? 1.f / (1.f - a.drop_p) : 1.f;
        const uint64_t step = a.step ? (uint64_t)*a.step : 0ull;
#pragma unroll
        for (int i = 0; i < kHeadMaxBp / 32; ++i) {
            const int b = lane + 32 * i;
            if (i < nb) {
                float y = 0.f;
                if (b < a.B) {
                    y = fmaxf(fmaf((v[i] - mean) * rstd, g, be), 0.f);
                    if (a.training && a.drop_p > 0.f) y = head_uniform(a.seed, step, (uint32_t)a.layer, (uint32_t)f, (uint32_t)b) >= a.drop_p ? y * keep_scale : 0.f;
                }
                const float hi = SPLIT ? tf32_hi(y) : y;
                Xt[(size_t)f * a.Bp + b] = hi;
                if (SPLIT) Xt_lo[(size_t)f * a.Bp + b] = y - hi;
                tile[w][b] = y;
            }
        }
    } else {
        for (int b = lane; b < a.Bp; b += 32) tile[w][b] = 0.f;
    }
    __syncthreads();
    // batch-major copy: row b gets the CTA's 32 features as one 128-byte segment
    for (int b = w; b < a.Bp; b += kHeadWarps) {
        const int ff = f0 + lane;
        if (ff < a.F) {
            const float y = tile[lane][b];
            const float hi = SPLIT ? tf32_hi(y) : y;
            X[(size_t)b * a.F + ff] = hi;
            if (SPLIT) X_lo[(size_t)b * a.F + ff] = y - hi;
        }
    }
}

// dXt [F, Bp] (gradient w.r.t. the post-dropout activation, feature-major) -> dYt [F, Bp] (gradient w.r.t. the GEMM output),
// dgamma, dbeta, dbias [F].  Regenerates the dropout mask, re-forms the ReLU mask from Yt, then BatchNorm1d backward:
// training: dY = gamma*rstd*(dy - mean(dy) - yhat*mean(dy*yhat)); eval: dY = gamma*rstd*dy.
__global__ void __launch_bounds__(32 * kHeadWarps)
head_act_bwd_kernel(const float *__restrict__ dXt, const float *__restrict__ Yt, HeadActArgs a, const float *__restrict__ mean_in,
                    const float *__restrict__ rstd_in, float *__restrict__ dYt, float *__restrict__ dgamma, float *__restrict__ dbeta,
                    float *__restrict__ dbias)
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kHeadWarps + w;
    if (f >= a.F) return;
    const int nb = a.Bp >> 5;
    const float bias = a.bias ? a.bias[f] : 0.f, mean = mean_in[f], rstd = rstd_in[f];
    const float g = a.gamma ? a.gamma[f] : 1.f, be = a.beta ? a.beta[f] : 0.f;
    const float keep_scale = (a.training && a.drop_p > 0.f) ? 1.f / (1.f - a.drop_p) : 1.f;
    const uint64_t step = a.step ? (uint64_t)*a.step : 0ull;
    float dy[kHeadMaxBp / 32], yh[kHeadMaxBp / 32];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadMaxBp / 32; ++i) {
        const int b = lane + 32 * i;
        dy[i] = 0.f, yh[i] = 0.f;
        if (i < nb && b < a.B) {
            yh[i] = (Yt[(size_t)f * a.Bp + b] + bias - mean) * rstd;
            float d = dXt[(size_t)f * a.Bp + b];
            if (a.training && a.drop_p > 0.f) d = head_uniform(a.seed, step, (uint32_t)a.layer, (uint32_t)f, (uint32_t)b) >= a.drop_p ? d * keep_scale : 0.f;
            dy[i] = fmaf(yh[i], g, be) > 0.f ? d : 0.f;
            s0 += dy[i];
            s1 = fmaf(dy[i], yh[i], s1);
        }
    }
    s0 = warp_sum(s0), s1 = warp_sum(s1);
    const float m0 = a.training ? s0 / (float)a.B : 0.f, m1 = a.training ? s1 / (float)a.B : 0.f;
    float sb = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadMaxBp / 32; ++i) {
        const int b = lane + 32 * i;
        if (i < nb) {
            const float d = b < a.B ? g * rstd * (dy[i] - m0 - yh[i] * m1) : 0.f;
            dYt[(size_t)f * a.Bp + b] = d;
            sb += d;
        }
    }
    sb = warp_sum(sb);
    if (lane == 0) {
        if (dgamma) dgamma[f] = s1;
        if (dbeta) dbeta[f] = s0;
        if (dbias) dbias[f] = sb;
    }
}

// Layout changes at the ends of the chain, 32 x 32 tiles through shared memory.
//   to_feature_major : X [B, F]  -> Xt [F, Bp]  (+ optional split; padding columns zero; optional dbias[f] = sum_b X[b, f])
//   to_batch_major   : Yt [F, Bp] (+ bias[f]) -> Y [B, F]
template <bool SPLIT>
__global__ void __launch_bounds__(1024)
to_feature_major_kernel(const float *__restrict__ X, int B, int Bp, int F, float *__restrict__ Xt, float *__restrict__ Xt_lo,
                        float *__restrict__ colsum)
{
    __shared__ float tile[32][33];
    const int f0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float cs = 0.f;
    for (int b0 = 0; b0 < Bp; b0 += 32) {
        const int b = b0 + ty, f = f0 + tx;
        tile[ty][tx] = (b < B && f < F) ? X[(size_t)b * F + f] : 0.f;
        __syncthreads();
        const int fo = f0 + ty, bo = b0 + tx;
        if (fo < F) {
            const float v = tile[tx][ty];
            const float hi = SPLIT ? tf32_hi(v) : v;
            Xt[(size_t)fo * Bp + bo] = hi;
            if (SPLIT) Xt_lo[(size_t)fo * Bp + bo] = v - hi;
            cs += v;
        }
        __syncthreads();
    }
    if (colsum) {
        cs = warp_sum(cs);
        if (tx == 0 && f0 + ty < F) colsum[f0 + ty] = cs;
    }
}

__global__ void __launch_bounds__(1024)
to_batch_major_kernel(const float *__restrict__ Yt, const float *__restrict__ bias, int B, int Bp, int F, float *__restrict__ Y)
{
    __shared__ float tile[32][33];
    const int f0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int b0 = 0; b0 < B; b0 += 32) {
        const int f = f0 + ty, b = b0 + tx;
        tile[ty][tx] = (f < F && b < Bp) ? Yt[(size_t)f * Bp + b] + (bias ? bias[f] : 0.f) : 0.f;
        __syncthreads();
        const int bo = b0 + ty, fo = f0 + tx;
        if (bo < B && fo < F) Y[(size_t)bo * F + fo] = tile[tx][ty];
        __syncthreads();
    }
}

// Pose output (:332-339): pose j of sample b = [ fc3[3j..3j+2] + b3 , w * normalize(tanh(fc_normals[3j..3j+2] + bn)) ].
// One warp per pose j: lanes walk the batch; Yt3 / Ytn rows 3j..3j+2 are read coalesced, out [B, P, 6] is written as
// 24-byte records (2.7 MB in total).
__global__ void __launch_bounds__(256)
pose_out_fwd_kernel(const float *__restrict__ Yt3, const float *__restrict__ b3, const float *__restrict__ Ytn, const float *__restrict__ bn,
                    int B, int Bp, int P, float weight_orient, float *__restrict__ out)
{
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= P) return;
    float bs[3], bo[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) bs[c] = b3[3 * j + c], bo[c] = bn[3 * j + c];
    for (int b = lane; b < B; b += 32) {
        float s[3], t[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s[c] = Yt3[(size_t)(3 * j + c) * Bp + b] + bs[c];
            t[c] = tanhf(Ytn[(size_t)(3 * j + c) * Bp + b] + bo[c]);
        }
        const float nrm = fmaxf(sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]), 1e-12f);      // F.normalize eps
        float *o = out + ((size_t)b * P + j) * 6;
        o[0] = s[0], o[1] = s[1], o[2] = s[2];
        o[3] = t[0] / nrm * weight_orient, o[4] = t[1] / nrm * weight_orient, o[5] = t[2] / nrm * weight_orient;
    }
}

// d_out [B, P, 6] -> dYt3, dYtn [3P, Bp] (padding columns zero) and the two bias gradients (row sums).
__global__ void __launch_bounds__(256)
pose_out_bwd_kernel(const float *__restrict__ d_out, const float *__restrict__ Ytn, const float *__restrict__ bn, int B, int Bp, int P,
                    float weight_orient, float *__restrict__ dYt3, float *__restrict__ dYtn, float *__restrict__ db3, float *__restrict__ dbn)
{
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= P) return;
    float bo[3], sum_s[3] = {0.f, 0.f, 0.f}, sum_n[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 3; ++c) bo[c] = bn[3 * j + c];
    for (int b = lane; b < Bp; b += 32) {
        float ds[3] = {0.f, 0.f, 0.f}, dn[3] = {0.f, 0.f, 0.f};
        if (b < B) {
            const float *g = d_out + ((size_t)b * P + j) * 6;
            float t[3], go[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                ds[c] = g[c];
                go[c] = g[3 + c] * weight_orient;
                t[c] = tanhf(Ytn[(size_t)(3 * j + c) * Bp + b] + bo[c]);
            }
            // u = t / max(|t|, eps): du/dt = (I - u u^T) / |t|  (the clamp is never active for tanh outputs of real data)
            const float nrm = fmaxf(sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]), 1e-12f);
            const float u0 = t[0] / nrm, u1 = t[1] / nrm, u2 = t[2] / nrm;
            const float dot = go[0] * u0 + go[1] * u1 + go[2] * u2;
            const float dt[3] = {(go[0] - dot * u0) / nrm, (go[1] - dot * u1) / nrm, (go[2] - dot * u2) / nrm};
#pragma unroll
            for (int c = 0; c < 3; ++c) dn[c] = dt[c] * (1.f - t[c] * t[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            dYt3[(size_t)(3 * j + c) * Bp + b] = ds[c];
            dYtn[(size_t)(3 * j + c) * Bp + b] = dn[c];
            sum_s[c] += ds[c], sum_n[c] += dn[c];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = warp_sum(sum_s[c]), bsum = warp_sum(sum_n[c]);
        if (lane == 0) db3[3 * j + c] = a, dbn[3 * j + c] = bsum;
    }
}

}  // namespace mpb

extern "C" int mpb_rng_advance(int64_t *counter, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(counter, "null pointer");
    rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    return check_launch("rng_advance_kernel");
}

#define MPB_HEAD_CHECKS()                                                                            \
    MPB_REQUIRE(F > 0 && B > 0 && Bp >= B && Bp % 32 == 0 && Bp <= mpb::kHeadMaxBp, "bad size (Bp: batch padded to a multiple of 32, at most 128)")

extern "C" int mpb_head_act_fwd(const float *Yt, int F, int B, int Bp, const float *bias, const float *gamma, const float *beta,
                                float *running_mean, float *running_var, float momentum, float eps, int training, float drop_p,
                                uint64_t seed, const int64_t *step, int layer, float *Xt, float *Xt_lo, float *X, float *X_lo,
                                float *mean, float *rstd, void *stream)
{
    using namespace mpb;
    MPB_HEAD_CHECKS();
    MPB_REQUIRE(Yt && Xt && X && mean && rstd, "null pointer");
    MPB_REQUIRE(training || (running_mean && running_var), "eval mode needs the running statistics");
    MPB_REQUIRE((Xt_lo != nullptr) == (X_lo != nullptr), "Xt_lo / X_lo must come together");
    MPB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout probability out of range");
    HeadActArgs a;
    a.F = F, a.B = B, a.Bp = Bp, a.bias = bias, a.gamma = gamma, a.beta = beta, a.running_mean = running_mean, a.running_var = running_var;
    a.momentum = momentum, a.eps = eps, a.drop_p = drop_p, a.training = training, a.layer = layer, a.seed = seed, a.step = step;
    const int grid = (F + kHeadWarps - 1) / kHeadWarps;
    if (Xt_lo)
        head_act_fwd_kernel<true><<<grid, 32 * kHeadWarps, 0, (cudaStream_t)stream>>>(Yt, a, Xt, Xt_lo, X, X_lo, mean, rstd);
    else
        head_act_fwd_kernel<false><<<grid, 32 * kHeadWarps, 0, (cudaStream_t)stream>>>(Yt, a, Xt, nullptr, X, nullptr, mean, rstd);
    return check_launch("head_act_fwd_kernel");
}

extern "C" int mpb_head_act_bwd(const float *dXt, const float *Yt, int F, int B, int Bp, const float *bias, const float *gamma,
                                const float *beta, const float *mean, const float *rstd, int training, float drop_p, uint64_t seed,
                                const int64_t *step, int layer, float *dYt, float *dgamma, float *dbeta, float *dbias, void *stream)
{
    using namespace mpb;
    MPB_HEAD_CHECKS();
    MPB_REQUIRE(dXt && Yt && mean && rstd && dYt, "null pointer");
    HeadActArgs a;
    a.F = F, a.B = B, a.Bp = Bp, a.bias = bias, a.gamma = gamma, a.beta = beta, a.running_mean = nullptr, a.running_var = nullptr;
    a.momentum = 0.f, a.eps = 0.f, a.drop_p = drop_p, a.training = training, a.layer = layer, a.seed = seed, a.step = step;
    head_act_bwd_kernel<<<(F + kHeadWarps - 1) / kHeadWarps, 32 * kHeadWarps, 0, (cudaStream_t)stream>>>(dXt, Yt, a, mean, rstd, dYt, dgamma,
                                                                                                      dbeta, dbias);
    return check_launch("head_act_bwd_kernel");
}

extern "C" int mpb_head_to_feature_major(const float *X, int B, int Bp, int F, float *Xt, float *Xt_lo, float *colsum, void *stream)
{
    using namespace mpb;
    MPB_HEAD_CHECKS();
    MPB_REQUIRE(X && Xt, "null pointer");
    const int grid = (F + 31) / 32;
    if (Xt_lo)
        to_feature_major_kernel<true><<<grid, 1024, 0, (cudaStream_t)stream>>>(X, B, Bp, F, Xt, Xt_lo, colsum);
    else
        to_feature_major_kernel<false><<<grid, 1024, 0, (cudaStream_t)stream>>>(X, B, Bp, F, Xt, nullptr, colsum);
    return check_launch("to_feature_major_kernel");
}

extern "C" int mpb_head_to_batch_major(const float *Yt, const float *bias, int B, int Bp, int F, float *Y, void *stream)
{
    using namespace mpb;
    MPB_HEAD_CHECKS();
    MPB_REQUIRE(Yt && Y, "null pointer");
    to_batch_major_kernel<<<(F + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(Yt, bias, B, Bp, F, Y);
    return check_launch("to_batch_major_kernel");
}

extern "C" int mpb_head_pose_out_fwd(const float *Yt3, const float *b3, const float *Ytn, const float *bn, int B, int Bp, int P,
                                     float weight_orient, float *out, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(Yt3 && b3 && Ytn && bn && out && B > 0 && Bp >= B && P > 0, "bad argument");
    pose_out_fwd_kernel<<<(P + 7) / 8, 256, 0, (cudaStream_t)stream>>>(Yt3, b3, Ytn, bn, B, Bp, P, weight_orient, out);
    return check_launch("pose_out_fwd_kernel");
}

extern "C" int mpb_head_pose_out_bwd(const float *d_out, const float *Ytn, const float *bn, int B, int Bp, int P, float weight_orient,
                                     float *dYt3, float *dYtn, float *db3, float *dbn, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(d_out && Ytn && bn && dYt3 && dYtn && db3 && dbn && B > 0 && Bp >= B && P > 0, "bad argument");
    pose_out_bwd_kernel<<<(P + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_out, Ytn, bn, B, Bp, P, weight_orient, dYt3, dYtn, db3, dbn);
    return check_launch("pose_out_bwd_kernel");
}
