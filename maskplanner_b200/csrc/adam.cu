// Multi-tensor Adam step for sm_100a: ONE launch updates every parameter of the model.
// Reference: train_maskplanner.py:159, :221 (torch.optim.Adam(model.parameters(), lr) + opt.step()).
//
// The step is pure HBM streaming (read p, g, m, v; write p, m, v: 28 bytes per parameter), so the only design
// questions are launch count and access width: pointer tables travel BY VALUE in the kernel parameters (no
// device-side metadata to keep in sync with autograd's changing .grad tensors, and the addresses are baked into a
// captured CUDA graph node), every CTA owns one fixed-size chunk of one tensor, accesses are 16-byte vectors when
// the four arrays of a tensor are 16-byte aligned.  The step counter lives on the device (graph replays advance
// it): every CTA reads it first, and the LAST CTA to finish (ticket counter) stores the incremented value.
//
// Arithmetic (torch/optim/adam.py, _single_tensor_adam, amsgrad = False, maximize = False):
//   g += weight_decay * p;  m += (g - m) * (1 - beta1);  v = beta2 * v + (1 - beta2) * g * g
//   p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
#include "common.cuh"

namespace mpb {

constexpr int kAdamMaxTensors = 80;   // keeps the by-value table below the classic 4 KB kernel-parameter limit
constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = kAdamThreads * 16;  // elements per CTA: four float4 per thread per array

struct AdamTable {
    float *p[kAdamMaxTensors];
    const float *g[kAdamMaxTensors];
    float *m[kAdamMaxTensors];
    float *v[kAdamMaxTensors];
    int64_t numel[kAdamMaxTensors];
    int chunk_end[kAdamMaxTensors];  // exclusive prefix of chunk counts: tensor t owns chunks [chunk_end[t-1], chunk_end[t])
    int ntensors;
};

struct AdamHyper {   // betas arrive as doubles: torch forms 1 - beta and beta^t in double before rounding to fp32
    double beta1_d, beta2_d;
    float lr, beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay, grad_scale;
};

__device__ __forceinline__ void adam_update(float &p, float g, float &m, float &v, const AdamHyper &h, float step_size, float inv_bc2_sqrt)
{
    g *= h.grad_scale;   // 1 on one GPU; 1/world_size when the gradient buffer holds the SUM over ranks (exact for powers of two)
    if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
    m = fmaf(g - m, h.one_minus_beta1, m);
    v = fmaf(h.beta2, v, h.one_minus_beta2 * g * g);
    const float denom = fmaf(sqrtf(v), inv_bc2_sqrt, h.eps);
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(kAdamThreads, 4)
adam_kernel(const __grid_constant__ AdamTable tab, AdamHyper h, const float *__restrict__ lr_dev, float *__restrict__ step,
            unsigned *__restrict__ ticket)
{
    const float t = *step + 1.f;
    if (lr_dev) h.lr = *lr_dev;
    const double bc1 = 1.0 - pow(h.beta1_d, (double)t), bc2 = 1.0 - pow(h.beta2_d, (double)t);
    const float step_size = (float)((double)h.lr / bc1), inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));

    int ti = 0;  // tensor owning this chunk: the table is short, a linear scan by one thread is cheap
    __shared__ int s_ti;
    if (threadIdx.x == 0) {
        while (ti < tab.ntensors - 1 && (int)blockIdx.x >= tab.chunk_end[ti]) ++ti;
        s_ti = ti;
    }
    __syncthreads();
    ti = s_ti;
    const int first = ti == 0 ? 0 : tab.chunk_end[ti - 1];
    const int64_t n = tab.numel[ti], e0 = (int64_t)((int)blockIdx.x - first) * kAdamChunk;
    float *p = tab.p[ti], *m = tab.m[ti], *v = tab.v[ti];
    const float *g = tab.g[ti];
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (vec && e0 + kAdamChunk <= n) {
        // two passes of two float4 per array: 8 loads in flight per thread at ~56 registers (4 CTAs/SM); four float4 per
        // array at once needed 90 registers and ran at 2 CTAs/SM (ncu: 22 % warps active, 3.96 TB/s)
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            float4 P[2], G[2], M[2], V[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int64_t e = e0 + (int64_t)((2 * half + u) * kAdamThreads + threadIdx.x) * 4;
                P[u] = *reinterpret_cast<const float4 *>(p + e);
                G[u] = *reinterpret_cast<const float4 *>(g + e);
                M[u] = *reinterpret_cast<const float4 *>(m + e);
                V[u] = *reinterpret_cast<const float4 *>(v + e);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                adam_update(P[u].x, G[u].x, M[u].x, V[u].x, h, step_size, inv_bc2_sqrt);
                adam_update(P[u].y, G[u].y, M[u].y, V[u].y, h, step_size, inv_bc2_sqrt);
                adam_update(P[u].z, G[u].z, M[u].z, V[u].z, h, step_size, inv_bc2_sqrt);
                adam_update(P[u].w, G[u].w, M[u].w, V[u].w, h, step_size, inv_bc2_sqrt);
                const int64_t e = e0 + (int64_t)((2 * half + u) * kAdamThreads + threadIdx.x) * 4;
                *reinterpret_cast<float4 *>(p + e) = P[u];
                *reinterpret_cast<float4 *>(m + e) = M[u];
                *reinterpret_cast<float4 *>(v + e) = V[u];
            }
        }
    } else {
        const int64_t e1 = e0 + kAdamChunk < n ? e0 + kAdamChunk : n;
        for (int64_t e = e0 + threadIdx.x; e < e1; e += kAdamThreads) {
            float pp = p[e], mm = m[e], vv = v[e];
            adam_update(pp, g[e], mm, vv, h, step_size, inv_bc2_sqrt);
            p[e] = pp, m[e] = mm, v[e] = vv;
        }
    }
    // the last CTA to get here publishes the new step count (everyone has read the old one by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *step = t;
            *ticket = 0u;
        }
    }
}

}  // namespace mpb

extern "C" int mpb_adam_step_f32(int ntensors, float *const *params, const float *const *grads, float *const *exp_avg,
                                 float *const *exp_avg_sq, const int64_t *numel, float lr, const float *lr_dev, double beta1,
                                 double beta2, double eps, double weight_decay, float grad_scale, float *step, uint32_t *ticket,
                                 void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(ntensors >= 0, "negative tensor count");
    if (ntensors == 0) return MPB_OK;
    MPB_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && step && ticket, "null pointer");
    MPB_REQUIRE(ntensors <= kAdamMaxTensors, "more than 80 tensors per call: split the parameter list");
    AdamTable tab;
    int chunks = 0, nt = 0;
    for (int i = 0; i < ntensors; ++i) {
        MPB_REQUIRE(numel[i] >= 0, "negative numel");
        if (numel[i] == 0) continue;
        MPB_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "null tensor pointer");
        const int64_t c = (numel[i] + kAdamChunk - 1) / kAdamChunk;
        MPB_REQUIRE(chunks + c < (int64_t)1 << 30, "too many elements");
        tab.p[nt] = params[i], tab.g[nt] = grads[i], tab.m[nt] = exp_avg[i], tab.v[nt] = exp_avg_sq[i];
        tab.numel[nt] = numel[i];
        chunks += (int)c;
        tab.chunk_end[nt] = chunks;
        ++nt;
    }
    if (nt == 0) return MPB_OK;
    tab.ntensors = nt;
    AdamHyper h;
    h.beta1_d = beta1, h.beta2_d = beta2;
    h.lr = lr, h.beta1 = (float)beta1, h.beta2 = (float)beta2;
    h.one_minus_beta1 = (float)(1.0 - beta1), h.one_minus_beta2 = (float)(1.0 - beta2);
    h.eps = (float)eps, h.weight_decay = (float)weight_decay, h.grad_scale = grad_scale;
    adam_kernel<<<chunks, kAdamThreads, 0, (cudaStream_t)stream>>>(tab, h, lr_dev, step, ticket);
    return check_launch("adam_kernel");
}
