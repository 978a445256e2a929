// The MaskPlanner training loss around the chamfer nearest-neighbour kernels, as a handful of fused launches.
// Reference: loss_handler.py:596-666 (get_asymm_v6_chamfer_with_stroke_masks) and :816-935 (get_stroke_masks_loss)
// with the resolved weights of config=[maskplanner,<cat>,longx_v2].  The reference evaluates the loss with ~150 small
// torch ops, B device->host copies and a scipy call per sample; round 1 of this library kept the torch ops (about 130
// launches of 2-10 us between the chamfer kernels and the Hungarian solver: 0.55 ms of a 3.6 ms step in the CUPTI
// timeline, profiles/r02_graph_timeline.txt).  Here the same arithmetic is six launches forward and three backward:
//
//   loss_lengths_kernel        `padded=True` length scan of both ground-truth tensors          pytorch3d_chamfer.py:138-149
//   (mpb_chamfer_nn_f32 x2)    nearest neighbours: segments (both directions), poses (GT -> prediction)
//   mask_cost_kernel           ids of the matched GT segments (:838) and ALL B x P x T BCE cost matrices (:860-873)
//   (mpb_lap_f32)              per-sample Hungarian matching                                    :875
//   loss_value_kernel          the five loss terms and their weighted sum                        :604-645, :906, :930
//   loss_bwd_masks_kernel      d loss / d mask logits, d loss / d mask scores
//   loss_bwd_pred_own_kernel   d loss / d prediction, prediction -> GT direction (term 1)
//   loss_bwd_pred_scatter_kernel   GT -> prediction directions of terms 3 and 2 (scatter-add)
//
// BCE-with-logits against a binary target y is softplus(x) - x*y, so cost[b,p,t] = sum_i softplus(x[b,p,i]) - sum_{i in
// stroke t} x[b,p,i]; the matched BCE of the loss (:886-906) is exactly the selected cost entry, nothing is recomputed.
#include "common.cuh"

namespace mpb {

constexpr int kLossMaxMasks = 32;

__device__ __forceinline__ float softplus_stable(float x) { return fmaxf(x, 0.f) + log1pf(__expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// first[n] = index of the first row whose channel 0 equals the sentinel (or P): blocks [0, B) scan `a`, [B, 2B) scan `b`.
__global__ void __launch_bounds__(256) loss_lengths_kernel(const float *__restrict__ a, int Pa, int Da, const float *__restrict__ b, int Pb,
                                                           int Db, int B, float sentinel, int64_t *__restrict__ len_a,
                                                           int64_t *__restrict__ len_b)
{
    const bool second = (int)blockIdx.x >= B;
    const int n = second ? blockIdx.x - B : blockIdx.x;
    const float *y = second ? b : a;
    const int P = second ? Pb : Pa, D = second ? Db : Da;
    int best = P;
    for (int j = threadIdx.x; j < P; j += blockDim.x)
        if (y[((int64_t)n * P + j) * D] == sentinel) {
            best = j;
            break;
        }
    best = redux_min_s32(best);
    __shared__ int w[8];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < 8 ? w[threadIdx.x] : P;
        v = redux_min_s32(v);
        if (threadIdx.x == 0) (second ? len_b : len_a)[n] = v;
    }
}

// One CTA per (predicted mask p, sample b): ids[i] = stroke id of the GT segment matched to predicted segment i (:838),
// cost[b,p,t] = sum_i BCEWithLogits(x[b,p,i], [ids[i] == t]) for every t, present[b,t].  Fixed-order reductions only
// (the assignment is discrete: run-to-run noise in the costs could flip it).
__global__ void __launch_bounds__(128) mask_cost_kernel(const float *__restrict__ masks, const float *__restrict__ stroke_ids,
                                                        const int64_t *__restrict__ match, int NM, int P1, int P2,
                                                        float *__restrict__ cost, uint8_t *__restrict__ present,
                                                        int32_t *__restrict__ ids_out)
{
    extern __shared__ float sm[];
    float *sx = sm;
    int *sid = reinterpret_cast<int *>(sm + P1);
    __shared__ float red[4];
    const int p = blockIdx.x, b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *x = masks + ((size_t)b * NM + p) * P1;
    float sp = 0.f;
    for (int i = threadIdx.x; i < P1; i += 128) {
        const float v = x[i];
        const int id = (int)stroke_ids[(size_t)b * P2 + match[(size_t)b * P1 + i]];
        sx[i] = v;
        sid[i] = id;
        sp += softplus_stable(v);
        if (p == 0) ids_out[(size_t)b * P1 + i] = id;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sp += __shfl_xor_sync(0xffffffffu, sp, off);
    if (lane == 0) red[warp] = sp;
    __syncthreads();
    sp = (red[0] + red[1]) + (red[2] + red[3]);
    for (int t = warp; t < NM; t += 4) {
        float s = 0.f;
        int cnt = 0;
        for (int i = lane; i < P1; i += 32) {
            const bool in = sid[i] == t;
            s += in ? sx[i] : 0.f;
            cnt += in ? 1 : 0;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, off);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        }
        if (lane == 0) {
            cost[((size_t)b * NM + p) * NM + t] = sp - s;
            if (p == 0) present[(size_t)b * NM + t] = cnt > 0 ? 1 : 0;
        }
    }
}

__device__ __forceinline__ double block_sum_1024(double v, double *sh)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];   // same order in every thread and on every run
    __syncthreads();
    return t;
}

struct LossValueArgs {
    const float *d_x, *d_y, *d_y2;          // [B,P1], [B,P2], [B,P3] nearest-neighbour distances
    const int64_t *len_y, *len_y2;          // [B]
    const float *cost;                      // [B,NM,NM]
    const uint8_t *present;                 // [B,NM]
    const int64_t *row;                     // [B,NM] predicted mask matched to target t (-1: absent)
    const float *scores;                    // [B,NM]
    const float *weights;                   // [5] segment chamfer, reverse point chamfer, reverse segment chamfer, masks, confidence
    const float *n_pairs_in;                // optional: normaliser of the matched-pair mean (global count under data parallel)
    float no_stroke_w;
    int B, P1, P2, P3, NM;
    float *loss, *terms;                    // terms[8]: t1, t2, t3, weighted mask term, sum of matched BCE, local pair count, confidence loss, pair normaliser used
};

__global__ void __launch_bounds__(1024) loss_value_kernel(const LossValueArgs a)
{
    __shared__ double sh[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int B = a.B, NM = a.NM;
    double s = 0.0;
    for (int i = tid; i < B * a.P1; i += 1024) s += (double)a.d_x[i];
    const double t1 = 100.0 * block_sum_1024(s, sh) / ((double)B * a.P1);                       // :604-611
    double acc3 = 0.0, acc2 = 0.0;
    for (int b = warp; b < B; b += 32) {
        const int l3 = (int)a.len_y[b], l2 = (int)a.len_y2[b];
        double v3 = 0.0, v2 = 0.0;
        for (int j = lane; j < l3; j += 32) v3 += (double)a.d_y[(size_t)b * a.P2 + j];
        for (int j = lane; j < l2; j += 32) v2 += (double)a.d_y2[(size_t)b * a.P3 + j];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            v3 += __shfl_xor_sync(0xffffffffu, v3, off);
            v2 += __shfl_xor_sync(0xffffffffu, v2, off);
        }
        acc3 += v3 / (double)l3;                                                                 // point_reduction="mean" over the real rows
        acc2 += v2 / (double)l2;
    }
    const double t3 = 100.0 * block_sum_1024(lane == 0 ? acc3 : 0.0, sh) / B;                   // :642-645
    const double t2 = 100.0 * block_sum_1024(lane == 0 ? acc2 : 0.0, sh) / B;                   // :631-636
    double sm = 0.0, np = 0.0, cf = 0.0;
    for (int i = tid; i < B * NM; i += 1024) {
        const int b = i / NM, t = i - b * NM;
        if (a.present[i]) {
            const int p = (int)a.row[i];
            np += 1.0;
            if (p >= 0) sm += (double)a.cost[((size_t)b * NM + p) * NM + t];
        }
        // confidence target of predicted mask t' = i % NM: matched to any present stroke? (:920-921)
        const int pm = t;
        float y = 0.f;
        for (int tt = 0; tt < NM; ++tt)
            if (a.present[b * NM + tt] && (int)a.row[b * NM + tt] == pm) y = 1.f;
        const float sc = a.scores[i];
        const float w = a.no_stroke_w + (1.f - a.no_stroke_w) * y;                               // :924-925
        cf += (double)(w * (fmaxf(sc, 0.f) - sc * y + log1pf(__expf(-fabsf(sc)))));
    }
    const double S = block_sum_1024(sm, sh), n_local = block_sum_1024(np, sh);
    const double conf = block_sum_1024(cf, sh) / ((double)B * NM);                               // :930
    if (tid == 0) {
        const double n = a.n_pairs_in ? (double)a.n_pairs_in[0] : n_local;
        const double masks = (double)a.weights[3] * (S / n) + (double)a.weights[4] * conf;      // :906, :932-935
        a.terms[0] = (float)t1, a.terms[1] = (float)t2, a.terms[2] = (float)t3, a.terms[3] = (float)masks;
        a.terms[4] = (float)S, a.terms[5] = (float)n_local, a.terms[6] = (float)conf, a.terms[7] = (float)n;
        a.loss[0] = (float)((double)a.weights[0] * t1 + (double)a.weights[1] * t2 + (double)a.weights[2] * t3 + masks);
    }
}

// d loss / d mask logits [B,NM,P1] and d loss / d scores [B,NM]; g = upstream gradient of the scalar loss.
__global__ void __launch_bounds__(128) loss_bwd_masks_kernel(const float *__restrict__ masks, const float *__restrict__ scores,
                                                             const int32_t *__restrict__ ids, const uint8_t *__restrict__ present,
                                                             const int64_t *__restrict__ row, const float *__restrict__ weights,
                                                             const float *__restrict__ n_pairs_in, const float *__restrict__ g, float no_stroke_w,
                                                             int B, int NM, int P1, float *__restrict__ gmasks, float *__restrict__ gscores)
{
    const int p = blockIdx.x, b = blockIdx.y;
    int tm = -1;
    for (int t = 0; t < NM; ++t)
        if (present[b * NM + t] && (int)row[b * NM + t] == p) tm = t;
    // matched-pair normaliser (:906): the given (global) count, or the number of present strokes of this batch -- recomputed
    // here from B * NM bytes so that the backward pass does not wait for the loss-value kernel
    __shared__ int s_np[4];
    int np = 0;
    if (!n_pairs_in)
        for (int i = threadIdx.x; i < B * NM; i += 128) np += present[i] ? 1 : 0;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) np += __shfl_xor_sync(0xffffffffu, np, off);
    if ((threadIdx.x & 31) == 0) s_np[threadIdx.x >> 5] = np;
    __syncthreads();
    const float n_pairs = n_pairs_in ? n_pairs_in[0] : (float)(s_np[0] + s_np[1] + s_np[2] + s_np[3]);
    const float go = g[0];
    const float coef = go * weights[3] / n_pairs;
    const float *x = masks + ((size_t)b * NM + p) * P1;
    float *gx = gmasks + ((size_t)b * NM + p) * P1;
    const int32_t *id = ids + (size_t)b * P1;
    for (int i = threadIdx.x; i < P1; i += 128)
        gx[i] = tm >= 0 ? coef * (sigmoidf_(x[i]) - (id[i] == tm ? 1.f : 0.f)) : 0.f;
    if (threadIdx.x == 0) {
        const float y = tm >= 0 ? 1.f : 0.f, s = scores[b * NM + p];
        const float w = no_stroke_w + (1.f - no_stroke_w) * y;
        gscores[b * NM + p] = go * weights[4] / ((float)B * NM) * w * (sigmoidf_(s) - y);
    }
}

// term 1 (prediction -> nearest GT segment, mean over B * P1): grad[n,i,:] = c1 * 2 * (x[n,i] - y[n, idx_x[n,i]]).
__global__ void loss_bwd_pred_own_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ idx_x,
                                         const int64_t *__restrict__ len_y, const float *__restrict__ weights,
                                         const float *__restrict__ g, int B, int P1, int P2, int D, int64_t total, float *__restrict__ gx)
{
    const float c1 = g[0] * weights[0] * 100.f / ((float)B * P1);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(e % D);
        const int64_t r = e / D, n = r / P1;
        float v = 0.f;
        if (len_y[n] > 0) v = 2.0f * c1 * (x[e] - y[(n * P2 + idx_x[r]) * D + d]);
        gx[e] = v;
    }
}

// terms 3 and 2 (every real GT row -> its nearest prediction, mean over the sample's rows, mean over B): scatter-add of
// -2 c (y[n,j] - x[n, idx[n,j]]) onto the matched prediction.  Items [0, t3) belong to the segment term, [t3, t3 + t2) to
// the pose term (the prediction viewed as [B, P1 * D / D2, D2]).
__global__ void loss_bwd_pred_scatter_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ idx_y,
                                             const int64_t *__restrict__ len_y, const float *__restrict__ y2,
                                             const int64_t *__restrict__ idx_y2, const int64_t *__restrict__ len_y2,
                                             const float *__restrict__ weights, const float *__restrict__ g, int B, int P1, int P2, int D,
                                             int P3, int D2, int64_t t3, int64_t t2, float *__restrict__ gx)
{
    const float go = g[0] * 100.f / (float)B;
    const int P1b = P1 * (D / D2);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < t3 + t2; e += (int64_t)gridDim.x * blockDim.x) {
        if (e < t3) {
            const int d = (int)(e % D);
            const int64_t r = e / D, n = r / P2, j = r - n * P2;
            const int64_t l = len_y[n];
            if (j < l) {
                const int64_t i = idx_y[r];
                const float c = go * weights[2] / (float)l;
                atomicAdd(gx + (n * P1 + i) * D + d, -2.0f * c * (y[e] - x[(n * P1 + i) * D + d]));
            }
        } else {
            const int64_t e2 = e - t3;
            const int d = (int)(e2 % D2);
            const int64_t r = e2 / D2, n = r / P3, j = r - n * P3;
            const int64_t l = len_y2[n];
            if (j < l) {
                const int64_t i = idx_y2[r];
                const float c = go * weights[1] / (float)l;
                atomicAdd(gx + (n * P1b + i) * D2 + d, -2.0f * c * (y2[e2] - x[(n * P1b + i) * D2 + d]));
            }
        }
    }
}

static inline unsigned loss_grid(int64_t total, int threads)
{
    int64_t blocks = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace mpb

extern "C" int mpb_loss_lengths_f32(const float *traj, int P2, int D, const float *traj_as_pc, int P3, int D2, int B, float sentinel,
                                    int64_t *len_traj, int64_t *len_pc, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && P2 >= 0 && P3 >= 0 && D >= 1 && D2 >= 1, "bad size");
    if (B == 0) return MPB_OK;
    MPB_REQUIRE(traj && traj_as_pc && len_traj && len_pc, "null pointer");
    loss_lengths_kernel<<<2 * B, 256, 0, (cudaStream_t)stream>>>(traj, P2, D, traj_as_pc, P3, D2, B, sentinel, len_traj, len_pc);
    return check_launch("loss_lengths_kernel");
}

extern "C" int mpb_mask_cost_f32(const float *masks, const float *stroke_ids, const int64_t *match, int B, int NM, int P1, int P2,
                                 float *cost, uint8_t *present, int32_t *ids_out, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && NM >= 1 && NM <= kLossMaxMasks && P1 >= 1 && P2 >= 1, "bad size (1 <= n_masks <= 32)");
    if (B == 0) return MPB_OK;
    MPB_REQUIRE(B <= 65535, "B exceeds grid.y");
    MPB_REQUIRE(masks && stroke_ids && match && cost && present && ids_out, "null pointer");
    const size_t smem = (size_t)P1 * 8;
    MPB_REQUIRE(smem <= 48 * 1024, "P1 too large for the shared-memory row (<= 6144 segments)");
    mask_cost_kernel<<<dim3(NM, B), 128, smem, (cudaStream_t)stream>>>(masks, stroke_ids, match, NM, P1, P2, cost, present, ids_out);
    return check_launch("mask_cost_kernel");
}

extern "C" int mpb_asymm_v6_loss_value_f32(const float *d_x, const float *d_y, const int64_t *len_y, const float *d_y2,
                                           const int64_t *len_y2, const float *cost, const uint8_t *present, const int64_t *row,
                                           const float *scores, const float *weights5, const float *n_pairs_in, float no_stroke_w,
                                           int B, int P1, int P2, int P3, int NM, float *loss, float *terms8, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 1 && P1 >= 1 && P2 >= 1 && P3 >= 1 && NM >= 1 && NM <= kLossMaxMasks, "bad size");
    MPB_REQUIRE(d_x && d_y && len_y && d_y2 && len_y2 && cost && present && row && scores && weights5 && loss && terms8, "null pointer");
    LossValueArgs a;
    a.d_x = d_x, a.d_y = d_y, a.d_y2 = d_y2, a.len_y = len_y, a.len_y2 = len_y2, a.cost = cost, a.present = present, a.row = row;
    a.scores = scores, a.weights = weights5, a.n_pairs_in = n_pairs_in, a.no_stroke_w = no_stroke_w;
    a.B = B, a.P1 = P1, a.P2 = P2, a.P3 = P3, a.NM = NM, a.loss = loss, a.terms = terms8;
    loss_value_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(a);
    return check_launch("loss_value_kernel");
}

extern "C" int mpb_asymm_v6_loss_bwd_f32(const float *y_pred, const float *traj, const float *traj_as_pc, const float *masks,
                                         const float *scores, const int64_t *idx_x, const int64_t *idx_y, const int64_t *len_y,
                                         const int64_t *idx_y2, const int64_t *len_y2, const int32_t *ids, const uint8_t *present,
                                         const int64_t *row, const float *weights5, const float *n_pairs_in, const float *grad_loss,
                                         float no_stroke_w, int B, int P1, int P2, int D, int P3, int D2, int NM, float *grad_pred,
                                         float *grad_masks, float *grad_scores, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 1 && P1 >= 1 && P2 >= 1 && P3 >= 1 && D >= 1 && D2 >= 1 && D % D2 == 0 && NM >= 1 && NM <= kLossMaxMasks, "bad size");
    MPB_REQUIRE(y_pred && traj && traj_as_pc && masks && scores && idx_x && idx_y && len_y && idx_y2 && len_y2 && ids && present && row &&
                    weights5 && grad_loss && grad_pred && grad_masks && grad_scores,
                "null pointer");
    MPB_REQUIRE(B <= 65535, "B exceeds grid.y");
    cudaStream_t st = (cudaStream_t)stream;
    loss_bwd_masks_kernel<<<dim3(NM, B), 128, 0, st>>>(masks, scores, ids, present, row, weights5, n_pairs_in, grad_loss, no_stroke_w, B, NM, P1,
                                                      grad_masks, grad_scores);
    const int64_t tx = (int64_t)B * P1 * D, t3 = (int64_t)B * P2 * D, t2 = (int64_t)B * P3 * D2;
    loss_bwd_pred_own_kernel<<<loss_grid(tx, 256), 256, 0, st>>>(y_pred, traj, idx_x, len_y, weights5, grad_loss, B, P1, P2, D, tx, grad_pred);
    loss_bwd_pred_scatter_kernel<<<loss_grid(t3 + t2, 256), 256, 0, st>>>(y_pred, traj, idx_y, len_y, traj_as_pc, idx_y2, len_y2, weights5,
                                                                        grad_loss, B, P1, P2, D, P3, D2, t3, t2, grad_pred);
    return check_launch("asymm_v6 loss backward kernels");
}
