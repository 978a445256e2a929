// Batched linear assignment on the device (the mask-loss host section).
// Reference: loss_handler.py:860-877 -- for every sample the [n_pred_masks, n_target_masks] BCE cost
// matrix is copied to the host and solved with scipy.optimize.linear_sum_assignment (B device->host
// synchronisations per step).  Here one WARP solves one sample: shortest-augmenting-path Hungarian
// algorithm with dual potentials, jobs (= predicted masks, <= 32) mapped to lanes, every inner loop a
// lane-parallel step and the path minimum a warp shuffle reduction.  fp64 arithmetic like scipy, so the
// (generically unique) optimum is the same assignment.
//
// cost [B, P, T] fp32: cost[b, p, t] of giving predicted mask p to target stroke t;
// present [B, T] uint8: target t exists in sample b (absent targets are skipped, like the reference's
// per-sample torch.unique); requires (number of present targets) <= P <= 32.
// out_row [B, T] int64: predicted mask assigned to target t, or -1 for absent targets.
#include "common.cuh"

namespace mpb {

__global__ void __launch_bounds__(32) lap_kernel(const float *__restrict__ cost, const uint8_t *__restrict__ present, int P, int T,
                                                 int64_t *__restrict__ out_row)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    const float *c = cost + (size_t)b * P * T;
    __shared__ double u[33];      // potential of worker i (1-based; worker = present target)
    __shared__ int job_of[33];    // p[j]: worker holding job j (1-based jobs; 0 = none); job_of[0] = worker being inserted
    __shared__ int way[33];
    __shared__ int worker_col[33];
    // the cost matrix, transposed (target-major, pitch 33: lane = job reads consecutive words).  Every step of the path search
    // reads one column of it; from global memory that is one dependent L2 round trip (~0.3 us) per step, ~250 steps per sample
    __shared__ float cs[32 * 33];
    for (int e = lane; e < P * T; e += 32) {
        const int p = e / T, t = e - p * T;
        cs[t * 33 + p] = c[e];
    }
    // compact the present targets into workers 1..n
    int n = 0;
    for (int t = 0; t < T; ++t)
        if (present[(size_t)b * T + t]) {
            ++n;
            if (lane == 0) worker_col[n] = t;
        }
    for (int i = lane; i <= 32; i += 32) u[i] = 0.0, job_of[i] = 0, way[i] = 0;
    if (lane == 0) u[32] = 0.0, job_of[32] = 0, way[32] = 0;
    __syncwarp();
    const int j = lane + 1;          // this lane's job (predicted mask lane), valid if j <= P
    const bool live = j <= P;
    double v = 0.0;                  // potential of job j
    for (int i = 1; i <= n && n <= P; ++i) {
        if (lane == 0) job_of[0] = i;
        __syncwarp();
        int j0 = 0;
        double minv = 1e300;
        bool used = false;
        while (true) {
            if (lane == j0 - 1) used = true;
            const int i0 = job_of[j0];
            const int col = worker_col[i0];
            const double ui0 = u[i0];
            double cand = 1e300;
            if (live && !used) {
                const double cur = (double)cs[col * 33 + (j - 1)] - ui0 - v;
                if (cur < minv) {
                    minv = cur;
                    way[j] = j0;
                }
                cand = minv;
            }
            // delta = min over unused jobs, j1 = lowest job index attaining it
            double delta = cand;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) delta = fmin(delta, __shfl_xor_sync(0xffffffffu, delta, o));
            const unsigned who = __ballot_sync(0xffffffffu, live && !used && cand == delta);
            const int j1 = __ffs(who);   // 1-based job index
            // update potentials
            if (used && live) {
                u[job_of[j]] += delta;   // distinct workers per used job
                v -= delta;
            } else if (live) {
                minv -= delta;
            }
            if (lane == 0) u[job_of[0]] += delta;   // the virtual job 0 is always "used"
            __syncwarp();
            j0 = j1;
            if (job_of[j0] == 0) break;
        }
        // augment along the path
        if (lane == 0) {
            int jj = j0;
            while (jj) {
                const int jn = way[jj];
                job_of[jj] = job_of[jn];
                jj = jn;
            }
        }
        __syncwarp();
    }
    for (int t = lane; t < T; t += 32) out_row[(size_t)b * T + t] = -1;
    __syncwarp();
    if (live && n <= P) {
        const int w = job_of[j];
        if (w > 0) out_row[(size_t)b * T + worker_col[w]] = j - 1;
    }
}

}  // namespace mpb

extern "C" int mpb_lap_f32(const float *cost, const uint8_t *present, int B, int P, int T, int64_t *out_row, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(B >= 0 && P >= 1 && T >= 1, "bad size");
    MPB_REQUIRE(P <= 32 && T <= 32, "at most 32 predicted masks / target strokes");
    if (B == 0) return MPB_OK;
    MPB_REQUIRE(cost && present && out_row, "null pointer");
    lap_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(cost, present, P, T, out_row);
    return check_launch("lap_kernel");
}
