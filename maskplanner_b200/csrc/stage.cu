// One-launch staging of a training batch into the fixed buffers a captured CUDA graph reads.
// Reference: the per-step input handling of the loop, train_maskplanner.py:207-208 (permute + .cuda() of the batch) together with
// the collate function's padding (utils/dataset/paintnet_ODv1.py:738-747: ground-truth rows padded with -100, stroke ids with -1).
// A replayed graph reads its inputs from fixed addresses, so every step copies the batch there; with stock torch ops that is
// three fill + copy pairs (padding to the configured maxima) and half a dozen device-to-device copies -- ~17 launches of
// 2-8 us in front of every replay.  Here all of it is ONE launch: up to 8 segments of 32-bit words, each
//     dst[b, r, :] = r < src_rows ? src[b, r, :] : pad          b < batch, r < dst_rows
// (int64 tensors travel as pairs of words; a segment with src_rows == dst_rows is a plain copy).
#include "common.cuh"

namespace mpb {

constexpr int kStageMaxSegs = 8;

struct StageTable {
    const uint32_t *src[kStageMaxSegs];
    uint32_t *dst[kStageMaxSegs];
    int64_t src_rows[kStageMaxSegs], dst_rows[kStageMaxSegs], row_words[kStageMaxSegs];
    int64_t end[kStageMaxSegs];       // exclusive prefix of destination word counts
    uint32_t pad[kStageMaxSegs];
    int nsegs;
};

__global__ void __launch_bounds__(256) stage_batch_kernel(const __grid_constant__ StageTable t)
{
    const int64_t total = t.end[t.nsegs - 1];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int s = 0;
        while (e >= t.end[s]) ++s;
        const int64_t o = e - (s ? t.end[s - 1] : 0);
        const int64_t per_b = t.dst_rows[s] * t.row_words[s];
        const int64_t b = o / per_b, rem = o - b * per_b;
        const int64_t r = rem / t.row_words[s], w = rem - r * t.row_words[s];
        t.dst[s][o] = r < t.src_rows[s] ? t.src[s][(b * t.src_rows[s] + r) * t.row_words[s] + w] : t.pad[s];
    }
}

}  // namespace mpb

extern "C" int mpb_stage_batch(int nsegs, const void *const *src, void *const *dst, const int64_t *batch, const int64_t *src_rows,
                               const int64_t *dst_rows, const int64_t *row_words, const uint32_t *pad_bits, void *stream)
{
    using namespace mpb;
    MPB_REQUIRE(nsegs >= 1 && nsegs <= kStageMaxSegs, "1 <= nsegs <= 8");
    MPB_REQUIRE(src && dst && batch && src_rows && dst_rows && row_words && pad_bits, "null pointer");
    StageTable t;
    int64_t total = 0;
    for (int s = 0; s < nsegs; ++s) {
        MPB_REQUIRE(src[s] && dst[s] && batch[s] > 0 && dst_rows[s] > 0 && row_words[s] > 0 && src_rows[s] >= 0 && src_rows[s] <= dst_rows[s],
                    "bad segment (src_rows <= dst_rows, positive sizes)");
        t.src[s] = (const uint32_t *)src[s], t.dst[s] = (uint32_t *)dst[s];
        t.src_rows[s] = src_rows[s], t.dst_rows[s] = dst_rows[s], t.row_words[s] = row_words[s], t.pad[s] = pad_bits[s];
        total += batch[s] * dst_rows[s] * row_words[s];
        t.end[s] = total;
    }
    for (int s = nsegs; s < kStageMaxSegs; ++s) t.end[s] = total;
    t.nsegs = nsegs;
    int64_t blocks = (total + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)sm_count() * 8;
    blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
    stage_batch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(t);
    return check_launch("stage_batch_kernel");
}
