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

// Work is counted in ITEMS: 16-byte vectors for a segment whose rows are a whole number of vectors (or a plain copy of a
// multiple of 4 words) with 16-byte aligned ends, single words otherwise.  Per-sample sizes fit 32 bits (checked on the host),
// so the row / word split of a padded segment is two 32-bit divisions; a plain copy needs none.
struct StageTable {
    const uint32_t *src[kStageMaxSegs];
    uint32_t *dst[kStageMaxSegs];
    int64_t end[kStageMaxSegs];       // exclusive prefix of destination ITEM counts
    uint32_t src_rows[kStageMaxSegs], dst_rows[kStageMaxSegs], row_items[kStageMaxSegs];   // row length in items
    uint32_t pad[kStageMaxSegs];
    uint8_t vec[kStageMaxSegs], plain[kStageMaxSegs];
    int nsegs;
};

__global__ void __launch_bounds__(256) stage_batch_kernel(const __grid_constant__ StageTable t)
{
    const int64_t total = t.end[t.nsegs - 1];
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int s = 0;
        while (e >= t.end[s]) ++s;
        const int64_t o = e - (s ? t.end[s - 1] : 0);
        int64_t from = o;            // source item; -1: padding
        if (!t.plain[s]) {
            const uint32_t per_b = t.dst_rows[s] * t.row_items[s];
            const int64_t b = o < ((int64_t)1 << 32) ? (int64_t)((uint32_t)o / per_b) : o / per_b;
            const uint32_t rem = (uint32_t)(o - b * per_b);
            const uint32_t r = rem / t.row_items[s], w = rem - r * t.row_items[s];
            from = r < t.src_rows[s] ? (b * t.src_rows[s] + r) * t.row_items[s] + w : -1;
        }
        if (t.vec[s]) {
            const uint32_t p = t.pad[s];
            reinterpret_cast<uint4 *>(t.dst[s])[o] = from >= 0 ? reinterpret_cast<const uint4 *>(t.src[s])[from] : make_uint4(p, p, p, p);
        } else {
            t.dst[s][o] = from >= 0 ? t.src[s][from] : t.pad[s];
        }
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
        MPB_REQUIRE((src[s] || src_rows[s] == 0) && dst[s] && batch[s] > 0 && dst_rows[s] > 0 && row_words[s] > 0 && src_rows[s] >= 0 && src_rows[s] <= dst_rows[s],
                    "bad segment (src_rows <= dst_rows, positive sizes)");
        const bool plain = src_rows[s] == dst_rows[s];
        const int64_t words = batch[s] * dst_rows[s] * row_words[s];
        MPB_REQUIRE(plain || dst_rows[s] * row_words[s] < ((int64_t)1 << 32), "a padded segment's per-sample size must fit 32 bits");
        const bool aligned = (((uintptr_t)src[s] | (uintptr_t)dst[s]) & 15) == 0;
        const bool vec = aligned && (plain ? words % 4 == 0 : row_words[s] % 4 == 0);
        t.src[s] = (const uint32_t *)src[s], t.dst[s] = (uint32_t *)dst[s];
        t.plain[s] = plain, t.vec[s] = vec, t.pad[s] = pad_bits[s];
        t.src_rows[s] = (uint32_t)(plain ? 0 : src_rows[s]), t.dst_rows[s] = (uint32_t)(plain ? 0 : dst_rows[s]);
        t.row_items[s] = (uint32_t)(plain ? 0 : row_words[s] / (vec ? 4 : 1));
        total += words / (vec ? 4 : 1);
        t.end[s] = total;
    }
    for (int s = nsegs; s < kStageMaxSegs; ++s) {
        t.end[s] = total, t.src[s] = nullptr, t.dst[s] = nullptr;
        t.src_rows[s] = t.dst_rows[s] = t.row_items[s] = t.pad[s] = 0, t.vec[s] = t.plain[s] = 0;
    }
    t.nsegs = nsegs;
    int64_t blocks = (total + 256 * 2 - 1) / (256 * 2);
    const int64_t cap = (int64_t)sm_count() * 8;
    blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
    stage_batch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(t);
    return check_launch("stage_batch_kernel");
}
