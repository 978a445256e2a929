/*
 * oracle/c/mp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the index-producing arithmetic of MaskPlanner's point-cloud hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (maskplanner_b200/) never does.
 *
 * Every function cites the reference lines whose arithmetic it restates (paths relative to
 * /root/reference).  Build: `make -C oracle/c` (gcc -O2 -ffp-contract=off: every float product
 * and sum below is individually rounded unless fmaf() is written out).
 *
 * Parity pin: checked bit-for-bit against the reference's own torch code imported from
 * /root/reference (oracle/make_golden.py) and against the committed fixtures in tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Farthest point sampling.  models/pointnet2_utils.py:65-86
 *   :76  distance = ones(B,N) * 1e10
 *   :77  farthest = randint(0,N,(B,))          -> passed in as `seed` (host draws it)
 *   :80  centroids[:, i] = farthest
 *   :82  dist = sum((xyz - centroid) ** 2, -1) -> ((dx*dx)+(dy*dy))+(dz*dz), each op rounded
 *   :83-84 distance[dist < distance] = dist    -> strict '<' (NaN never replaces)
 *   :85  farthest = max(distance, -1)[1]       -> lowest index among equal maxima
 * xyz is addressed through explicit element strides so permuted views can be fed unchanged.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_fps_f32(const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                         const int64_t *seed, int npoint, int64_t *out /* [B,npoint] */)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (int64_t)b * sb;
        float *mind = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
        for (int i = 0; i < N; ++i) mind[i] = 1e10f;
        int64_t far = seed[b];
        for (int s = 0; s < npoint; ++s) {
            out[(int64_t)b * npoint + s] = far;
            const float cx = p[far * sn], cy = p[far * sn + sc], cz = p[far * sn + 2 * sc];
            float best = -INFINITY;
            int64_t besti = 0;
            for (int i = 0; i < N; ++i) {
                const float dx = p[i * sn] - cx, dy = p[i * sn + sc] - cy, dz = p[i * sn + 2 * sc] - cz;
                const float d = ((dx * dx) + (dy * dy)) + (dz * dz);
                if (d < mind[i]) mind[i] = d;
                /* torch.max: first index of the maximum; mind[] never holds NaN (init 1e10,
                 * only replaced by values that compared '<'), so no NaN branch is needed. */
                if (mind[i] > best) { best = mind[i]; besti = i; }
            }
            far = besti;
        }
        free(mind);
    }
}

/* ------------------------------------------------------------------------------------------
 * Expanded-form squared distance.  models/pointnet2_utils.py:21-42
 *   :39 dist  = -2 * matmul(src, dst^T)   inner dim 3 on CPU == fma(s2,d2, fma(s1,d1, s0*d0))
 *   :40 dist += sum(src**2, -1)           (s0*s0 + s1*s1) + s2*s2, unfused
 *   :41 dist += sum(dst**2, -1)
 * ------------------------------------------------------------------------------------------ */
static inline float orc_sqdist_expanded(const float *s, const float *d)
{
    const float dot = fmaf(s[2], d[2], fmaf(s[1], d[1], s[0] * d[0]));
    const float sn = ((s[0] * s[0]) + (s[1] * s[1])) + (s[2] * s[2]);
    const float dn = ((d[0] * d[0]) + (d[1] * d[1])) + (d[2] * d[2]);
    return ((-2.0f * dot) + sn) + dn;
}

ORC_API void orc_square_distance_f32(const float *src, const float *dst, int B, int N, int M,
                                     float *out /* [B,N,M] */)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < N; ++i) {
            const float *s = src + ((int64_t)b * N + i) * 3;
            for (int j = 0; j < M; ++j)
                out[((int64_t)b * N + i) * M + j] = orc_sqdist_expanded(s, dst + ((int64_t)b * M + j) * 3);
        }
}

/* ------------------------------------------------------------------------------------------
 * Ball query.  models/pointnet2_utils.py:89-109
 *   :102 group_idx = arange(N) per query
 *   :103 sqrdists = square_distance(new_xyz, xyz)          (queries are `src`)
 *   :104 group_idx[sqrdists > radius**2] = N               -> in-ball iff !(d > fp32(r^2))
 *   :105 sort ascending, keep first nsample
 *   :106-108 entries still equal to N are replaced by the first entry
 *            (so an empty ball yields N everywhere, exactly like the reference).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_ball_query_f32(const float *xyz, const float *new_xyz, int B, int N, int S, float r2,
                                int nsample, int64_t *out /* [B,S,nsample] */)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < S; ++s) {
            const float *q = new_xyz + ((int64_t)b * S + s) * 3;
            int64_t *o = out + ((int64_t)b * S + s) * nsample;
            int cnt = 0;
            for (int i = 0; i < N && cnt < nsample; ++i) {
                const float d = orc_sqdist_expanded(q, xyz + ((int64_t)b * N + i) * 3);
                if (!(d > r2)) o[cnt++] = i;
            }
            const int64_t first = cnt > 0 ? o[0] : (int64_t)N;
            for (int k = cnt; k < nsample; ++k) o[k] = first;
        }
}

/* ------------------------------------------------------------------------------------------
 * K-nearest neighbours, pytorch3d.ops.knn.knn_points semantics (pytorch3d v0.7.x, not vendored in
 * the reference; call sites pytorch3d_chamfer.py:182-183, 205-206, 257-258):
 *   dists = squared L2 in direct form, accumulated over d = 0..D-1 in order;
 *   only the first len2[n] points of p2 are candidates; rows i >= len1[n] stay zero;
 *   K results sorted ascending, ties resolved to the lowest index (strict '<' insertion).
 * `use_fma` selects acc = fmaf(diff,diff,acc) (what nvcc emits for pytorch3d's CUDA loop) versus
 * acc += diff*diff (what its CPU loop does without -mfma); both are within the stated tolerance.
 * If len2[n] < K the missing slots keep dist 0 / idx 0... pytorch3d pads with -1 idx / inf? It
 * leaves them at their zero-initialised value; we do the same.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_knn_f32(const float *p1, const float *p2, int N, int P1, int P2, int D,
                         const int64_t *len1, const int64_t *len2, int K, int use_fma,
                         float *dists /* [N,P1,K] */, int64_t *idx /* [N,P1,K] */)
{
    memset(dists, 0, sizeof(float) * (size_t)N * P1 * K);
    memset(idx, 0, sizeof(int64_t) * (size_t)N * P1 * K);
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int i = 0; i < P1; ++i) {
            const int64_t l1 = len1 ? len1[n] : P1, l2 = len2 ? len2[n] : P2;
            if (i >= l1) continue;
            const float *a = p1 + ((int64_t)n * P1 + i) * D;
            float bd[16];
            int64_t bi[16];
            int have = 0;
            const int KK = K > 16 ? 16 : K;
            for (int64_t j = 0; j < l2; ++j) {
                const float *c = p2 + ((int64_t)n * P2 + j) * D;
                float acc = 0.0f;
                if (use_fma)
                    for (int d = 0; d < D; ++d) { const float df = a[d] - c[d]; acc = fmaf(df, df, acc); }
                else
                    for (int d = 0; d < D; ++d) { const float df = a[d] - c[d]; acc += df * df; }
                if (have < KK) {
                    int pos = have++;
                    while (pos > 0 && acc < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                    bd[pos] = acc; bi[pos] = j;
                } else if (acc < bd[KK - 1]) {
                    int pos = KK - 1;
                    while (pos > 0 && acc < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                    bd[pos] = acc; bi[pos] = j;
                }
            }
            for (int k = 0; k < have; ++k) {
                dists[((int64_t)n * P1 + i) * K + k] = bd[k];
                idx[((int64_t)n * P1 + i) * K + k] = bi[k];
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * kNN grouping for the 100k-point stress configuration (BASELINE.json configs[4]).  No reference
 * function exists; SURVEY.md section 8(d) defines the oracle as square_distance (:21-42, expanded
 * form) followed by topk(k, largest=False, sorted=True).  Ties are unordered in torch, so tests
 * compare distance multisets; this restatement breaks ties towards the lowest index.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_knn_group_f32(const float *xyz, const float *new_xyz, int B, int N, int S, int K,
                               int64_t *out_idx /* [B,S,K] */, float *out_d /* [B,S,K] or NULL */)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int s = 0; s < S; ++s) {
            const float *q = new_xyz + ((int64_t)b * S + s) * 3;
            float bd[128];
            int64_t bi[128];
            int have = 0;
            const int KK = K > 128 ? 128 : K;
            for (int i = 0; i < N; ++i) {
                const float d = orc_sqdist_expanded(q, xyz + ((int64_t)b * N + i) * 3);
                if (have < KK || d < bd[KK - 1]) {
                    int pos = have < KK ? have++ : KK - 1;
                    while (pos > 0 && d < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                    bd[pos] = d; bi[pos] = i;
                }
            }
            for (int k = 0; k < KK; ++k) {
                out_idx[((int64_t)b * S + s) * K + k] = k < have ? bi[k] : 0;
                if (out_d) out_d[((int64_t)b * S + s) * K + k] = k < have ? bd[k] : 0.0f;
            }
        }
}

ORC_API int orc_version(void) { return 1; }

/* Thread count of the OpenMP loops above (torchrun exports OMP_NUM_THREADS=1; the timed CPU arm asks for all cores). */
#ifdef _OPENMP
#include <omp.h>
ORC_API int orc_set_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
#else
ORC_API int orc_set_threads(int n) { (void)n; return 1; }
#endif
