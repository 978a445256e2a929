/*
 * maskplanner_b200.h -- C ABI of libmaskplanner_b200.so (sm_100a).
 *
 * The reference (gabrieletiboni/MaskPlanner) has no FFI layer: its boundary for this path is the
 * Python module surface of models/pointnet2_utils.py and pytorch3d_chamfer.py.  Each entry point
 * below is what a binding for one of those functions would call; the citation names the reference
 * interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - strides are in ELEMENTS (not bytes) so permuted torch views can be passed without a copy;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing allocates: the caller owns outputs and workspaces;
 *   - return value: 0 on success, negative mpb_status otherwise; mpb_last_error_string() describes
 *     the last failure on the calling thread.  Nothing throws, nothing synchronises the device.
 *   - index tensors are int64 (torch.long), exactly as the reference returns them.
 */
#ifndef MASKPLANNER_B200_H_
#define MASKPLANNER_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MPB_API __attribute__((visibility("default")))
#else
#define MPB_API
#endif

typedef enum {
    MPB_OK = 0,
    MPB_ERR_INVALID_ARGUMENT = -1,
    MPB_ERR_UNSUPPORTED = -2,
    MPB_ERR_CUDA = -3,
    MPB_ERR_WORKSPACE = -4
} mpb_status;

MPB_API int mpb_version(void);
MPB_API const char *mpb_last_error_string(void);
/* Number of SMs / compute capability (major*10+minor) of the current device; <0 on error. */
MPB_API int mpb_device_sm_count(void);
MPB_API int mpb_device_arch(void);

/* ---- measurement helper (the FP32 SIMT roof "must be measured by the builder") ----
 * Launches sm_count*ctas_per_sm CTAs of 256 threads, each thread running `iters` rounds of 16 independent
 * fma chains (packed != 0: fma.rn.f32x2, two per lane per instruction).  out holds one float per thread
 * (mpb_peak_fp32_threads(ctas_per_sm) of them).  flops = 2 * threads * iters * 16 * (packed ? 2 : 1). */
MPB_API int64_t mpb_peak_fp32_threads(int ctas_per_sm);
MPB_API int mpb_peak_fp32_ffma(int packed, int iters, int ctas_per_sm, float *out, void *stream);

/* ---- a1: farthest_point_sample(xyz, npoint)            models/pointnet2_utils.py:65-86 -------
 * xyz [B,N,3] f32 through (sb,sn,sc); seed_idx [B] i64 = the reference's CPU randint draw (:77),
 * made by the caller; out_idx [B,npoint] i64.  Distance ((dx*dx)+(dy*dy))+(dz*dz), every op
 * rounded; strict '<' running minimum from 1e10; lowest index among equal maxima.
 * Register/cluster-resident up to N = 131072; beyond that `workspace` must hold B*N floats.
 * mpb_fps_workspace_bytes() returns the workspace the call needs (0 if none). */
MPB_API int64_t mpb_fps_workspace_bytes(int B, int N);
MPB_API int mpb_fps_f32(const float *xyz, int64_t sb, int64_t sn, int64_t sc, int B, int N,
                        const int64_t *seed_idx, int npoint, int64_t *out_idx, void *workspace,
                        void *stream);

/* ---- a2: square_distance(src, dst)                     models/pointnet2_utils.py:21-42 -------
 * Expanded form ((-2*dot)+|s|^2)+|d|^2 with dot = fma(s2,d2,fma(s1,d1,s0*d0)); C = 3 only (the
 * reference's other use, C = feature dim in feature propagation, is off the path).
 * src [B,N,3], dst [B,M,3] contiguous; out [B,N,M] f32. */
MPB_API int mpb_square_distance_f32(const float *src, const float *dst, int B, int N, int M,
                                    float *out, void *stream);

/* ---- a3: query_ball_point(radius, nsample, xyz, new_xyz)   models/pointnet2_utils.py:89-109 --
 * First `nsample` point indices in ascending order with !(d > r2), d as in a2, padded with the
 * first hit; a query with an empty ball gets N in every slot, like the reference.
 * r2 is fp32(radius**2), rounded by the caller exactly as torch rounds the Python scalar. */
MPB_API int mpb_ball_query_f32(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc,
                               const float *new_xyz, int64_t qsb, int64_t qsn, int64_t qsc,
                               int B, int N, int S, float r2, int nsample, int64_t *out_idx,
                               void *stream);

/* ---- stress config: kNN grouping (no reference function; SURVEY.md 8d defines it as
 * square_distance + topk(k, smallest, sorted)).  Ties go to the lowest index.  k <= 64.
 * out_dist may be NULL. */
MPB_API int mpb_knn_group_f32(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc,
                              const float *new_xyz, int64_t qsb, int64_t qsn, int64_t qsc,
                              int B, int N, int S, int k, int64_t *out_idx, float *out_dist,
                              void *stream);

/* ---- a4: index_points(points, idx)                     models/pointnet2_utils.py:45-62 -------
 * out[b,m,:] = points[b, idx[b,m], :]; points [B,N,C] through strides, idx [B,M] i64 (M = S or
 * S*K), out [B,M,C] contiguous.  Out-of-range indices (the reference would raise) produce zeros.
 * _bwd: grad_points[b, idx[b,m], :] += grad_out[b,m,:] (grad_points [B,N,C] contiguous, caller
 * zero-fills it first). */
MPB_API int mpb_index_points_f32(const float *points, int64_t sb, int64_t sn, int64_t sc, int B,
                                 int N, int C, const int64_t *idx, int64_t M, float *out,
                                 void *stream);
MPB_API int mpb_index_points_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int C,
                                     int64_t M, float *grad_points, void *stream);

/* ---- a5: sample_and_group's gather/centre/concat       models/pointnet2_utils.py:133-141 -----
 * out[b,s,k,0:3]   = xyz[b, idx[b,s,k], :] - new_xyz[b,s,:]
 * out[b,s,k,3:3+D] = feats[b, idx[b,s,k], :]          (feats may be NULL, D = 0)
 * out is [B,S,K,ldo] f32 with ldo >= 3+D (columns beyond 3+D are zero-filled: GEMM K padding).
 * _bwd scatters grad_out[..., 3:3+D] into grad_feats [B,N,D] (contiguous, caller zero-fills);
 * xyz carries no gradient in the reference model (input cloud / index-selected coordinates are
 * never parameters), grad_xyz may be NULL; when given it receives the :134 terms too. */
MPB_API int mpb_group_points_f32(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc,
                                 const float *feats, int64_t fsb, int64_t fsn, int64_t fsc,
                                 const float *new_xyz, const int64_t *idx, int B, int N, int S,
                                 int K, int D, int ldo, float *out, void *stream);
MPB_API int mpb_group_points_bwd_f32(const float *grad_out, int ldo, const int64_t *idx, int B,
                                     int N, int S, int K, int D, float *grad_feats,
                                     float *grad_xyz, float *grad_new_xyz, void *stream);

/* ---- a8: chamfer_distance's nearest-neighbour core     pytorch3d_chamfer.py:257-258 ----------
 * (pytorch3d.ops.knn.knn_points with K = 1, both directions from one launch.)
 * x [N,P1,D], y [N,P2,D] f32 contiguous; x_len / y_len [N] i64 or NULL (= full length).
 * dist_x/idx_x [N,P1] : for every x point the squared distance to / index of its nearest y point
 * among the first y_len[n]; rows >= x_len[n] are written as 0.  dist_y/idx_y [N,P2] likewise.
 * Either output pair may be NULL: that direction is skipped (asymmetric modes, :329-332).
 * Squared distance in direct form, fma-accumulated over d = 0..D-1; strict '<' => lowest index
 * wins ties.  D <= 64. */
MPB_API int mpb_chamfer_nn_f32(const float *x, const float *y, int N, int P1, int P2, int D,
                               const int64_t *x_len, const int64_t *y_len, float *dist_x,
                               int64_t *idx_x, float *dist_y, int64_t *idx_y, void *stream);

/* Backward of the above (pytorch3d knn_points_backward, K = 1):
 * grad_x[n,i] = 2*gdx[n,i]*(x[n,i]-y[n,idx_x[n,i]]) + sum_{j: idx_y[n,j]=i} 2*gdy[n,j]*(x[n,i]-y[n,j])
 * and symmetrically for grad_y.  gdx/gdy (and their idx) may be NULL.  grad_x / grad_y are
 * fully written (no pre-zeroing needed). */
MPB_API int mpb_chamfer_nn_bwd_f32(const float *x, const float *y, int N, int P1, int P2, int D,
                                   const int64_t *x_len, const int64_t *y_len, const int64_t *idx_x,
                                   const int64_t *idx_y, const float *gdx, const float *gdy,
                                   float *grad_x, float *grad_y, void *stream);

/* General-K variant for the K = 2 branches (pytorch3d_chamfer.py:205-206): one direction,
 * dists/idx [N,P1,K] sorted ascending, K <= 8. */
MPB_API int mpb_knn_points_f32(const float *p1, const float *p2, int N, int P1, int P2, int D,
                               const int64_t *len1, const int64_t *len2, int K, float *dists,
                               int64_t *idx, void *stream);
MPB_API int mpb_knn_points_bwd_f32(const float *p1, const float *p2, int N, int P1, int P2, int D,
                                   const int64_t *len1, const int64_t *len2, const int64_t *idx,
                                   int K, const float *grad_dists, float *grad_p1, float *grad_p2,
                                   void *stream);

/* ---- a7: PointNetSetAbstraction's shared MLP        models/pointnet2_utils.py:208-214 -------------
 * relu(bn(conv1x1(x))) per layer then max over the K neighbours.  The 1x1 conv over [B,C,K,S] is the
 * row-wise GEMM Z[M,Cout] = A[M,Cin] * W[Cout,Cin]^T (M = B*S*K); activations travel as bf16 or fp32 [M,C]
 * row-major with C a multiple of 64 (zero padded), accumulation / statistics in fp32.
 *
 * mpb_group_points_bf16 / _bwd_bf16: a5 emitted as bf16 GEMM rows [B*S*K, ldo], columns in the order
 *                      [feats (D), centred xyz (3), zero padding]; ldo % 8 == 0.  _bwd adds the feature
 *                      columns of grad_out into grad_feats [B,N,D] (fp32, caller zero-fills).
 * mpb_sa_gemm_tn:      C[M,N] = f(A)[M,K] * B[N,K]^T on tcgen05 + TMA.  Forward (B = W) and dgrad (A = dZ, B = W^T).
 *                      dtype 0: bf16 operands/outputs (K % 64 == 0); 1: fp32 storage, one TF32 pass; 2: fp32 storage,
 *                      3xTF32 (hi/lo split of A in shared memory, B_lo = W - tf32(W) supplied by the caller) --
 *                      the reference-precision mode (K % 32 == 0).  N % 32 == 0.
 *                      a_scale/a_shift (optional, [K]): f(a) = relu(a_scale[k]*a + a_shift[k]) applied to the A tile
 *                      between TMA landing and MMA issue -- the previous layer's BatchNorm + ReLU (:211-212), so
 *                      the normalised activation never exists in HBM.
 *                      epi 1: also write per-column (sum, sum of squares) partial rows of the stored C (the
 *                      layer's training-mode BatchNorm statistics); epi 2: also write per-column (sum dY,
 *                      sum dY*z) with dY = C * [z_scale*z + z_shift > 0] against Z [M,N] (same dtype as C): the
 *                      BatchNorm-backward statistics of the layer below, fused into the dgrad GEMM.
 *                      partials: [nparts][2][N] fp32, nparts = mpb_sa_gemm_stat_partials(...) (0: cannot fuse).
 * mpb_sa_gemm_wgrad:   dW[cout,cin] (fp32, reference layout, overwritten) = crop/unpermute(dZ[M,N]^T * f(A)[M,K]);
 *                      both operands read as MN-major straight from their row-major layout; the contraction
 *                      rows are split over CTAs whose partial tiles go to `workspace`
 *                      (mpb_sa_gemm_wgrad_workspace bytes) and are summed in a fixed order: bitwise reproducible.
 *                      xyz_last undoes mpb_pack_weight_*'s column permutation; accumulate != 0 adds to dW instead.
 * mpb_bn_*:            training-mode BatchNorm2d statistics, normalise + ReLU (+ max-pool with arg-max),
 *                      and the matching backward; `partials` is a [nparts, 2, C] fp32 workspace with
 *                      nparts = mpb_bn_stat_partials(rows, C).  running_mean / running_var are updated in
 *                      place (momentum, unbiased variance) like nn.BatchNorm2d; the conv bias only enters
 *                      the running mean.  coef is a [3, C] fp32 workspace produced by _bwd_finalize. */
MPB_API int mpb_group_points_bf16(const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc,
                                  const float *feats, int64_t fsb, int64_t fsn, int64_t fsc,
                                  const float *new_xyz, const int64_t *idx, int B, int N, int S,
                                  int K, int D, int ldo, void *out, void *stream);
MPB_API int mpb_group_points_bwd_bf16(const void *grad_out, int ldo, const int64_t *idx, int B,
                                      int N, int S, int K, int D, float *grad_feats, void *stream);
MPB_API int mpb_sa_gemm_stat_partials(int dtype, int M, int N, int K, int xform, int epi);
MPB_API int mpb_sa_gemm_tn(int dtype, const void *A, const void *B, const void *B_lo, void *C, int M,
                           int N, int K, const float *a_scale, const float *a_shift, int epi,
                           float *partials, int nparts, const void *Z, const float *z_scale,
                           const float *z_shift, void *stream);
MPB_API int64_t mpb_sa_gemm_wgrad_workspace(int dtype, int M, int N, int K, int xform);
MPB_API int mpb_sa_gemm_wgrad(int dtype, const void *dZ, const void *A, int M, int N, int K,
                              const float *a_scale, const float *a_shift, float *workspace, int cout,
                              int cin, int xyz_last, int accumulate, float *dW, void *stream);
/* accumulate = -1 in mpb_sa_gemm_wgrad writes the per-split partial tiles only; this call sums them (fixed order) into dW.
 * Split so the caller can run the reduction on another stream: it feeds the optimizer, not the next GEMM. */
/* The two GEMMs that consume dZ of a MAX-POOLED layer, with dZ rebuilt in shared memory from that layer's stored
 * pre-activation Zl (bf16) instead of being written by mpb_bn_bwd_apply and read back twice:
 *   dZ[m,c] = pgo[g,c] * [m - g*pool_k == argmax[g,c]] + negw_e[0][c] * Zl[m,c] + negw_e[1][c],  g = m / pool_k.
 * Everything else as in mpb_sa_gemm_tn (epi 0 / 2) and mpb_sa_gemm_wgrad.          pointnet2_utils.py:210-214 (autograd) */
MPB_API int mpb_sa_gemm_tn_pool(int dtype, const void *Zl, const void *B, void *C, int M, int N, int K, int pool_k,
                                const int32_t *argmax, const float *pgo, const float *negw_e, int epi,
                                float *partials, int nparts, const void *Z, const float *z_scale,
                                const float *z_shift, void *stream);
MPB_API int mpb_sa_gemm_wgrad_pool(int dtype, const void *Zl, const void *A, int M, int N, int K,
                                   const float *a_scale, const float *a_shift, int pool_k, const int32_t *argmax,
                                   const float *pgo, const float *negw_e, float *workspace, int cout, int cin,
                                   int xyz_last, int accumulate, float *dW, void *stream);
MPB_API int mpb_sa_gemm_wgrad_reduce(int dtype, int M, int N, int K, int xform, const float *workspace, int cout,
                                     int cin, int xyz_last, int accumulate, float *dW, void *stream);
MPB_API int mpb_bn_stat_partials(int64_t rows, int C);
/* dtype of the mpb_bn_* / mpb_sa_first_layer* calls: storage type of the activations, 0 = bf16, 1 = fp32. */
MPB_API int mpb_bn_colstats(int dtype, const void *Z, int64_t M, int C, float *partials, int nparts,
                            void *stream);
/* C_valid <= C: channels >= C_valid are alignment padding (scale = shift = 0, no parameter access).
 * num_batches_tracked (optional, device int64): BatchNorm's batch counter, incremented by this launch. */
MPB_API int mpb_bn_finalize_f32(const float *partials, int nparts, int C, int C_valid, int64_t M,
                                const float *bias, const float *gamma, const float *beta,
                                float *running_mean, float *running_var, float momentum, float eps,
                                float *scale, float *shift, float *mean, float *rstd,
                                int64_t *num_batches_tracked, void *stream);
/* A = relu(scale*Z + shift) as a separate pass (A/B switch MPB_FUSE_APPLY=0; the default path applies it inside
 * the consumer GEMM's operand path instead, mpb_sa_gemm_tn a_scale/a_shift). */
MPB_API int mpb_bn_relu(int dtype, const void *Z, const float *scale, const float *shift, int64_t M,
                        int C, void *A, void *stream);
/* zmax (optional, fp32 [G,C]): the pre-activation Z at the arg-max row, consumed by _bwd_stats (pooled). */
MPB_API int mpb_bn_relu_max(int dtype, const void *Z, const float *scale, const float *shift,
                            int64_t G, int K, int C, float *out, int32_t *argmax, float *zmax,
                            void *stream);
/* Backward: exactly one of dA (dense upstream gradient, activation dtype [M,C]) and dOut (pooled upstream
 * gradient, fp32 [M/K,C], with argmax int32 [M/K,C]) is non-NULL.  partials hold the raw moments
 * (sum dY, sum dY*z), dY = upstream * [scale*z + shift > 0]; _bwd_finalize turns them into dgamma, dbeta and coef
 * and zero-fills `clear` (clear_count floats, multiple of 4) as a side job. */
MPB_API int mpb_bn_bwd_stats(int dtype, const void *dA, const float *dOut, const int32_t *argmax,
                             const float *zmax, int K, const void *Z, const float *scale,
                             const float *shift, const float *mean, const float *rstd, int64_t M, int C,
                             float *partials, int nparts, float *pgo, void *stream);
MPB_API int mpb_bn_bwd_finalize_f32(const float *partials, int nparts, int C, int C_valid, int64_t M,
                                    const float *gamma, const float *mean, const float *rstd,
                                    float *dgamma, float *dbeta, float *coef, float *clear,
                                    int64_t clear_count, float *negw_e, void *stream);
/* dZ of a max-pooled layer from the by-products of the statistics pass (pgo, negw_e below): the lean form of the pooled
 * mpb_bn_bwd_apply (16 instead of 40 per-channel registers, no ReLU re-evaluation). */
MPB_API int mpb_bn_bwd_apply_pooled(int dtype, const float *pgo, const int32_t *argmax, int K, const void *Z,
                                    const float *negw_e, int64_t M, int C, void *dZ, void *stream);
/* pgo (pooled form of _bwd_stats, optional, fp32 [G,C]) and negw_e (_bwd_finalize, optional, fp32 [2,C]) feed the POOLED
 * operand transform of mpb_sa_gemm_tn_pool / mpb_sa_gemm_wgrad_pool: with dZ = p*dY - w*z + e and dY non-zero only at the
 * arg-max row of each (group, channel), pgo[g,c] = p[c]*dY[g,c] and negw_e = (-w, e); the GEMMs then rebuild the dZ tile
 * from the stored pre-activation z in shared memory and mpb_bn_bwd_apply (one read + one write of [M,C]) is not needed. */
MPB_API int mpb_bn_bwd_apply(int dtype, const void *dA, const float *dOut, const int32_t *argmax, int K,
                             const void *Z, const float *scale, const float *shift, const float *mean,
                             const float *rstd, const float *coef, int64_t M, int C, void *dZ,
                             void *stream);
/* Weight packing.  _bf16: conv weight fp32 [cout,cin] -> bf16 Wp [cout_p,cin_p] (zero padded; xyz_last moves the 3
 * leading input channels behind the features) and its transpose Wt.  _tf32: the same as fp32 with the 3xTF32 split,
 * hi = tf32(w) in (Wp_hi, Wt_hi) and lo = w - hi in (Wp_lo, Wt_lo); both lo pointers NULL: unsplit fp32 copy. */
MPB_API int mpb_pack_weight_bf16(const float *W, int cout, int cin, int cout_p, int cin_p, int xyz_last,
                                 void *Wp, void *Wt, void *stream);
MPB_API int mpb_pack_weight_tf32(const float *W, int cout, int cin, int cout_p, int cin_p, int xyz_last,
                                 float *Wp_hi, float *Wp_lo, float *Wt_hi, float *Wt_lo, void *stream);
/* On-the-fly first SA layer (3 + D <= 8 input channels): Z = row . W^T with the row gathered from xyz / feats /
 * new_xyz / idx (never materialised), BatchNorm statistic partials fused; _bwd re-gathers the row, forms dZ like
 * mpb_bn_bwd_apply and ACCUMULATES dW [C, ldw] (caller zero-fills).  W [C, ldw] in the activation dtype, xyz-last
 * column order (fp32: the unsplit weight -- the kernel's CUDA-core FMAs are exact fp32). */
MPB_API int mpb_sa_first_layer(int dtype, const float *xyz, int64_t xsb, int64_t xsn, int64_t xsc,
                               const float *feats, int64_t fsb, int64_t fsn, int64_t fsc,
                               const float *new_xyz, const int64_t *idx, int B, int N, int S, int K,
                               int D, const void *W, int ldw, int C, void *Z, float *partials,
                               int nparts, void *stream);
MPB_API int mpb_sa_first_layer_bwd(int dtype, const void *dA, const void *Z, const float *scale,
                                   const float *shift, const float *mean, const float *rstd,
                                   const float *coef, const float *xyz, int64_t xsb, int64_t xsn,
                                   int64_t xsc, const float *feats, int64_t fsb, int64_t fsn,
                                   int64_t fsc, const float *new_xyz, const int64_t *idx, int B, int N,
                                   int S, int K, int D, int C, float *dW, int ldw, void *stream);

/* ---- f3: regression heads                             models/pointnet2_cls_ssg.py:270-295, :309-341 ----------
 * The head GEMMs reuse mpb_sa_gemm_tn / mpb_sa_gemm_wgrad (dtype 1 = TF32, 2 = 3xTF32) with the WEIGHT as the 128-row
 * operand, so head activations are FEATURE-MAJOR [F, Bp], Bp = batch padded to a multiple of 32 (<= 128):
 *   forward  Yt[Nout,Bp] = W[Nout,Kin] * X[Bp,Kin]^T      mpb_sa_gemm_tn(dtype, A=W, B=X, C=Yt, M=Nout, N=Bp, K=Kin)
 *   dW       dW[Nout,Kin] = dYt[Nout,Bp] * Xt[Kin,Bp]^T   mpb_sa_gemm_tn(dtype, A=dYt, B=Xt, C=dW, M=Nout, N=Kin, K=Bp)
 *   dX       dXt[Kin,Bp] = W^T * dYt                       mpb_sa_gemm_wgrad(dtype, dZ=W, A=dYt, M=Nout, N=Kin, K=Bp)
 * mpb_head_act_fwd:  bias + BatchNorm1d (training: batch statistics over the B real columns, running stats updated;
 *                    eval: running stats) + ReLU + dropout(p), one warp per feature.  Outputs the activation in both
 *                    layouts: Xt [F,Bp] and X [Bp,F] (padding rows/columns zero); Xt_lo/X_lo non-NULL: hi/lo TF32 split.
 *                    Dropout keeps element (f,b) iff hash(seed, *step, layer, f, b) >= p: nothing is stored.
 * mpb_head_act_bwd:  dXt [F,Bp] -> dYt [F,Bp], dgamma, dbeta, dbias [F] (mask regenerated from the same hash).
 * mpb_rng_advance:   *counter += 1 on the stream (once per forward; a captured step advances it on replay).
 * mpb_head_to_feature_major / _to_batch_major: [B,F] <-> [F,Bp] (+ bias on the way out, + column sums on the way in).
 * mpb_head_pose_out_*: out[b,j,:] = [fc3[3j..3j+2], w * normalize(tanh(fc_normals[3j..3j+2]))]  (:332-339) and its backward. */
MPB_API int mpb_rng_advance(int64_t *counter, void *stream);
MPB_API int mpb_head_act_fwd(const float *Yt, int F, int B, int Bp, const float *bias, const float *gamma,
                             const float *beta, float *running_mean, float *running_var, float momentum,
                             float eps, int training, float drop_p, uint64_t seed, const int64_t *step,
                             int layer, float *Xt, float *Xt_lo, float *X, float *X_lo, float *mean,
                             float *rstd, void *stream);
MPB_API int mpb_head_act_bwd(const float *dXt, const float *Yt, int F, int B, int Bp, const float *bias,
                             const float *gamma, const float *beta, const float *mean, const float *rstd,
                             int training, float drop_p, uint64_t seed, const int64_t *step, int layer,
                             float *dYt, float *dgamma, float *dbeta, float *dbias, void *stream);
MPB_API int mpb_head_to_feature_major(const float *X, int B, int Bp, int F, float *Xt, float *Xt_lo,
                                      float *colsum, void *stream);
MPB_API int mpb_head_to_batch_major(const float *Yt, const float *bias, int B, int Bp, int F, float *Y,
                                    void *stream);
MPB_API int mpb_head_pose_out_fwd(const float *Yt3, const float *b3, const float *Ytn, const float *bn,
                                  int B, int Bp, int P, float weight_orient, float *out, void *stream);
MPB_API int mpb_head_pose_out_bwd(const float *d_out, const float *Ytn, const float *bn, int B, int Bp, int P,
                                  float weight_orient, float *dYt3, float *dYtn, float *db3, float *dbn,
                                  void *stream);

/* `padded=True` length scan                             pytorch3d_chamfer.py:138-149 ----------
 * first[n] = first j with y[n,j,0] == sentinel, else P2; *any_flag (int32, caller zero-fills) is
 * set to 1 if any sample is padded.  No host synchronisation (the reference does 2N of them). */
MPB_API int mpb_padded_lengths_f32(const float *y, int N, int P2, int D, float sentinel,
                                   int64_t *first, int32_t *any_flag, void *stream);

/* ---- f3: optimizer step                               train_maskplanner.py:159, :221 -------------------
 * torch.optim.Adam semantics (amsgrad = False, maximize = False) for up to 80 fp32 tensors in ONE launch.
 * params / grads / exp_avg / exp_avg_sq / numel are HOST arrays of length ntensors (device pointers and
 * element counts); they are copied into the kernel's parameter block, so the call is CUDA-graph capturable
 * and needs no device-side tables.  step (device float, the 1-based count of completed steps) is read by
 * the launch and advanced by its last CTA; ticket is a zero-initialised device counter owned by the caller.
 * lr_dev (optional device float) overrides lr, so a scheduler can change it between graph replays.
 * grad_scale multiplies every gradient on load (1/world_size when the buffer holds the SUM over ranks). */
MPB_API int mpb_adam_step_f32(int ntensors, float *const *params, const float *const *grads,
                              float *const *exp_avg, float *const *exp_avg_sq, const int64_t *numel,
                              float lr, const float *lr_dev, double beta1, double beta2, double eps,
                              double weight_decay, float grad_scale, float *step, uint32_t *ticket,
                              void *stream);

/* ---- f1 (next row): the mask loss's per-sample Hungarian matching   loss_handler.py:860-877 ------
 * One warp per sample solves min sum_t cost[b, row(t), t] over injective row(.) for the PRESENT
 * targets t (present[b,t] != 0; the reference's per-sample torch.unique), fp64 like scipy's
 * linear_sum_assignment.  cost [B,P,T] f32, present [B,T] u8, out_row [B,T] i64 (-1 for absent
 * targets).  Needs #present <= P <= 32, T <= 32.  Replaces B device->host copies per step. */
MPB_API int mpb_lap_f32(const float *cost, const uint8_t *present, int B, int P, int T,
                        int64_t *out_row, void *stream);

/* ---- a9 / f1: the asymm_v6 training loss around the nearest-neighbour kernels   loss_handler.py:596-666, :816-935 ----
 * The caller issues mpb_chamfer_nn_f32 twice (segments, both directions; poses, GT -> prediction) and mpb_lap_f32;
 * these entry points replace the ~130 small torch ops between them.
 *  mpb_loss_lengths_f32         `padded=True` length scan of traj [B,P2,D] and traj_as_pc [B,P3,D2] in one launch
 *                               (pytorch3d_chamfer.py:138-149: first row whose channel 0 is the sentinel).
 *  mpb_mask_cost_f32            ids_out[b,i] = (int) stroke_ids[b, match[b,i]] (:838); cost[b,p,t] = sum_i
 *                               BCEWithLogits(masks[b,p,i], [ids == t]) for all p, t (:860-873); present[b,t].
 *  mpb_asymm_v6_loss_value_f32  loss = w0*t1 + w1*t2 + w2*t3 + w3*mean matched BCE + w4*confidence BCE; weights5 is a
 *                               DEVICE array (schedulable between CUDA-graph replays); n_pairs_in (optional, device)
 *                               overrides the matched-pair normaliser (global count under data parallelism).
 *                               terms8 = {t1, t2, t3, weighted mask term, sum matched BCE, local pairs, confidence, pairs used}.
 *  mpb_asymm_v6_loss_bwd_f32    d loss / d y_pred [B,P1,D], d masks [B,NM,P1], d scores [B,NM], scaled by grad_loss[0];
 *                               independent of the value kernel (n_pairs_in: the same optional override). */
MPB_API int mpb_loss_lengths_f32(const float *traj, int P2, int D, const float *traj_as_pc, int P3, int D2, int B,
                                 float sentinel, int64_t *len_traj, int64_t *len_pc, void *stream);
MPB_API int mpb_mask_cost_f32(const float *masks, const float *stroke_ids, const int64_t *match, int B, int NM, int P1,
                              int P2, float *cost, uint8_t *present, int32_t *ids_out, void *stream);
MPB_API int mpb_asymm_v6_loss_value_f32(const float *d_x, const float *d_y, const int64_t *len_y, const float *d_y2,
                                        const int64_t *len_y2, const float *cost, const uint8_t *present,
                                        const int64_t *row, const float *scores, const float *weights5,
                                        const float *n_pairs_in, float no_stroke_w, int B, int P1, int P2, int P3,
                                        int NM, float *loss, float *terms8, void *stream);
MPB_API int mpb_asymm_v6_loss_bwd_f32(const float *y_pred, const float *traj, const float *traj_as_pc,
                                      const float *masks, const float *scores, const int64_t *idx_x,
                                      const int64_t *idx_y, const int64_t *len_y, const int64_t *idx_y2,
                                      const int64_t *len_y2, const int32_t *ids, const uint8_t *present,
                                      const int64_t *row, const float *weights5, const float *n_pairs_in,
                                      const float *grad_loss, float no_stroke_w, int B, int P1, int P2, int D, int P3,
                                      int D2, int NM, float *grad_pred, float *grad_masks, float *grad_scores,
                                      void *stream);

/* ---- the step's input boundary                        train_maskplanner.py:207-208, paintnet_ODv1.py:738-747 ----
 * One launch copies up to 8 tensors (as 32-bit words) into the fixed buffers a captured step reads, padding rows:
 * dst[b, r, :] = r < src_rows ? src[b, r, :] : pad_bits.  All arrays are HOST arrays of length nsegs. */
MPB_API int mpb_stage_batch(int nsegs, const void *const *src, void *const *dst, const int64_t *batch,
                            const int64_t *src_rows, const int64_t *dst_rows, const int64_t *row_words,
                            const uint32_t *pad_bits, void *stream);

/* Developer aid: in a library built with -DMPB_MBAR_DEBUG, where the first timed-out mbarrier wait of the tcgen05 GEMM
 * pipelines happened (source line, CTA, thread, ...); all zeros in a normal build. */
MPB_API int mpb_debug_mbar_state(int *out8, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MASKPLANNER_B200_H_ */
