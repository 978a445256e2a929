"""Tensor-core shared MLP of PointNetSetAbstraction: relu(bn(conv1x1)) x L, then max over neighbours.

Reference: models/pointnet2_utils.py:208-214.  Host-side orchestration only -- every stage is a
hand-written sm_100a kernel behind the C ABI (include/maskplanner_b200.h, section a7):

    forward, per layer   Z = A @ W^T            mpb_gemm_bf16_tn      (tcgen05 + TMA, fp32 accumulate in TMEM)
                         batch statistics       mpb_bn_colstats_bf16 + mpb_bn_finalize_f32 (running stats updated)
                         A' = relu(s*Z + t)     mpb_bn_relu_bf16      (last layer: mpb_bn_relu_max_bf16 = fused max-pool)
    backward, per layer  sum dY, sum dY*zhat    mpb_bn_bwd_stats_bf16 + mpb_bn_bwd_finalize_f32 (-> dgamma, dbeta)
                         dZ                     mpb_bn_bwd_apply_bf16
                         dW = dZ^T @ A          mpb_gemm_bf16_wgrad   (MN-major operands, no transposed copies)
                         dA = dZ @ W            mpb_gemm_bf16_tn

Activations are bf16 [M, C] row-major (M = B*S*K neighbourhood rows), channel counts zero-padded to
multiples of 64; statistics, pooled outputs and all parameter gradients are fp32.  In training mode the
conv bias cannot influence the output (BatchNorm removes it); it only enters the running mean, and its
gradient is exactly zero (the reference's value there is rounding noise).

Optional L2-resident chunking (off by default, see L2_CHUNK_BYTES).  Every [M, C] tensor of SA1/SA2
(134-268 MB at B = 64) is larger than the 126 MB L2, and training-mode BatchNorm puts a full-batch barrier
between layers, so producer and consumer kernels of one tensor cannot be fused across the whole batch.
They CAN be run back to back on a row chunk small enough to stay in L2: forward, per chunk
`bn_relu(l-1) -> GEMM(l) -> colstats(l)`; backward, per chunk `apply(l) -> wgrad(l) -> dgrad(l) ->
bwd_stats(l-1)`; only launch order and pointer offsets differ.  Measured on B200 this is a net loss at
B = 64, so the schedule below degenerates to one chunk unless MPB_L2_CHUNK_MB is set.
"""
import ctypes
import os

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr

# Rows per L2-resident chunk are chosen so that the widest producer/consumer pair of the stack
# (A chunk + Z chunk, bf16) stays below this many bytes; 0 disables chunking (one launch per stage).
# Measured on B200 at B = 64 (profiles/r01_notes.md): with 8 chunks the per-launch fixed cost of the two
# GEMM kernels (barrier/TMEM set-up, split-M reductions) outweighs the L2 hits, so the default is off.
L2_CHUNK_BYTES = int(os.environ.get("MPB_L2_CHUNK_MB", "0")) << 20


def pad64(c):
    return (c + 63) // 64 * 64


# Optional per-launch timing of the two GEMM kernels (bench.py's roofline leg): set to a list and every GEMM
# launch appends (kernel name, algorithmic bytes, flops, start event, end event), recorded on the launching stream.
GEMM_TIMELINE = None


FUSE_STATS = os.environ.get("MPB_FUSE_STATS", "1") == "1"   # BatchNorm statistics in the GEMM epilogue (N <= 256)


def _gemm_tn(lib, a_ptr, b_ptr, c_ptr, M, N, K, st, stats=None):
    """C[M,N] (bf16) = A[M,K] @ B[N,K]^T.  Algorithmic bytes: A and B read once, C written once, all bf16.
    stats = (partials pointer, nparts): also emit the per-column sum / sum-of-squares partials of C."""
    ev = None
    if GEMM_TIMELINE is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if stats is not None:
        check(lib.mpb_gemm_bf16_tn_stats(a_ptr, b_ptr, c_ptr, M, N, K, stats[0], stats[1], st), "mpb_gemm_bf16_tn_stats")
    else:
        check(lib.mpb_gemm_bf16_tn(a_ptr, b_ptr, c_ptr, M, N, K, 0, st), "mpb_gemm_bf16_tn")
    if ev is not None:
        ev[1].record()
        GEMM_TIMELINE.append(("gemm_tn_kernel", 2 * (M * K + N * K + M * N), 2 * M * N * K, ev[0], ev[1]))


def _gemm_wgrad(lib, dz_ptr, a_ptr, dw_ptr, M, N, K, st):
    """dW[N,K] (fp32) += dZ[M,N]^T @ A[M,K].  Algorithmic bytes: dZ and A read once (bf16), dW written once (fp32)."""
    ev = None
    if GEMM_TIMELINE is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    check(lib.mpb_gemm_bf16_wgrad(dz_ptr, a_ptr, dw_ptr, M, N, K, st), "mpb_gemm_bf16_wgrad")
    if ev is not None:
        ev[1].record()
        GEMM_TIMELINE.append(("wgrad_kernel", 2 * M * (N + K) + 4 * N * K, 2 * M * N * K, ev[0], ev[1]))


def _padded_weight(conv_weight, cout_p, cin_p, xyz_last):
    """[Cout,Cin,1,1] fp32 -> bf16 [cout_p, cin_p] (zero padded) and its transpose [cin_p, cout_p].
    xyz_last: the rows come from mpb_group_points_bf16 (features first, the 3 centred coordinates last),
    so the reference's xyz-first input channels (:137) move to the end."""
    cout, cin = conv_weight.shape[0], conv_weight.shape[1]
    dev = conv_weight.device
    w = torch.empty(cout_p, cin_p, dtype=torch.bfloat16, device=dev)
    wt = torch.empty(cin_p, cout_p, dtype=torch.bfloat16, device=dev)
    src = conv_weight.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    check(_cabi.load().mpb_pack_weight_bf16(ptr(src), cout, cin, cout_p, cin_p, 1 if xyz_last else 0, ptr(w), ptr(wt),
                                            stream_ptr()), "mpb_pack_weight_bf16")
    return w, wt


NARROW_LDW = 8          # leading dimension of the packed first-layer weight on the narrow path (3 + D <= 8 channels)


def narrow_rows_supported(points, K):
    """The on-the-fly first layer applies when the grouped row has at most 8 channels, its features (if any) need no
    gradient, and the schedule is not L2-chunked."""
    D = 0 if points is None else points.shape[2]
    return 3 + D <= NARROW_LDW and L2_CHUNK_BYTES <= 0 and not (points is not None and points.requires_grad) \
        and os.environ.get("MPB_NARROW_FIRST", "1") == "1"


def _narrow_args(narrow):
    """(xyz, feats|None, new_xyz, idx) -> the leading arguments of mpb_sa_first_layer[_bwd]_bf16."""
    xyz, feats, new_xyz, idx = narrow
    B, N, _ = xyz.shape
    _, S, K = idx.shape
    fs = tuple(feats.stride()) if feats is not None else (0, 0, 0)
    D = 0 if feats is None else feats.shape[2]
    return (ptr(xyz), *xyz.stride(), ptr(feats), *fs, ptr(new_xyz), ptr(idx), B, N, S, K, D)


def _unpermute_wgrad(dw, cout, cin, xyz_last):
    """Inverse of the column order used by _padded_weight, cropped to the real [Cout, Cin]."""
    if xyz_last and cin > 3:
        return torch.cat([dw[:cout, cin - 3:cin], dw[:cout, :cin - 3]], dim=1).reshape(cout, cin, 1, 1)
    return dw[:cout, :cin].reshape(cout, cin, 1, 1)


def _row_chunks(M, K, widest_pair_bytes_per_row):
    """Row ranges [(r0, r1), ...]: whole groups of K rows, a multiple of 128 rows (GEMM tile) where possible."""
    if L2_CHUNK_BYTES <= 0:
        return [(0, M)]
    rows = max(K, L2_CHUNK_BYTES // max(widest_pair_bytes_per_row, 1))
    unit = K * 128 // _gcd(K, 128)               # lcm(K, 128)
    rows = max(unit, rows // unit * unit) if rows >= unit else max(K, rows // K * K)
    return [(r0, min(M, r0 + rows)) for r0 in range(0, M, rows)]


def _gcd(a, b):
    while b:
        a, b = b, a % b
    return a


def _off(t, row):
    """Device pointer of row `row` of a contiguous 2-D tensor."""
    return ctypes.c_void_p(t.data_ptr() + row * t.shape[1] * t.element_size())


class SharedMLPMax(torch.autograd.Function):
    """pooled[G, C_L] = max_k relu(bn_L(... relu(bn_1(a0 @ W_1^T)) ...)) over the K rows of each group.

    apply(a0, K, training, momentum_eps, xyz_last, *flat) with
      a0    bf16 [M, pad64(Cin)], M = G*K -- or a NarrowRows tuple (xyz, feats|None, new_xyz, idx) when the grouped row
            has at most 8 channels: the first layer then gathers its input on the fly (mpb_sa_first_layer_bf16) and
            the [M, 64] operand is never written
      flat  per layer: conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var
      momentum_eps  tuple of (momentum, eps) per layer
    """

    @staticmethod
    def forward(ctx, a0, K, training, momentum_eps, xyz_last, *flat):
        lib = _cabi.load()
        L = len(flat) // 6
        narrow = a0 if isinstance(a0, tuple) else None
        if narrow is not None:
            n_xyz, n_feats, n_new, n_idx = narrow
            M = n_idx.numel()
            dev = n_xyz.device
            c_in_p = NARROW_LDW
            a0 = None
        else:
            M = a0.shape[0]
            dev = a0.device
            c_in_p = a0.shape[1]
        G = M // K
        st = stream_ptr()
        dims = []
        for l in range(L):
            cout, cin = flat[6 * l].shape[0], flat[6 * l].shape[1]
            dims.append((cout, cin, pad64(cout), c_in_p))
            c_in_p = pad64(cout)
        chunks = [(0, M)] if narrow is not None else _row_chunks(M, K, max(2 * (d[2] + d[3]) for d in dims))
        acts, zs, stats, wts = [a0], [], [], []
        a = a0
        out = argmax = None
        for l in range(L):
            W, bias, gamma, beta, rmean, rvar = flat[6 * l:6 * l + 6]
            cout, cin, cout_p, cin_p = dims[l]
            w, wt = _padded_weight(W, cout_p, cin_p, xyz_last and l == 0)
            z = torch.empty(M, cout_p, dtype=torch.bfloat16, device=dev)
            sc = torch.empty(4, cout_p, dtype=torch.float32, device=dev)      # rows: scale, shift, mean, rstd
            mom, eps = momentum_eps[l]
            # statistics fused into the GEMM epilogue when the layer fits one column tile (N <= 256), else a separate pass
            gemm_layer = l > 0 or narrow is None
            fuse = [lib.mpb_gemm_tn_stat_partials(r1 - r0, cout_p, cin_p) if (training and FUSE_STATS and gemm_layer) else 0 for r0, r1 in chunks]
            np_c = [f or lib.mpb_bn_stat_partials(r1 - r0, cout_p) for f, (r0, r1) in zip(fuse, chunks)]
            part = torch.empty(sum(np_c), 2, cout_p, dtype=torch.float32, device=dev) if training else None
            if l == 0 and narrow is not None:
                np_c = [lib.mpb_bn_stat_partials(M, cout_p)]
                part = torch.empty(np_c[0], 2, cout_p, dtype=torch.float32, device=dev)
                check(lib.mpb_sa_first_layer_bf16(*_narrow_args(narrow), ptr(w), cin_p, cout_p, ptr(z), ptr(part), np_c[0], st),
                      "mpb_sa_first_layer_bf16")
            if l > 0:
                a = torch.empty(M, cin_p, dtype=torch.bfloat16, device=dev)
                acts.append(a)
            p0 = 0
            for ci, (r0, r1) in enumerate(chunks if (l > 0 or narrow is None) else []):
                if l > 0:   # previous layer's normalise + ReLU for this chunk, consumed from L2 by the GEMM below
                    ps = stats[l - 1]
                    check(lib.mpb_bn_relu_bf16(_off(zs[l - 1], r0), ptr(ps[0]), ptr(ps[1]), r1 - r0, cin_p, _off(a, r0), st),
                          "mpb_bn_relu_bf16")
                pptr = _off(part.view(-1, 2 * cout_p), p0) if training else None
                _gemm_tn(lib, _off(a, r0), ptr(w), _off(z, r0), r1 - r0, cout_p, cin_p, st, stats=(pptr, fuse[ci]) if fuse[ci] else None)
                if training:
                    if not fuse[ci]:
                        check(lib.mpb_bn_colstats_bf16(_off(z, r0), r1 - r0, cout_p, pptr, np_c[ci], st), "mpb_bn_colstats_bf16")
                    p0 += np_c[ci]
            if training:
                check(lib.mpb_bn_finalize_f32(ptr(part), sum(np_c), cout_p, cout, M, ptr(bias), ptr(gamma), ptr(beta), ptr(rmean),
                                              ptr(rvar), mom, eps, ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), st),
                      "mpb_bn_finalize_f32")
            else:
                # eval: running statistics; the conv bias folds into the shift (tiny per-channel host-side math)
                sc.zero_()
                rstd = torch.rsqrt(rvar + eps)
                s = (gamma if gamma is not None else 1.0) * rstd
                b0 = bias if bias is not None else 0.0
                sc[0, :cout] = s
                sc[1, :cout] = (beta if beta is not None else 0.0) + (b0 - rmean) * s
                sc[2, :cout] = rmean - b0
                sc[3, :cout] = rstd
            zs.append(z)
            stats.append(sc)
            wts.append(wt)
        cl_p = dims[-1][2]
        out = torch.empty(G, cl_p, dtype=torch.float32, device=dev)
        argmax = torch.empty(G, cl_p, dtype=torch.int32, device=dev)
        zmax = torch.empty(G, cl_p, dtype=torch.float32, device=dev) if training else None
        sc = stats[-1]
        check(lib.mpb_bn_relu_max_bf16(ptr(zs[-1]), ptr(sc[0]), ptr(sc[1]), G, K, cl_p, ptr(out), ptr(argmax), ptr(zmax), st),
              "mpb_bn_relu_max_bf16")
        ctx.K, ctx.L, ctx.dims, ctx.training, ctx.xyz_last, ctx.chunks = K, L, dims, training, xyz_last, chunks
        ctx.narrow = narrow is not None
        ctx.save_for_backward(argmax, *acts, *zs, *stats, *wts, *[flat[6 * l + 2] for l in range(L)], zmax,
                              *(narrow if narrow is not None else ()))
        c_last = dims[-1][0]
        return out[:, :c_last] if c_last != out.shape[1] else out

    @staticmethod
    def backward(ctx, d_out):
        if not ctx.training:
            raise RuntimeError("SharedMLPMax: backward through eval-mode BatchNorm is not implemented on the tensor-core path; "
                               "use precision='fp32' for that")
        lib = _cabi.load()
        K, L, dims, chunks = ctx.K, ctx.L, ctx.dims, ctx.chunks
        saved = ctx.saved_tensors
        argmax = saved[0]
        acts, zs = saved[1:1 + L], saved[1 + L:1 + 2 * L]
        stats, wts, gammas = saved[1 + 2 * L:1 + 3 * L], saved[1 + 3 * L:1 + 4 * L], saved[1 + 4 * L:1 + 5 * L]
        zmax = saved[1 + 5 * L]
        narrow = tuple(saved[2 + 5 * L:6 + 5 * L]) if ctx.narrow else None
        M = zs[0].shape[0]
        G = M // K
        dev = d_out.device
        st = stream_ptr()
        cl, cl_p = dims[-1][0], dims[-1][2]
        if cl != cl_p:
            d_pool = torch.zeros(G, cl_p, dtype=torch.float32, device=dev)
            d_pool[:, :cl] = d_out
        else:
            d_pool = d_out.contiguous().float()
        grads = [None] * (6 * L)

        def alloc_partials(l, pooled):
            cp = dims[l][2]
            np_c = [lib.mpb_bn_stat_partials((r1 - r0) // K if pooled else r1 - r0, cp) for r0, r1 in chunks]
            return np_c, torch.empty(sum(np_c), 2 * cp, dtype=torch.float32, device=dev)

        # statistics of the last layer: the pooled upstream gradient touches one row per (group, channel)
        np_c, part = alloc_partials(L - 1, True)
        sc = stats[L - 1]
        p0 = 0
        for ci, (r0, r1) in enumerate(chunks):
            check(lib.mpb_bn_bwd_stats_bf16(None, _off(d_pool, r0 // K), _off(argmax, r0 // K), _off(zmax, r0 // K), K, _off(zs[L - 1], r0), ptr(sc[0]),
                                            ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), r1 - r0, cl_p, _off(part, p0), np_c[ci], st),
                  "mpb_bn_bwd_stats_bf16")
            p0 += np_c[ci]
        d_a = None
        for l in range(L - 1, -1, -1):
            cout, cin, cout_p, cin_p = dims[l]
            z, sc, gamma = zs[l], stats[l], gammas[l]
            pooled = l == L - 1
            coef = torch.empty(3, cout_p, dtype=torch.float32, device=dev)
            dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
            dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
            # weight-gradient accumulator + the (exactly zero) conv-bias gradient: one buffer, zero-filled by the finalize launch
            nw = cout_p * cin_p
            wbuf = torch.empty(nw + (cout + 3) // 4 * 4, dtype=torch.float32, device=dev)
            dw, dbias = wbuf[:nw].view(cout_p, cin_p), wbuf[nw:nw + cout]
            check(lib.mpb_bn_bwd_finalize_f32(ptr(part), sum(np_c), cout_p, cout, M, ptr(gamma), ptr(sc[2]), ptr(sc[3]), ptr(dgamma),
                                              ptr(dbeta), ptr(coef), ptr(wbuf), wbuf.numel(), st), "mpb_bn_bwd_finalize_f32")
            if l == 0 and narrow is not None:
                # fused dZ + weight gradient against the re-gathered rows; nothing upstream of the grouping needs a gradient
                check(lib.mpb_sa_first_layer_bwd_bf16(ptr(d_a), ptr(z), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), ptr(coef),
                                                      *_narrow_args(narrow), cout_p, ptr(dw), cin_p, st), "mpb_sa_first_layer_bwd_bf16")
                grads[0] = _unpermute_wgrad(dw, cout, cin, ctx.xyz_last)
                grads[1], grads[2], grads[3] = dbias, dgamma, dbeta
                d_a = None
                break
            dz = torch.empty(M, cout_p, dtype=torch.bfloat16, device=dev)
            need_da = l > 0 or ctx.needs_input_grad[0]
            d_prev = torch.empty(M, cin_p, dtype=torch.bfloat16, device=dev) if need_da else None
            if l > 0:
                np_n, part_n = alloc_partials(l - 1, False)
                ps = stats[l - 1]
            p0 = 0
            for ci, (r0, r1) in enumerate(chunks):   # apply -> wgrad -> dgrad -> next layer's statistics, chunk by chunk (L2)
                rows = r1 - r0
                if pooled:
                    check(lib.mpb_bn_bwd_apply_bf16(None, _off(d_pool, r0 // K), _off(argmax, r0 // K), K, _off(z, r0), ptr(sc[0]),
                                                    ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), ptr(coef), rows, cout_p, _off(dz, r0), st),
                          "mpb_bn_bwd_apply_bf16")
                else:
                    check(lib.mpb_bn_bwd_apply_bf16(_off(d_a, r0), None, None, K, _off(z, r0), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]),
                                                    ptr(sc[3]), ptr(coef), rows, cout_p, _off(dz, r0), st), "mpb_bn_bwd_apply_bf16")
                _gemm_wgrad(lib, _off(dz, r0), _off(acts[l], r0), ptr(dw), rows, cout_p, cin_p, st)
                if need_da:
                    _gemm_tn(lib, _off(dz, r0), ptr(wts[l]), _off(d_prev, r0), rows, cin_p, cout_p, st)
                if l > 0:
                    check(lib.mpb_bn_bwd_stats_bf16(_off(d_prev, r0), None, None, None, K, _off(zs[l - 1], r0), ptr(ps[0]), ptr(ps[1]),
                                                    ptr(ps[2]), ptr(ps[3]), rows, cin_p, _off(part_n, p0), np_n[ci], st),
                          "mpb_bn_bwd_stats_bf16")
                    p0 += np_n[ci]
            grads[6 * l] = _unpermute_wgrad(dw, cout, cin, ctx.xyz_last and l == 0)
            grads[6 * l + 1] = dbias                                                   # exact zeros: BN removes the conv bias
            grads[6 * l + 2] = dgamma
            grads[6 * l + 3] = dbeta
            d_a = d_prev
            if l > 0:
                np_c, part = np_n, part_n
        return (d_a, None, None, None, None, *grads)


def shared_mlp_max(a0, K, convs, bns, training, xyz_last=False):
    """Run the stack on bf16 rows `a0` [G*K, pad64(Cin)] -- or on a NarrowRows tuple (xyz [B,N,3] fp32, feats [B,N,D]
    fp32 | None, new_xyz [B,S,3] contiguous, idx [B,S,K] int64 contiguous), see SharedMLPMax; returns pooled fp32
    [G, C_last].  xyz_last=True when the rows are in mpb_group_points_bf16 order (features first, centred xyz last)."""
    flat, me = [], []
    for conv, bn in zip(convs, bns):
        flat += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        me.append((bn.momentum if bn.momentum is not None else 0.1, bn.eps))
    out = SharedMLPMax.apply(a0, K, training, tuple(me), bool(xyz_last), *flat)
    if training:
        for bn in bns:
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
    return out
