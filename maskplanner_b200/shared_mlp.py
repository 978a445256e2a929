"""Tensor-core shared MLP of PointNetSetAbstraction: relu(bn(conv1x1)) x L, then max over neighbours.

Reference: models/pointnet2_utils.py:208-214.  Host-side orchestration only -- every stage is a
hand-written sm_100a kernel behind the C ABI (include/maskplanner_b200.h, section a7):

    forward, layer l     Z_l = f_{l-1}(Z_{l-1}) @ W_l^T     mpb_sa_gemm_tn   tcgen05 + TMA; f = the previous layer's
                                                            BatchNorm + ReLU applied to the A tile in shared memory
                                                            (a_scale / a_shift), batch statistics of Z_l in the epilogue
                         scale/shift                        mpb_bn_finalize_f32 (running stats updated)
                         pooled = max_k f_L(Z_L)            mpb_bn_relu_max (fused max-pool + arg-max)
    backward, layer l    sum dY, sum dY*z                   epilogue of the dgrad GEMM of layer l+1 (epi 2); the last
                                                            layer (pooled upstream gradient): mpb_bn_bwd_stats
                         dgamma, dbeta, coefficients        mpb_bn_bwd_finalize_f32
                         dZ_l                               mpb_bn_bwd_apply
                         dW_l = dZ_l^T @ f_{l-1}(Z_{l-1})   mpb_sa_gemm_wgrad (MN-major operands, same on-the-fly f,
                                                            fixed-order split-M reduction: deterministic)
                         dA_{l-1} = dZ_l @ W_l              mpb_sa_gemm_tn

Only the pre-activations Z_l are ever stored: the normalised activations relu(bn(Z_l)) exist in shared memory only.
Three arithmetic modes (`precision`):
    "bf16"  bf16 activations and operands, fp32 accumulation / statistics              tolerance rel 1e-2
    "tf32"  fp32 activations, one TF32 pass (what cuDNN does for the reference on GPU)  tolerance ~1e-3
    "fp32"  fp32 activations, 3xTF32 (hi/lo split of both operands)                    tolerance rel 1e-4
Channel counts are zero-padded to multiples of 64.  In training mode the conv bias cannot influence the output
(BatchNorm removes it); it only enters the running mean, and its gradient is exactly zero (the reference's value
there is rounding noise).
"""
import ctypes
import os

import torch

from . import _cabi, gradsink
from ._cabi import check, ptr, stream_ptr

# precision -> (GEMM dtype code of mpb_sa_gemm_*, activation dtype code of mpb_bn_*, torch storage dtype)
MODES = {"bf16": (0, 0, torch.bfloat16), "tf32": (1, 1, torch.float32), "fp32": (2, 1, torch.float32)}


def pad64(c):
    return (c + 63) // 64 * 64


# Optional per-launch timing of the two GEMM kernels (bench.py's roofline leg): set to a list and every GEMM
# launch appends (kernel name, algorithmic bytes, flops, start event, end event), recorded on the launching stream.
# Algorithmic bytes count REAL channels only (zero-pad columns are not work the reference has).
GEMM_TIMELINE = None

FUSE_STATS = os.environ.get("MPB_FUSE_STATS", "1") == "1"        # forward BatchNorm statistics in the GEMM epilogue
FUSE_APPLY = os.environ.get("MPB_FUSE_APPLY", "1") == "1"        # BatchNorm + ReLU of the previous layer in the operand path
FUSE_BWD_STATS = os.environ.get("MPB_FUSE_BWD_STATS", "1") == "1"  # BatchNorm-backward statistics in the dgrad epilogue
# Pooled layer: dZ rebuilt in the two consumer GEMMs' operand path (mpb_sa_gemm_tn_pool / _wgrad_pool) instead of written by
# mpb_bn_bwd_apply and read back twice.  Correct and tested, saves 1.07 GB of HBM traffic per step at B = 64 -- but OFF by
# default: the dgrad GEMM with the BatchNorm-backward statistics epilogue is bound by shared-memory bandwidth, not HBM, and the
# extra operand pass costs about as much (stand-alone at M = 1M: dgrad 92 -> 161 us, wgrad 89 -> 108 us) as the removed
# 105 us kernel saves (profiles/r02_notes.md has the three versions that were measured).
FUSE_POOL_APPLY = os.environ.get("MPB_FUSE_POOL_APPLY", "0") == "1"
LEAN_POOLED_APPLY = os.environ.get("MPB_LEAN_POOLED_APPLY", "1") == "1"   # pooled dZ from pgo / (-w, e) (mpb_bn_bwd_apply_pooled)


class _Timed:
    def __init__(self, name, nbytes, flops, tag=""):
        self.rec = None
        if GEMM_TIMELINE is not None:
            self.rec = (name, nbytes, flops, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), tag)

    def __enter__(self):
        if self.rec:
            self.rec[3].record()

    def __exit__(self, *exc):
        if self.rec:
            self.rec[4].record()
            GEMM_TIMELINE.append(self.rec)
        return False


def _gemm_tn(lib, gd, esz, a, b, b_lo, c, M, N, K, st, real, a_affine=None, epi=0, partials=None, nparts=0, z=None, z_affine=None, pool=None):
    """C[M,N] = f(A)[M,K] @ B[N,K]^T.  real = (real K channels, real N channels) for the algorithmic byte count:
    A and B read once, C written once (+ Z read once for epi 2).  pool = (group size, argmax, pgo, negw_e): A is the stored
    pre-activation of the max-pooled layer and the kernel rebuilds dZ from it (mpb_sa_gemm_tn_pool)."""
    rk, rn = real
    nbytes = esz * (M * rk + rn * rk + M * rn + (M * rn if epi == 2 else 0))
    xf = 2 if pool else (1 if a_affine else 0)
    with _Timed("gemm_tn_kernel", nbytes, 2 * M * rn * rk, "M=%d N=%d K=%d xform=%d epi=%d" % (M, N, K, xf, epi)):
        if pool:
            pk, p_arg, p_pgo, p_we = pool
            check(lib.mpb_sa_gemm_tn_pool(gd, ptr(a), ptr(b), ptr(c), M, N, K, pk, ptr(p_arg), ptr(p_pgo), ptr(p_we), epi, ptr(partials), nparts,
                                          ptr(z), ptr(z_affine[0]) if z_affine else None, ptr(z_affine[1]) if z_affine else None, st),
                  "mpb_sa_gemm_tn_pool")
            return
        check(lib.mpb_sa_gemm_tn(gd, ptr(a), ptr(b), ptr(b_lo), ptr(c), M, N, K,
                                 ptr(a_affine[0]) if a_affine else None, ptr(a_affine[1]) if a_affine else None,
                                 epi, ptr(partials), nparts, ptr(z), ptr(z_affine[0]) if z_affine else None,
                                 ptr(z_affine[1]) if z_affine else None, st), "mpb_sa_gemm_tn")


DEFER_WGRAD_REDUCE = os.environ.get("MPB_DEFER_WGRAD_REDUCE", "1") == "1"
_PENDING_REDUCE = []


def join_wgrad_reductions():
    """Make the current stream wait for the weight-gradient reductions issued on the side stream."""
    while _PENDING_REDUCE:
        fork, dw = _PENDING_REDUCE.pop()
        fork.join(dw)


def _gemm_wgrad(lib, gd, esz, dz, a, M, N, K, st, real, a_affine, cout, cin, xyz_last, dw, pool=None):
    """dW[cout,cin] = crop(dZ[M,N]^T @ f(A)[M,K]).  Algorithmic bytes: dZ and A read once, dW written once (fp32).
    The fixed-order sum of the per-split partial tiles only feeds the optimizer: it is issued on a side stream (joined at the
    end of the module's backward), so the dgrad GEMM that follows starts as soon as the partial tiles are written."""
    from .streams import Fork
    rk, rn = real
    xf = 1 if a_affine else 0
    ws_bytes = lib.mpb_sa_gemm_wgrad_workspace(gd, M, N, K, xf)
    if ws_bytes < 0:
        raise _cabi.MpbError("mpb_sa_gemm_wgrad: unsupported shape M=%d N=%d K=%d" % (M, N, K))
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dw.device)
    defer = DEFER_WGRAD_REDUCE
    with _Timed("wgrad_kernel", esz * M * (rn + rk) + 4 * rn * rk, 2 * M * rn * rk, "M=%d N=%d K=%d xform=%d pool=%d" % (M, N, K, xf, 1 if pool else 0)):
        if pool:
            pk, p_arg, p_pgo, p_we = pool
            check(lib.mpb_sa_gemm_wgrad_pool(gd, ptr(dz), ptr(a), M, N, K, ptr(a_affine[0]) if a_affine else None,
                                             ptr(a_affine[1]) if a_affine else None, pk, ptr(p_arg), ptr(p_pgo), ptr(p_we), ptr(ws), cout, cin,
                                             1 if xyz_last else 0, -1 if defer else 0, ptr(dw), st), "mpb_sa_gemm_wgrad_pool", launches=1 if defer else 2)
        else:
            check(lib.mpb_sa_gemm_wgrad(gd, ptr(dz), ptr(a), M, N, K, ptr(a_affine[0]) if a_affine else None,
                                        ptr(a_affine[1]) if a_affine else None, ptr(ws), cout, cin, 1 if xyz_last else 0, -1 if defer else 0, ptr(dw), st),
                  "mpb_sa_gemm_wgrad", launches=1 if defer else 2)
    if defer:
        with Fork(ws, dw, slot=4) as fork:
            check(lib.mpb_sa_gemm_wgrad_reduce(gd, M, N, K, xf, ptr(ws), cout, cin, 1 if xyz_last else 0, 0, ptr(dw), stream_ptr()),
                  "mpb_sa_gemm_wgrad_reduce")
        if fork.active:
            _PENDING_REDUCE.append((fork, dw))
        

def _f32_source(conv_weight):
    src = conv_weight.detach()
    if src.dtype != torch.float32 or not src.is_contiguous():
        src = src.float().contiguous()
    return src


def _alloc_packed(kind, cout_p, cin_p, mode, dev):
    """Operand buffers of one layer: kind "gemm" -> (w, w_lo, wt, wt_lo) [cout_p, cin_p] / [cin_p, cout_p]; kind "narrow" ->
    (w [cout_p, 8], scratch transpose)."""
    if mode == "bf16":
        return (torch.empty(cout_p, cin_p, dtype=torch.bfloat16, device=dev), None,
                torch.empty(cin_p, cout_p, dtype=torch.bfloat16, device=dev), None)
    if kind == "narrow":
        buf = torch.empty(2, cout_p * cin_p, dtype=torch.float32, device=dev)
        return buf[0].view(cout_p, cin_p), None, buf[1].view(cin_p, cout_p), None
    buf = torch.empty(4, cout_p * cin_p, dtype=torch.float32, device=dev)
    return buf[0].view(cout_p, cin_p), buf[1].view(cout_p, cin_p), buf[2].view(cin_p, cout_p), buf[3].view(cin_p, cout_p)


def _fill_packed(bufs, kind, conv_weight, cout_p, cin_p, xyz_last, mode):
    """One launch: [Cout,Cin,1,1] fp32 -> the zero-padded operand and its transpose in `bufs` (on the current stream)."""
    cout, cin = conv_weight.shape[0], conv_weight.shape[1]
    src = _f32_source(conv_weight)
    w, w_lo, wt, wt_lo = bufs
    lib = _cabi.load()
    if mode == "bf16":
        check(lib.mpb_pack_weight_bf16(ptr(src), cout, cin, cout_p, cin_p, 1 if xyz_last else 0, ptr(w), ptr(wt), stream_ptr()),
              "mpb_pack_weight_bf16")
    elif kind == "narrow":      # UNSPLIT: the first-layer kernel's CUDA-core FMAs are exact fp32
        check(lib.mpb_pack_weight_tf32(ptr(src), cout, cin, cout_p, cin_p, 1, ptr(w), None, ptr(wt), None, stream_ptr()),
              "mpb_pack_weight_tf32")
    else:
        check(lib.mpb_pack_weight_tf32(ptr(src), cout, cin, cout_p, cin_p, 1 if xyz_last else 0, ptr(w), ptr(w_lo), ptr(wt), ptr(wt_lo),
                                       stream_ptr()), "mpb_pack_weight_tf32")


class PackAhead:
    """The weight operands of a whole step, packed beside the first layer instead of in front of every GEMM.

    Packing a layer's weight ([Cout,Cin,1,1] fp32 -> padded bf16 / split-TF32 operand + transpose) is a 2 us launch, but it
    sat on the main stream between the BatchNorm finalize of layer l-1 and the GEMM of layer l: nine serial launches on the
    forward critical path.  The weights are known when the step starts, so a Trainer brackets its step with begin() / end():
    the first bracketed step records which operands the layers ask for (in order), later steps issue the first one in line
    and all the others on a side stream (a parallel branch of the captured graph) into persistent buffers; the layers pick
    them up (`take`), the first pick-up after the first layer joins the branch -- by then it finished long ago.  Outside a
    bracket, or for an operand that was not announced, the layer packs in line exactly as before.  MPB_PREPACK=0 disables."""

    enabled = os.environ.get("MPB_PREPACK", "1") == "1"

    def __init__(self):
        self.requests = None          # [(key, kind, weight tensor, cout_p, cin_p, xyz_last, mode)]
        self.bufs = {}
        self.ready = {}
        self.fork = None
        self.log = None

    def begin(self):
        global _PACK_AHEAD
        self.ready, self.fork, self.log = {}, None, None
        if not self.enabled:
            return
        _PACK_AHEAD = self
        if self.requests is None:
            self.log = []
            return
        from .streams import Fork

        def run(req):
            key, kind, W, cout_p, cin_p, xyz_last, mode = req
            bufs = self.bufs.get(key)
            if bufs is None:
                bufs = self.bufs[key] = _alloc_packed(kind, cout_p, cin_p, mode, W.device)
            _fill_packed(bufs, kind, W, cout_p, cin_p, xyz_last, mode)
            self.ready[key] = bufs

        if self.requests:
            run(self.requests[0])
        if len(self.requests) > 1:
            with Fork(self.requests[1][2], slot=5) as fork:
                for req in self.requests[1:]:
                    run(req)
            self.fork = fork if fork.active else None
            self._first = self.requests[0][0]

    def take(self, key, kind, W, cout_p, cin_p, xyz_last, mode):
        if self.log is not None:
            self.log.append((key, kind, W, cout_p, cin_p, xyz_last, mode))
            return None
        hit = self.ready.pop(key, None)
        if hit is not None and self.fork is not None and key != self._first:
            self.fork.join()
            self.fork = None
        return hit

    def end(self):
        global _PACK_AHEAD
        if self.log is not None:
            self.requests, self.log = self.log, None
        if self.fork is not None:
            self.fork.join()
            self.fork = None
        self.ready = {}
        if _PACK_AHEAD is self:
            _PACK_AHEAD = None


_PACK_AHEAD = None


def _packed(kind, conv_weight, cout_p, cin_p, xyz_last, mode):
    if _PACK_AHEAD is not None:
        key = (kind, conv_weight.data_ptr(), tuple(conv_weight.shape), cout_p, cin_p, bool(xyz_last), mode)
        hit = _PACK_AHEAD.take(key, kind, conv_weight, cout_p, cin_p, bool(xyz_last), mode)
        if hit is not None:
            return hit
    bufs = _alloc_packed(kind, cout_p, cin_p, mode, conv_weight.device)
    _fill_packed(bufs, kind, conv_weight, cout_p, cin_p, xyz_last, mode)
    return bufs


def _padded_weight(conv_weight, cout_p, cin_p, xyz_last, mode):
    """[Cout,Cin,1,1] fp32 -> the GEMM's B operand [cout_p, cin_p] (zero padded) and its transpose [cin_p, cout_p].
    bf16: (w, None, wt, None).  tf32 / fp32: hi = tf32(w) and lo = w - hi (lo is None for the single-pass mode).
    xyz_last: the rows come from mpb_group_points_bf16 (features first, the 3 centred coordinates last), so the
    reference's xyz-first input channels (:137) move to the end."""
    return _packed("gemm", conv_weight, cout_p, cin_p, xyz_last, mode)


def _narrow_weight(conv_weight, cout_p, mode):
    """First-layer weight for the on-the-fly kernel: [cout_p, 8] in the activation storage type, xyz-last column order,
    UNSPLIT in the fp32 modes (the kernel's CUDA-core FMAs are exact fp32)."""
    return _packed("narrow", conv_weight, cout_p, NARROW_LDW, True, mode)[0]


NARROW_LDW = 8          # leading dimension of the packed first-layer weight on the narrow path (3 + D <= 8 channels)


def narrow_rows_supported(points, K):
    """The on-the-fly first layer applies when the grouped row has at most 8 channels and its features (if any) need
    no gradient."""
    D = 0 if points is None else points.shape[2]
    return 3 + D <= NARROW_LDW and not (points is not None and points.requires_grad) \
        and os.environ.get("MPB_NARROW_FIRST", "1") == "1"


def _narrow_args(narrow):
    """(xyz, feats|None, new_xyz, idx) -> the leading arguments of mpb_sa_first_layer[_bwd]."""
    xyz, feats, new_xyz, idx = narrow
    B, N, _ = xyz.shape
    _, S, K = idx.shape
    fs = tuple(feats.stride()) if feats is not None else (0, 0, 0)
    D = 0 if feats is None else feats.shape[2]
    return (ptr(xyz), *xyz.stride(), ptr(feats), *fs, ptr(new_xyz), ptr(idx), B, N, S, K, D)


def _unpermute_narrow_wgrad(dw, cout, cin):
    """[cout_p, 8] accumulator of the narrow backward (xyz-last columns) -> the reference's [Cout, Cin, 1, 1]."""
    if cin > 3:
        return torch.cat([dw[:cout, cin - 3:cin], dw[:cout, :cin - 3]], dim=1).reshape(cout, cin, 1, 1)
    return dw[:cout, :cin].reshape(cout, cin, 1, 1)


class SharedMLPMax(torch.autograd.Function):
    """pooled[G, C_L] = max_k relu(bn_L(... relu(bn_1(a0 @ W_1^T)) ...)) over the K rows of each group.

    apply(a0, K, training, momentum_eps, xyz_last, mode, *flat) with
      a0    [M, pad64(Cin)] rows in the mode's storage type (bf16 / fp32), M = G*K -- or a NarrowRows tuple
            (xyz, feats|None, new_xyz, idx) when the grouped row has at most 8 channels: the first layer then gathers
            its input on the fly (mpb_sa_first_layer) and the [M, 64] operand is never written
      flat  per layer: conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var
      momentum_eps  tuple of (momentum, eps) per layer
    """

    @staticmethod
    def forward(ctx, a0, K, training, momentum_eps, xyz_last, mode, counters, *flat):
        lib = _cabi.load()
        gd, ad, tdt = MODES[mode]
        esz = 2 if tdt == torch.bfloat16 else 4
        L = len(flat) // 6
        narrow = a0 if isinstance(a0, tuple) else None
        if narrow is not None:
            n_xyz, n_feats, n_new, n_idx = narrow
            M = n_idx.numel()
            dev = n_xyz.device
            c_in_p = NARROW_LDW
            a0 = None
        else:
            if a0.dtype != tdt:
                raise TypeError("SharedMLPMax(%s): rows must be %s, got %s" % (mode, tdt, a0.dtype))
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
        zs, stats, wts = [], [], []
        for l in range(L):
            W, bias, gamma, beta, rmean, rvar = flat[6 * l:6 * l + 6]
            cout, cin, cout_p, cin_p = dims[l]
            z = torch.empty(M, cout_p, dtype=tdt, device=dev)
            sc = torch.empty(4, cout_p, dtype=torch.float32, device=dev)      # rows: scale, shift, mean, rstd
            mom, eps = momentum_eps[l]
            part, nparts = None, 0
            if l == 0 and narrow is not None:
                w = _narrow_weight(W, cout_p, mode)
                nparts = lib.mpb_bn_stat_partials(M, cout_p)
                part = torch.empty(nparts, 2, cout_p, dtype=torch.float32, device=dev)
                check(lib.mpb_sa_first_layer(ad, *_narrow_args(narrow), ptr(w), NARROW_LDW, cout_p, ptr(z), ptr(part), nparts, st),
                      "mpb_sa_first_layer")
                wts.append((None, None))
            else:
                w, w_lo, wt, wt_lo = _padded_weight(W, cout_p, cin_p, xyz_last and l == 0, mode)
                wts.append((wt, wt_lo))
                a_in = a0 if l == 0 else zs[l - 1]
                affine = None
                if l > 0:
                    if FUSE_APPLY:
                        affine = (stats[l - 1][0], stats[l - 1][1])
                    else:   # A/B switch: materialise relu(bn(Z_{l-1})) like round 1 did
                        a_in = torch.empty(M, cin_p, dtype=tdt, device=dev)
                        check(lib.mpb_bn_relu(ad, ptr(zs[l - 1]), ptr(stats[l - 1][0]), ptr(stats[l - 1][1]), M, cin_p, ptr(a_in), st), "mpb_bn_relu")
                xf = 1 if affine else 0
                if training and FUSE_STATS:
                    nparts = lib.mpb_sa_gemm_stat_partials(gd, M, cout_p, cin_p, xf, 1)
                if nparts:
                    part = torch.empty(nparts, 2, cout_p, dtype=torch.float32, device=dev)
                real = (cin if l == 0 else dims[l - 1][0], cout)
                _gemm_tn(lib, gd, esz, a_in, w, w_lo, z, M, cout_p, cin_p, st, real, a_affine=affine, epi=1 if nparts else 0,
                         partials=part, nparts=nparts)
                if training and not nparts:
                    nparts = lib.mpb_bn_stat_partials(M, cout_p)
                    part = torch.empty(nparts, 2, cout_p, dtype=torch.float32, device=dev)
                    check(lib.mpb_bn_colstats(ad, ptr(z), M, cout_p, ptr(part), nparts, st), "mpb_bn_colstats")
            if training:
                check(lib.mpb_bn_finalize_f32(ptr(part), nparts, cout_p, cout, M, ptr(bias), ptr(gamma), ptr(beta), ptr(rmean),
                                              ptr(rvar), mom, eps, ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]),
                                              ptr(counters[l]) if counters is not None else None, st), "mpb_bn_finalize_f32")
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
        cl_p = dims[-1][2]
        out = torch.empty(G, cl_p, dtype=torch.float32, device=dev)
        argmax = torch.empty(G, cl_p, dtype=torch.int32, device=dev)
        zmax = torch.empty(G, cl_p, dtype=torch.float32, device=dev) if training else None
        sc = stats[-1]
        check(lib.mpb_bn_relu_max(ad, ptr(zs[-1]), ptr(sc[0]), ptr(sc[1]), G, K, cl_p, ptr(out), ptr(argmax), ptr(zmax), st),
              "mpb_bn_relu_max")
        ctx.K, ctx.L, ctx.dims, ctx.training, ctx.xyz_last, ctx.mode = K, L, dims, training, xyz_last, mode
        ctx.narrow = narrow is not None
        ctx.has_a0 = a0 is not None
        flat_w = []
        for wt, wt_lo in wts:
            flat_w += [wt, wt_lo]
        ctx.sinks = [tuple(gradsink.lookup(flat[6 * l + j]) for j in range(4)) for l in range(L)] if gradsink.active() else None
        ctx.save_for_backward(argmax, *zs, *stats, *flat_w, *[flat[6 * l + 2] for l in range(L)], zmax, a0,
                              *(narrow if narrow is not None else ()))
        c_last = dims[-1][0]
        return out[:, :c_last] if c_last != out.shape[1] else out

    @staticmethod
    def backward(ctx, d_out):
        if not ctx.training:
            raise RuntimeError("SharedMLPMax: backward through eval-mode BatchNorm is not implemented on the tensor-core path; "
                               "PointNetSetAbstraction falls back to stock torch ops for that (off the training path)")
        lib = _cabi.load()
        K, L, dims, mode = ctx.K, ctx.L, ctx.dims, ctx.mode
        gd, ad, tdt = MODES[mode]
        esz = 2 if tdt == torch.bfloat16 else 4
        saved = ctx.saved_tensors
        argmax = saved[0]
        zs, stats = saved[1:1 + L], saved[1 + L:1 + 2 * L]
        wts = [(saved[1 + 2 * L + 2 * l], saved[2 + 2 * L + 2 * l]) for l in range(L)]
        gammas = saved[1 + 4 * L:1 + 5 * L]
        zmax, a0 = saved[1 + 5 * L], saved[2 + 5 * L]
        narrow = tuple(saved[3 + 5 * L:7 + 5 * L]) if ctx.narrow else None
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
        # statistics of the last layer: the pooled upstream gradient touches one row per (group, channel)
        nparts = lib.mpb_bn_stat_partials(G, cl_p)
        part = torch.empty(nparts, 2 * cl_p, dtype=torch.float32, device=dev)
        sc = stats[L - 1]
        # pooled layer: the dZ tensor is not materialised -- its two consumer GEMMs rebuild it from Z_L in their operand path
        pool_fuse = (FUSE_POOL_APPLY and FUSE_APPLY and mode == "bf16" and L >= 2 and K >= 16 and (K in (16, 32, 64) or K % 128 == 0))
        lean_apply = LEAN_POOLED_APPLY and L >= 1 and not (L == 1 and narrow is not None)
        pgo = torch.empty(G, cl_p, dtype=torch.float32, device=dev) if (pool_fuse or lean_apply) else None
        check(lib.mpb_bn_bwd_stats(ad, None, ptr(d_pool), ptr(argmax), ptr(zmax), K, ptr(zs[L - 1]), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]),
                                   ptr(sc[3]), M, cl_p, ptr(part), nparts, ptr(pgo), st), "mpb_bn_bwd_stats")
        d_a = None
        for l in range(L - 1, -1, -1):
            cout, cin, cout_p, cin_p = dims[l]
            z, sc, gamma = zs[l], stats[l], gammas[l]
            pooled = l == L - 1
            is_narrow = l == 0 and narrow is not None
            coef = torch.empty(3, cout_p, dtype=torch.float32, device=dev)
            # data-parallel runs: the kernels write straight into the parameter's view of the flat gradient buffer and
            # autograd gets None for it (maskplanner_b200.gradsink); the conv bias view is never written -- its gradient
            # is exactly zero and the buffer starts zeroed
            sink_w, sink_b, sink_g, sink_be = ctx.sinks[l] if ctx.sinks else (None, None, None, None)
            dgamma = sink_g if sink_g is not None else torch.empty(cout, dtype=torch.float32, device=dev)
            dbeta = sink_be if sink_be is not None else torch.empty(cout, dtype=torch.float32, device=dev)
            # zero-filled by the finalize launch: the (exactly zero) conv-bias gradient and, for the narrow first layer,
            # the weight-gradient accumulator of its fused backward kernel
            nw = cout_p * NARROW_LDW if is_narrow else 0
            wbuf = torch.empty(nw + (cout + 3) // 4 * 4, dtype=torch.float32, device=dev)
            dbias = wbuf[nw:nw + cout]
            pool = None
            negw_e = torch.empty(2, cout_p, dtype=torch.float32, device=dev) if (pooled and (pool_fuse or lean_apply)) else None
            check(lib.mpb_bn_bwd_finalize_f32(ptr(part), nparts, cout_p, cout, M, ptr(gamma), ptr(sc[2]), ptr(sc[3]), ptr(dgamma),
                                              ptr(dbeta), ptr(coef), ptr(wbuf), wbuf.numel(), ptr(negw_e), st), "mpb_bn_bwd_finalize_f32")
            if negw_e is not None and pool_fuse:
                pool = (K, argmax, pgo, negw_e)
            if is_narrow:
                # fused dZ + weight gradient against the re-gathered rows; nothing upstream of the grouping needs a gradient
                dw = wbuf[:nw].view(cout_p, NARROW_LDW)
                check(lib.mpb_sa_first_layer_bwd(ad, ptr(d_a), ptr(z), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), ptr(coef),
                                                 *_narrow_args(narrow), cout_p, ptr(dw), NARROW_LDW, st), "mpb_sa_first_layer_bwd")
                g0 = _unpermute_narrow_wgrad(dw, cout, cin)
                if sink_w is not None:
                    sink_w.copy_(g0)
                grads[0] = g0 if sink_w is None else None
                grads[1] = dbias if sink_b is None else None
                grads[2] = dgamma if sink_g is None else None
                grads[3] = dbeta if sink_be is None else None
                d_a = None
                break
            if pool is not None:
                dz = z          # the GEMMs transform the stored pre-activation into dZ tile by tile
            elif pooled and negw_e is not None:
                dz = torch.empty(M, cout_p, dtype=tdt, device=dev)
                check(lib.mpb_bn_bwd_apply_pooled(ad, ptr(pgo), ptr(argmax), K, ptr(z), ptr(negw_e), M, cout_p, ptr(dz), st), "mpb_bn_bwd_apply_pooled")
            elif pooled:
                dz = torch.empty(M, cout_p, dtype=tdt, device=dev)
                check(lib.mpb_bn_bwd_apply(ad, None, ptr(d_pool), ptr(argmax), K, ptr(z), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]),
                                           ptr(coef), M, cout_p, ptr(dz), st), "mpb_bn_bwd_apply")
            else:
                dz = torch.empty(M, cout_p, dtype=tdt, device=dev)
                check(lib.mpb_bn_bwd_apply(ad, ptr(d_a), None, None, K, ptr(z), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), ptr(coef),
                                           M, cout_p, ptr(dz), st), "mpb_bn_bwd_apply")
            # weight gradient against the layer's input f_{l-1}(Z_{l-1}) (or the stored first-layer rows)
            ps = stats[l - 1] if l > 0 else None
            real = (cin if l == 0 else dims[l - 1][0], cout)
            dw = sink_w.view(cout, cin) if sink_w is not None else torch.empty(cout, cin, dtype=torch.float32, device=dev)
            if l > 0 and not FUSE_APPLY:
                a_prev = torch.empty(M, cin_p, dtype=tdt, device=dev)
                check(lib.mpb_bn_relu(ad, ptr(zs[l - 1]), ptr(ps[0]), ptr(ps[1]), M, cin_p, ptr(a_prev), st), "mpb_bn_relu")
                _gemm_wgrad(lib, gd, esz, dz, a_prev, M, cout_p, cin_p, st, real, None, cout, cin, False, dw)
            else:
                _gemm_wgrad(lib, gd, esz, dz, zs[l - 1] if l > 0 else a0, M, cout_p, cin_p, st, real, (ps[0], ps[1]) if l > 0 else None,
                            cout, cin, ctx.xyz_last and l == 0, dw, pool=pool)
            grads[6 * l] = dw.view(cout, cin, 1, 1) if sink_w is None else None
            grads[6 * l + 1] = dbias if sink_b is None else None                       # exact zeros: BN removes the conv bias
            grads[6 * l + 2] = dgamma if sink_g is None else None
            grads[6 * l + 3] = dbeta if sink_be is None else None
            need_da = l > 0 or (ctx.has_a0 and ctx.needs_input_grad[0])
            d_prev = None
            if need_da:
                d_prev = torch.empty(M, cin_p, dtype=tdt, device=dev)
                wt, wt_lo = wts[l]
                np_n = lib.mpb_sa_gemm_stat_partials(gd, M, cin_p, cout_p, 2 if pool else 0, 2) if (l > 0 and FUSE_BWD_STATS) else 0
                if np_n:   # the layer below's BatchNorm-backward statistics come out of this GEMM's epilogue
                    part_n = torch.empty(np_n, 2 * cin_p, dtype=torch.float32, device=dev)
                    _gemm_tn(lib, gd, esz, dz, wt, wt_lo, d_prev, M, cin_p, cout_p, st, (cout, real[0]), epi=2, partials=part_n, nparts=np_n,
                             z=zs[l - 1], z_affine=(ps[0], ps[1]), pool=pool)
                else:
                    _gemm_tn(lib, gd, esz, dz, wt, wt_lo, d_prev, M, cin_p, cout_p, st, (cout, real[0]), pool=pool)
                    if l > 0:
                        np_n = lib.mpb_bn_stat_partials(M, cin_p)
                        part_n = torch.empty(np_n, 2 * cin_p, dtype=torch.float32, device=dev)
                        check(lib.mpb_bn_bwd_stats(ad, ptr(d_prev), None, None, None, K, ptr(zs[l - 1]), ptr(ps[0]), ptr(ps[1]), ptr(ps[2]),
                                                   ptr(ps[3]), M, cin_p, ptr(part_n), np_n, None, st), "mpb_bn_bwd_stats")
                if l > 0:
                    part, nparts = part_n, np_n
            d_a = d_prev
        join_wgrad_reductions()
        return (d_a, None, None, None, None, None, None, *grads)


def shared_mlp_max(a0, K, convs, bns, training, xyz_last=False, mode="bf16"):
    """Run the stack on rows `a0` [G*K, pad64(Cin)] (bf16 for mode "bf16", fp32 for "tf32" / "fp32") -- or on a
    NarrowRows tuple (xyz [B,N,3] fp32, feats [B,N,D] fp32 | None, new_xyz [B,S,3] contiguous, idx [B,S,K] int64
    contiguous), see SharedMLPMax; returns pooled fp32 [G, C_last].  xyz_last=True when the rows are in
    mpb_group_points_bf16 order (features first, centred xyz last)."""
    if mode not in MODES:
        raise ValueError("mode must be one of %s" % (sorted(MODES),))
    flat, me = [], []
    for conv, bn in zip(convs, bns):
        flat += [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        me.append((bn.momentum if bn.momentum is not None else 0.1, bn.eps))
    # BatchNorm's num_batches_tracked counters are advanced by the finalize launches (device int64 scalars)
    counters = None
    if training:
        counters = tuple(bn.num_batches_tracked if (bn.num_batches_tracked is not None and bn.num_batches_tracked.is_cuda
                                                    and bn.num_batches_tracked.dtype == torch.int64) else None for bn in bns)
    out = SharedMLPMax.apply(a0, K, training, tuple(me), bool(xyz_last), mode, counters, *flat)
    if training:
        for bn, c in zip(bns, counters):
            if c is None and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
    return out
