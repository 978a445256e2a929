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
"""
import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr


def pad64(c):
    return (c + 63) // 64 * 64


def _padded_weight(conv_weight, cout_p, cin_p, xyz_last):
    """[Cout,Cin,1,1] fp32 -> bf16 [cout_p, cin_p] (zero padded) and its transpose [cin_p, cout_p].
    xyz_last: the rows come from mpb_group_points_bf16 (features first, the 3 centred coordinates last),
    so the reference's xyz-first input channels (:137) move to the end."""
    cout, cin = conv_weight.shape[0], conv_weight.shape[1]
    w = torch.zeros(cout_p, cin_p, dtype=torch.bfloat16, device=conv_weight.device)
    src = conv_weight.detach().reshape(cout, cin)
    if xyz_last and cin > 3:
        w[:cout, :cin - 3] = src[:, 3:]
        w[:cout, cin - 3:cin] = src[:, :3]
    else:
        w[:cout, :cin] = src
    return w, w.t().contiguous()


def _unpermute_wgrad(dw, cout, cin, xyz_last):
    """Inverse of the column order used by _padded_weight, cropped to the real [Cout, Cin]."""
    if xyz_last and cin > 3:
        return torch.cat([dw[:cout, cin - 3:cin], dw[:cout, :cin - 3]], dim=1).reshape(cout, cin, 1, 1)
    return dw[:cout, :cin].reshape(cout, cin, 1, 1)


class SharedMLPMax(torch.autograd.Function):
    """pooled[G, C_L] = max_k relu(bn_L(... relu(bn_1(a0 @ W_1^T)) ...)) over the K rows of each group.

    apply(a0, K, training, momentum_eps, xyz_last, *flat) with
      a0    bf16 [M, pad64(Cin)], M = G*K
      flat  per layer: conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var
      momentum_eps  tuple of (momentum, eps) per layer
    """

    @staticmethod
    def forward(ctx, a0, K, training, momentum_eps, xyz_last, *flat):
        lib = _cabi.load()
        L = len(flat) // 6
        M = a0.shape[0]
        G = M // K
        dev = a0.device
        st = stream_ptr()
        acts, zs, stats, wts, dims = [a0], [], [], [], []
        a = a0
        out = argmax = None
        for l in range(L):
            W, bias, gamma, beta, rmean, rvar = flat[6 * l:6 * l + 6]
            cout, cin = W.shape[0], W.shape[1]
            cout_p, cin_p = pad64(cout), a.shape[1]
            w, wt = _padded_weight(W, cout_p, cin_p, xyz_last and l == 0)
            z = torch.empty(M, cout_p, dtype=torch.bfloat16, device=dev)
            check(lib.mpb_gemm_bf16_tn(ptr(a), ptr(w), ptr(z), M, cout_p, cin_p, 0, st), "mpb_gemm_bf16_tn")
            sc = torch.empty(4, cout_p, dtype=torch.float32, device=dev)      # rows: scale, shift, mean, rstd
            mom, eps = momentum_eps[l]
            if training:
                nparts = lib.mpb_bn_stat_partials(M, cout_p)
                part = torch.empty(nparts, 2, cout_p, dtype=torch.float32, device=dev)
                check(lib.mpb_bn_colstats_bf16(ptr(z), M, cout_p, ptr(part), nparts, st), "mpb_bn_colstats_bf16")
                check(lib.mpb_bn_finalize_f32(ptr(part), nparts, cout_p, cout, M, ptr(bias), ptr(gamma), ptr(beta), ptr(rmean),
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
            if l < L - 1:
                a = torch.empty(M, cout_p, dtype=torch.bfloat16, device=dev)
                check(lib.mpb_bn_relu_bf16(ptr(z), ptr(sc[0]), ptr(sc[1]), M, cout_p, ptr(a), st), "mpb_bn_relu_bf16")
                acts.append(a)
            else:
                out = torch.empty(G, cout_p, dtype=torch.float32, device=dev)
                argmax = torch.empty(G, cout_p, dtype=torch.int32, device=dev)
                check(lib.mpb_bn_relu_max_bf16(ptr(z), ptr(sc[0]), ptr(sc[1]), G, K, cout_p, ptr(out), ptr(argmax), st),
                      "mpb_bn_relu_max_bf16")
            zs.append(z)
            stats.append(sc)
            wts.append(wt)
            dims.append((cout, cin, cout_p, cin_p))
        ctx.K, ctx.L, ctx.dims, ctx.training, ctx.xyz_last = K, L, dims, training, xyz_last
        ctx.save_for_backward(argmax, *acts, *zs, *stats, *wts, *[flat[6 * l + 2] for l in range(L)])
        c_last = dims[-1][0]
        return out[:, :c_last] if c_last != out.shape[1] else out

    @staticmethod
    def backward(ctx, d_out):
        if not ctx.training:
            raise RuntimeError("SharedMLPMax: backward through eval-mode BatchNorm is not implemented on the tensor-core path; "
                               "use precision='fp32' for that")
        lib = _cabi.load()
        K, L, dims = ctx.K, ctx.L, ctx.dims
        saved = ctx.saved_tensors
        argmax = saved[0]
        acts, zs = saved[1:1 + L], saved[1 + L:1 + 2 * L]
        stats, wts, gammas = saved[1 + 2 * L:1 + 3 * L], saved[1 + 3 * L:1 + 4 * L], saved[1 + 4 * L:1 + 5 * L]
        M = acts[0].shape[0]
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
        d_a = None
        for l in range(L - 1, -1, -1):
            cout, cin, cout_p, cin_p = dims[l]
            z, sc, gamma = zs[l], stats[l], gammas[l]
            pooled = l == L - 1
            rows = G if pooled else M
            nparts = lib.mpb_bn_stat_partials(rows, cout_p)
            part = torch.empty(nparts, 2, cout_p, dtype=torch.float32, device=dev)
            coef = torch.empty(3, cout_p, dtype=torch.float32, device=dev)
            dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
            dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
            up_dense = None if pooled else ptr(d_a)
            up_pool = ptr(d_pool) if pooled else None
            am = ptr(argmax) if pooled else None
            check(lib.mpb_bn_bwd_stats_bf16(up_dense, up_pool, am, K, ptr(z), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), M, cout_p,
                                            ptr(part), nparts, st), "mpb_bn_bwd_stats_bf16")
            check(lib.mpb_bn_bwd_finalize_f32(ptr(part), nparts, cout_p, cout, M, ptr(gamma), ptr(sc[3]), ptr(dgamma), ptr(dbeta),
                                              ptr(coef), st), "mpb_bn_bwd_finalize_f32")
            dz = torch.empty(M, cout_p, dtype=torch.bfloat16, device=dev)
            check(lib.mpb_bn_bwd_apply_bf16(up_dense, up_pool, am, K, ptr(z), ptr(sc[0]), ptr(sc[1]), ptr(sc[2]), ptr(sc[3]), ptr(coef),
                                            M, cout_p, ptr(dz), st), "mpb_bn_bwd_apply_bf16")
            dw = torch.zeros(cout_p, cin_p, dtype=torch.float32, device=dev)
            check(lib.mpb_gemm_bf16_wgrad(ptr(dz), ptr(acts[l]), ptr(dw), M, cout_p, cin_p, st), "mpb_gemm_bf16_wgrad")
            grads[6 * l] = _unpermute_wgrad(dw, cout, cin, ctx.xyz_last and l == 0)
            grads[6 * l + 1] = torch.zeros(cout, dtype=torch.float32, device=dev)      # exact: BN removes the conv bias
            grads[6 * l + 2] = dgamma
            grads[6 * l + 3] = dbeta
            if l > 0 or ctx.needs_input_grad[0]:
                d_a = torch.empty(M, cin_p, dtype=torch.bfloat16, device=dev)
                check(lib.mpb_gemm_bf16_tn(ptr(dz), ptr(wts[l]), ptr(d_a), M, cin_p, cout_p, 0, st), "mpb_gemm_bf16_tn")
            else:
                d_a = None
        return (d_a, None, None, None, None, *grads)


def shared_mlp_max(a0, K, convs, bns, training, xyz_last=False):
    """Run the stack on bf16 rows `a0` [G*K, pad64(Cin)]; returns pooled fp32 [G, C_last].
    xyz_last=True when `a0` comes from mpb_group_points_bf16 (features first, centred xyz last)."""
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
