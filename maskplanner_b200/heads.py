"""Regression heads of the MaskPlanner regressor as ONE autograd node over hand-written kernels.

Reference: models/pointnet2_cls_ssg.py:270-295 (modules) and :309-341 (forward) -- the pose head
(fc1/bn1, fc2/bn2, fc3, fc_normals) and the stroke-mask head (sm_fc1/sm_bn1, sm_fc2/sm_bn2, sm_fc3, mask_conf_out),
both fed by the encoder's global feature.  Parameters stay where the reference keeps them (the regressor's
``nn.Linear`` / ``nn.BatchNorm1d`` modules, so state_dicts interchange); this module only replaces the arithmetic:

* every GEMM runs on the tcgen05 kernels of csrc/sa_gemm.cu with the weight as the 128-row operand read straight from
  its fp32 parameter tensor (TF32 in the "bf16"/"tf32" encoder modes, 3xTF32 in "fp32"), activations feature-major
  [F, Bp] (csrc/heads.cu explains the layout);
* bias + BatchNorm1d + ReLU + dropout forward/backward are one row-local kernel each;
* weight gradients are written by the dW GEMM straight into the gradient tensors handed back to autograd.

The two heads only share the global feature, so the stroke-mask head runs on a side stream (fork/join with events; inside
CUDA-graph capture they become parallel branches), forward and backward.
Dropout: counter-based hash (seed from torch's CPU generator at construction, a device-side step counter advanced once per
forward), so a captured step draws fresh masks on every replay.  The masks are NOT torch's Philox stream: like the
reference on CUDA vs CPU, runs are reproducible per seed but not bit-identical across implementations.
"""
import torch

from . import _cabi, gradsink, streams
from ._cabi import check, ptr, stream_ptr

GEMM_DTYPE = {"bf16": 1, "tf32": 1, "fp32": 2}     # encoder precision mode -> head GEMM arithmetic (TF32 / 3xTF32)


def _pad32(b):
    return (b + 31) // 32 * 32


class _Ctx:
    """Per-forward scratch: kernels, stream pointer, arithmetic mode."""

    def __init__(self, B, dev, gd):
        self.lib = _cabi.load()
        self.B, self.Bp, self.dev, self.gd = B, _pad32(B), dev, gd
        self.split = gd == 2

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.dev)

    # Yt [Nout, Bp] = W [Nout, Kin] @ X [Bp, Kin]^T
    def linear_fwd(self, W, X, X_lo):
        Nout, Kin = W.shape
        Yt = self.empty(Nout, self.Bp)
        check(self.lib.mpb_sa_gemm_tn(self.gd, ptr(W), ptr(X), ptr(X_lo), ptr(Yt), Nout, self.Bp, Kin, None, None, 0, None, 0, None, None, None,
                                      stream_ptr()), "mpb_sa_gemm_tn(head fwd)")
        return Yt

    # dW [Nout, Kin] = dYt [Nout, Bp] @ Xt [Kin, Bp]^T
    def linear_dw(self, dYt, Xt, Xt_lo, Kin, out=None):
        Nout = dYt.shape[0]
        dW = out.view(Nout, Kin) if out is not None else self.empty(Nout, Kin)
        check(self.lib.mpb_sa_gemm_tn(self.gd, ptr(dYt), ptr(Xt), ptr(Xt_lo), ptr(dW), Nout, Kin, self.Bp, None, None, 0, None, 0, None, None, None,
                                      stream_ptr()), "mpb_sa_gemm_tn(head dW)")
        return dW

    # dXt [Kin, Bp] (+)= W^T @ dYt
    def linear_dx(self, W, dYt, out=None):
        Nout, Kin = W.shape
        ws_bytes = self.lib.mpb_sa_gemm_wgrad_workspace(self.gd, Nout, Kin, self.Bp, 0)
        if ws_bytes < 0:
            raise _cabi.MpbError("head dX: unsupported shape %s" % ((Nout, Kin, self.Bp),))
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=self.dev)
        acc = out is not None
        if out is None:
            out = self.empty(Kin, self.Bp)
        check(self.lib.mpb_sa_gemm_wgrad(self.gd, ptr(W), ptr(dYt), Nout, Kin, self.Bp, None, None, ptr(ws), Kin, self.Bp, 0, 1 if acc else 0,
                                         ptr(out), stream_ptr()), "mpb_sa_gemm_wgrad(head dX)", launches=2)
        return out


class HeadsFunction(torch.autograd.Function):
    """(traj_pred [B,P/λ... see regressor], masks [B, n_masks*out_vectors], scores [B, n_masks]) = heads(feat [B, 1024]).

    apply(feat, cfg, *params) with params in the order of PARAM_NAMES (weights/biases of the 8 linears, affine parameters
    and running statistics of the 4 BatchNorm1d modules) and cfg = dict(training, drop_p, seed, step (device int64),
    weight_orient, gemm_dtype, momentum (4), eps (4))."""

    LINEARS = ("fc1", "fc2", "fc3", "fc_normals", "sm_fc1", "sm_fc2", "sm_fc3", "mask_conf_out")
    BNS = ("bn1", "bn2", "sm_bn1", "sm_bn2")

    @staticmethod
    def forward(ctx, feat, cfg, *params):
        P = dict(zip(HeadsFunction.param_names(), params))
        B = feat.shape[0]
        dev = feat.device
        k = _Ctx(B, dev, cfg["gemm_dtype"])
        lib, Bp = k.lib, k.Bp
        training, drop_p = cfg["training"], cfg["drop_p"] if cfg["training"] else 0.0
        seed, step = cfg["seed"], cfg["step"]
        if training and drop_p > 0:
            check(lib.mpb_rng_advance(ptr(step), stream_ptr()), "mpb_rng_advance")
        feat = feat.contiguous().float()
        # the global feature in both layouts (batch-major rows padded to Bp for the GEMM's B operand)
        Kf = feat.shape[1]
        if k.split or Bp != B:
            X0 = torch.zeros(Bp, Kf, dtype=torch.float32, device=dev)
            X0[:B] = feat
        else:
            X0 = feat
        X0_lo = None
        if k.split:
            hi = _tf32_round(X0)
            X0_lo, X0 = X0 - hi, hi
        Ft, Ft_lo = k.empty(Kf, Bp), (k.empty(Kf, Bp) if k.split else None)
        check(lib.mpb_head_to_feature_major(ptr(feat), B, Bp, Kf, ptr(Ft), ptr(Ft_lo), None, stream_ptr()), "mpb_head_to_feature_major")

        saved = {}

        def act(name_fc, name_bn, layer, Yt):
            F = Yt.shape[0]
            Xt, X = k.empty(F, Bp), k.empty(Bp, F)
            Xt_lo, X_lo = (k.empty(F, Bp), k.empty(Bp, F)) if k.split else (None, None)
            mean, rstd = k.empty(F), k.empty(F)
            i = HeadsFunction.BNS.index(name_bn)
            check(lib.mpb_head_act_fwd(ptr(Yt), F, B, Bp, ptr(P[name_fc + ".bias"]), ptr(P[name_bn + ".weight"]), ptr(P[name_bn + ".bias"]),
                                       ptr(P[name_bn + ".running_mean"]), ptr(P[name_bn + ".running_var"]), cfg["momentum"][i], cfg["eps"][i],
                                       1 if training else 0, drop_p, seed, ptr(step), layer, ptr(Xt), ptr(Xt_lo), ptr(X), ptr(X_lo), ptr(mean), ptr(rstd),
                                       stream_ptr()), "mpb_head_act_fwd")
            saved[name_fc] = (Yt, mean, rstd)
            return X, X_lo, Xt, Xt_lo

        # stroke-mask head on the side stream (it only shares the global feature with the pose head)
        with streams.Fork(X0, Ft) as fork:
            m1 = act("sm_fc1", "sm_bn1", 2, k.linear_fwd(P["sm_fc1.weight"], X0, X0_lo))
            m2 = act("sm_fc2", "sm_bn2", 3, k.linear_fwd(P["sm_fc2.weight"], m1[0], m1[1]))
            Ys3 = k.linear_fwd(P["sm_fc3.weight"], m2[0], m2[1])
            Yc = k.linear_fwd(P["mask_conf_out.weight"], m2[0], m2[1])
            n3, nc = Ys3.shape[0], Yc.shape[0]
            masks = k.empty(B, n3)
            scores = k.empty(B, nc)
            check(lib.mpb_head_to_batch_major(ptr(Ys3), ptr(P["sm_fc3.bias"]), B, Bp, n3, ptr(masks), stream_ptr()), "mpb_head_to_batch_major")
            check(lib.mpb_head_to_batch_major(ptr(Yc), ptr(P["mask_conf_out.bias"]), B, Bp, nc, ptr(scores), stream_ptr()), "mpb_head_to_batch_major")
        h1 = act("fc1", "bn1", 0, k.linear_fwd(P["fc1.weight"], X0, X0_lo))
        h2 = act("fc2", "bn2", 1, k.linear_fwd(P["fc2.weight"], h1[0], h1[1]))
        Yt3 = k.linear_fwd(P["fc3.weight"], h2[0], h2[1])
        Ytn = k.linear_fwd(P["fc_normals.weight"], h2[0], h2[1])
        n_pose = Yt3.shape[0] // 3
        out = k.empty(B, n_pose, 6)
        check(lib.mpb_head_pose_out_fwd(ptr(Yt3), ptr(P["fc3.bias"]), ptr(Ytn), ptr(P["fc_normals.bias"]), B, Bp, n_pose, cfg["weight_orient"],
                                        ptr(out), stream_ptr()), "mpb_head_pose_out_fwd")
        fork.join(masks, scores)

        ctx.cfg, ctx.B = cfg, B
        ctx.names = HeadsFunction.param_names()
        ctx.sinks = {n: gradsink.lookup(t) for n, t in P.items()} if gradsink.active() else {}
        keep = [Ft, Ft_lo, h1[2], h1[3], h2[2], h2[3], m1[2], m1[3], m2[2], m2[3], Ytn]
        for n in ("fc1", "fc2", "sm_fc1", "sm_fc2"):
            keep += list(saved[n])
        ctx.save_for_backward(*keep, *params)
        ctx.n_keep = len(keep)
        return out, masks, scores

    @staticmethod
    def backward(ctx, d_out, d_masks, d_scores):
        cfg, B = ctx.cfg, ctx.B
        sv = ctx.saved_tensors
        keep, params = sv[:ctx.n_keep], sv[ctx.n_keep:]
        P = dict(zip(ctx.names, params))
        (Ft, Ft_lo, H1t, H1t_lo, H2t, H2t_lo, M1t, M1t_lo, M2t, M2t_lo, Ytn) = keep[:11]
        pre = {n: keep[11 + 3 * i:14 + 3 * i] for i, n in enumerate(("fc1", "fc2", "sm_fc1", "sm_fc2"))}
        dev = Ft.device
        k = _Ctx(B, dev, cfg["gemm_dtype"])
        lib, Bp = k.lib, k.Bp
        training, drop_p = cfg["training"], cfg["drop_p"] if cfg["training"] else 0.0
        seed, step = cfg["seed"], cfg["step"]
        G = {}
        S = ctx.sinks

        def vec(name, n):
            """gradient vector of a small parameter: its view of the flat buffer when one is registered"""
            t = S.get(name)
            return t if t is not None else k.empty(n)

        def act_bwd(name_fc, name_bn, layer, dXt):
            Yt, mean, rstd = pre[name_fc]
            F = Yt.shape[0]
            dYt = k.empty(F, Bp)
            dg, db, dbias = vec(name_bn + ".weight", F), vec(name_bn + ".bias", F), vec(name_fc + ".bias", F)
            check(lib.mpb_head_act_bwd(ptr(dXt), ptr(Yt), F, B, Bp, ptr(P[name_fc + ".bias"]), ptr(P[name_bn + ".weight"]), ptr(P[name_bn + ".bias"]),
                                       ptr(mean), ptr(rstd), 1 if training else 0, drop_p, seed, ptr(step), layer, ptr(dYt), ptr(dg), ptr(db), ptr(dbias),
                                       stream_ptr()), "mpb_head_act_bwd")
            G[name_bn + ".weight"], G[name_bn + ".bias"], G[name_fc + ".bias"] = dg, db, dbias
            return dYt

        def to_feature_major(dY, F, bias_name):
            dYt, dbias = k.empty(F, Bp), vec(bias_name, F)
            check(lib.mpb_head_to_feature_major(ptr(dY), B, Bp, F, ptr(dYt), None, ptr(dbias), stream_ptr()), "mpb_head_to_feature_major")
            return dYt, dbias

        Kf = Ft.shape[0]
        d_masks = d_masks.contiguous().float()
        d_scores = d_scores.contiguous().float()
        d_out = d_out.contiguous().float()
        # stroke-mask head backward on the side stream
        with streams.Fork(d_masks, d_scores, M2t, M1t, Ft) as fork:
            n3, nc = P["sm_fc3.weight"].shape[0], P["mask_conf_out.weight"].shape[0]
            dYs3, G["sm_fc3.bias"] = to_feature_major(d_masks, n3, "sm_fc3.bias")
            dYc, G["mask_conf_out.bias"] = to_feature_major(d_scores, nc, "mask_conf_out.bias")
            G["sm_fc3.weight"] = k.linear_dw(dYs3, M2t, M2t_lo, M2t.shape[0], out=S.get("sm_fc3.weight"))
            G["mask_conf_out.weight"] = k.linear_dw(dYc, M2t, M2t_lo, M2t.shape[0], out=S.get("mask_conf_out.weight"))
            dM2t = k.linear_dx(P["sm_fc3.weight"], dYs3)
            k.linear_dx(P["mask_conf_out.weight"], dYc, out=dM2t)
            dYs2 = act_bwd("sm_fc2", "sm_bn2", 3, dM2t)
            G["sm_fc2.weight"] = k.linear_dw(dYs2, M1t, M1t_lo, M1t.shape[0], out=S.get("sm_fc2.weight"))
            dM1t = k.linear_dx(P["sm_fc2.weight"], dYs2)
            dYs1 = act_bwd("sm_fc1", "sm_bn1", 2, dM1t)
            G["sm_fc1.weight"] = k.linear_dw(dYs1, Ft, Ft_lo, Kf, out=S.get("sm_fc1.weight"))
            dF_mask = k.linear_dx(P["sm_fc1.weight"], dYs1)
        n_pose = P["fc3.weight"].shape[0] // 3
        dYt3, dYtn = k.empty(3 * n_pose, Bp), k.empty(3 * n_pose, Bp)
        G["fc3.bias"], G["fc_normals.bias"] = vec("fc3.bias", 3 * n_pose), vec("fc_normals.bias", 3 * n_pose)
        check(lib.mpb_head_pose_out_bwd(ptr(d_out), ptr(Ytn), ptr(P["fc_normals.bias"]), B, Bp, n_pose, cfg["weight_orient"], ptr(dYt3), ptr(dYtn),
                                        ptr(G["fc3.bias"]), ptr(G["fc_normals.bias"]), stream_ptr()), "mpb_head_pose_out_bwd")
        G["fc3.weight"] = k.linear_dw(dYt3, H2t, H2t_lo, H2t.shape[0], out=S.get("fc3.weight"))
        G["fc_normals.weight"] = k.linear_dw(dYtn, H2t, H2t_lo, H2t.shape[0], out=S.get("fc_normals.weight"))
        dH2t = k.linear_dx(P["fc3.weight"], dYt3)
        k.linear_dx(P["fc_normals.weight"], dYtn, out=dH2t)
        dYt2 = act_bwd("fc2", "bn2", 1, dH2t)
        G["fc2.weight"] = k.linear_dw(dYt2, H1t, H1t_lo, H1t.shape[0], out=S.get("fc2.weight"))
        dH1t = k.linear_dx(P["fc2.weight"], dYt2)
        dYt1 = act_bwd("fc1", "bn1", 0, dH1t)
        G["fc1.weight"] = k.linear_dw(dYt1, Ft, Ft_lo, Kf, out=S.get("fc1.weight"))
        dFt = k.linear_dx(P["fc1.weight"], dYt1)
        fork.join(dF_mask, *[G[n] for n in G if n.startswith(("sm_", "mask_conf"))])
        dFt += dF_mask
        d_feat = k.empty(B, Kf)
        check(lib.mpb_head_to_batch_major(ptr(dFt), None, B, Bp, Kf, ptr(d_feat), stream_ptr()), "mpb_head_to_batch_major")
        # running statistics: None; parameters whose gradient went straight into the flat buffer: None as well
        grads = [None if S.get(n) is not None else G.get(n) for n in ctx.names]
        if gradsink.HOOKS["heads_done"] is not None:
            gradsink.HOOKS["heads_done"]()
        return (d_feat, None, *grads)

    @staticmethod
    def param_names():
        names = []
        for l in HeadsFunction.LINEARS:
            names += [l + ".weight", l + ".bias"]
        for b in HeadsFunction.BNS:
            names += [b + ".weight", b + ".bias", b + ".running_mean", b + ".running_var"]
        return names


def _tf32_round(x):
    """cvt.rna.tf32.f32 on a torch tensor (round to nearest, ties away): add half an ulp of the 10-bit mantissa, truncate."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


class FusedHeads:
    """Callable bound to a PointNet2Regressor_StrokeMasks: runs its head modules' parameters through HeadsFunction."""

    def __init__(self, model):
        self.model = model
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item())       # from torch's CPU generator: follows torch.manual_seed
        self.step = None

    @staticmethod
    def supported(model):
        return (model.pred_stroke_masks and model.mask_confidence_scores and not model.segment_confidence_scores and model.outdim_orient > 0
                and model.outdim == model.outdim_orient and model.outdim % 3 == 0)

    def __call__(self, feat, precision):
        m = self.model
        if self.step is None or self.step.device != feat.device:
            self.step = torch.zeros(1, dtype=torch.int64, device=feat.device)
        params = []
        for name in HeadsFunction.param_names():
            mod, attr = name.split(".")
            params.append(getattr(getattr(m, mod), attr))
        bns = [getattr(m, b) for b in HeadsFunction.BNS]
        cfg = dict(training=m.training, drop_p=float(m.dropout.p), seed=self.seed, step=self.step, weight_orient=float(m.weight_orient),
                   gemm_dtype=GEMM_DTYPE[precision], momentum=tuple(b.momentum if b.momentum is not None else 0.1 for b in bns),
                   eps=tuple(b.eps for b in bns))
        out, masks, scores = HeadsFunction.apply(feat, cfg, *params)
        if m.training:
            counters = [b.num_batches_tracked for b in bns if b.num_batches_tracked is not None]
            if counters:
                torch._foreach_add_(counters, 1)      # BatchNorm1d's own bookkeeping, one launch for the four layers
        B = feat.shape[0]
        return out.view(B, m.out_vectors, -1), masks.view(B, m.n_stroke_masks, -1), scores
