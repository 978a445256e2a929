"""Tensor-core shared MLP (tcgen05 GEMMs + bf16 BN/ReLU/max-pool kernels, maskplanner_b200.shared_mlp):
GEMM kernels against fp32 matmul, and the fused forward/backward against a torch emulation that rounds to
bf16 at exactly the same points (so what is compared is the kernels, not the conditioning of the network).
Tolerances: fp32-output GEMM rel 1e-5; bf16-output GEMM rel 1e-2 (one bf16 rounding = 2^-9); fused stack
rel-L2 1e-2 on outputs and gradients."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from maskplanner_b200 import _cabi
    return _cabi


def _check_gemm():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("check_gemm", os.path.join(os.path.dirname(__file__), "..", "tools", "check_gemm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("dt", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1000, 128, 192), (4096, 256, 128), (8192, 512, 320), (8192, 1024, 512),
                                   (300, 160, 64), (70000, 64, 64), (5, 32, 64)])
def test_gemm_tn_matches_float64_matmul(M, N, K, dt):
    """mpb_sa_gemm_tn against float64 matmul for bf16 (rel 1.5e-2: one bf16 output rounding), single-pass TF32 (2e-3) and
    3xTF32 (2e-5), each with and without the on-the-fly BatchNorm+ReLU of the A operand and with the fused forward
    (sum, sum of squares) and backward (sum dY, sum dY*z) statistics checked against the stored output."""
    cg = _check_gemm()
    for xform in (False, True):
        for epi in (0, 1, 2):
            if xform and epi == 2:
                continue
            r = cg.run_tn(dt, M, N, K, xform, epi)
            assert not r.startswith("FAIL"), (xform, epi, r)


@pytest.mark.parametrize("dt", [0, 1, 2])
@pytest.mark.parametrize("M,N,K", [(64, 128, 64), (1024, 64, 64), (5000, 128, 192), (100000, 256, 128), (8192, 1024, 512),
                                   (8192, 256, 320), (777, 64, 64)])
def test_gemm_wgrad_matches_float64_matmul_and_is_deterministic(M, N, K, dt):
    """mpb_sa_gemm_wgrad (MN-major operands straight from the row-major layout, fixed-order split-M reduction) against
    float64, with and without the operand transform; two runs must agree bit for bit."""
    cg = _check_gemm()
    for xform in (False, True):
        r = cg.run_wg(dt, M, N, K, xform)
        assert not r.startswith("FAIL"), (xform, r)
    r = cg.run_wg(dt, 4096, 128, 192, True, cout=100, cin=131, xyz_last=True)
    assert not r.startswith("FAIL"), r


@pytest.mark.parametrize("M,N,K,Kg", [(4096, 64, 128, 32), (8192, 128, 256, 64), (70016, 64, 128, 32), (1024, 320, 256, 128), (1024, 64, 128, 256), (128, 64, 64, 16)])
def test_gemm_tn_pool_rebuilds_dz_from_the_stored_preactivation(M, N, K, Kg):
    """mpb_sa_gemm_tn_pool: the dgrad GEMM of a max-pooled layer with its dZ operand rebuilt in shared memory from the stored
    pre-activation, the arg-max rows, p*dY and (-w, e) -- against float64 matmul on the explicitly formed (bf16-rounded) dZ."""
    cg = _check_gemm()
    for epi in (0, 2):
        r = cg.run_tn_pool(M, N, K, Kg, epi)
        assert not r.startswith("FAIL"), (epi, r)


@pytest.mark.parametrize("M,N,K,Kg", [(4096, 128, 64, 32), (8192, 256, 128, 64), (70016, 128, 64, 32), (1024, 64, 192, 128), (1024, 128, 64, 256), (64, 64, 64, 16)])
def test_gemm_wgrad_pool_rebuilds_dz_and_is_deterministic(M, N, K, Kg):
    cg = _check_gemm()
    for xform in (False, True):
        r = cg.run_wg_pool(M, N, K, Kg, xform)
        assert not r.startswith("FAIL"), (xform, r)


def _emulate(a0, K, convs, bns, mode="bf16"):
    """Same math as the kernels in plain torch.  mode "bf16": rounding to bf16 where the kernels store or feed bf16
    (weights, pre-activations, the normalised operand of the next GEMM); "tf32"/"fp32": float64 ground truth."""
    r = (lambda t: t.bfloat16().float()) if mode == "bf16" else (lambda t: t)
    x = a0.float() if mode == "bf16" else a0.double()
    M = x.shape[0]
    for i, (conv, bn) in enumerate(zip(convs, bns)):
        cout, cin = conv.weight.shape[:2]
        w = r(conv.weight.view(cout, cin)).to(x.dtype)
        z = r(x[:, :cin] @ w.t())
        mean, var = z.mean(0), z.var(0, unbiased=False)
        s = bn.weight.to(x.dtype) / torch.sqrt(var + bn.eps)
        x = F.relu(z * s + (bn.bias.to(x.dtype) - mean * s))
        if i < len(convs) - 1:
            x = r(x)
    return x.view(M // K, K, -1).max(1)[0]


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


@pytest.mark.parametrize("mode,tol_out,tol_grad", [("bf16", 1e-3, 1e-2), ("tf32", 5e-3, 0.15), ("fp32", 2e-5, 1e-4)])
@pytest.mark.parametrize("G,K,cin,mlp", [(64, 12, 9, [16, 24, 32]), (2, 40, 35, [32, 48]), (512, 32, 3, [64, 64, 128]),
                                          (256, 64, 131, [128, 128, 256]), (4, 128, 259, [256, 512, 1024])])
def test_fused_stack_forward_backward_matches_emulation(G, K, cin, mlp, mode, tol_out, tol_grad):
    """bf16: against the torch emulation with the kernels' rounding points (rel-L2 1e-3 outputs, 1e-2 gradients).
    tf32 (one TF32 pass) and fp32 (3xTF32): against a float64 run of the same stack -- the fp32 mode must sit at
    fp32 accuracy (2e-5 outputs, 1e-4 gradients: BatchNorm backward amplifies rounding by the conditioning of the
    batch statistics)."""
    from maskplanner_b200.shared_mlp import pad64, shared_mlp_max
    torch.manual_seed(G + K)
    convs, bns, c = nn.ModuleList(), nn.ModuleList(), cin
    for co in mlp:
        convs.append(nn.Conv2d(c, co, 1))
        bns.append(nn.BatchNorm2d(co))
        c = co
    convs.cuda(), bns.cuda()
    for bn in bns:
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.uniform_(-0.3, 0.3)
    M = G * K
    a0 = F.pad(torch.randn(M, cin, device="cuda"), (0, pad64(cin) - cin))
    if mode == "bf16":
        a0 = a0.bfloat16()
    wout = torch.randn(G, mlp[-1], device="cuda")
    params = list(convs.parameters()) + list(bns.parameters())
    rm0 = [bn.running_mean.clone() for bn in bns]
    a1 = a0.clone().requires_grad_(True)
    out1 = shared_mlp_max(a1, K, convs, bns, True, mode=mode)
    (out1 * wout).sum().backward()
    g1 = [a1.grad.float()] + [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    a2 = (a0.clone().float() if mode == "bf16" else a0.clone().double()).requires_grad_(True)
    out2 = _emulate(a2, K, convs, bns, mode)
    (out2 * wout.to(out2.dtype)).sum().backward()
    g2 = [a2.grad] + [(p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for p in params]
    assert tuple(out1.shape) == (G, mlp[-1])
    assert _rel_l2(out1, out2) < tol_out, _rel_l2(out1, out2)
    names = ["a0"] + ["conv." + n for n, _ in convs.named_parameters()] + ["bn." + n for n, _ in bns.named_parameters()]
    for n, x, y in zip(names, g1, g2):
        if n.startswith("conv.") and n.endswith("bias"):
            assert float(x.abs().max()) == 0.0, n          # exact zero: training-mode BN removes the conv bias
            continue
        assert _rel_l2(x, y) < tol_grad, (n, _rel_l2(x, y))
    # running statistics: momentum update with the unbiased variance and the conv bias folded into the mean
    for i, bn in enumerate(bns):
        assert int(bn.num_batches_tracked) == 1
        assert not torch.equal(bn.running_mean, rm0[i])


def test_running_stats_update_matches_batchnorm2d():
    from maskplanner_b200.shared_mlp import pad64, shared_mlp_max
    torch.manual_seed(1)
    conv, bn = nn.Conv2d(20, 40, 1).cuda(), nn.BatchNorm2d(40).cuda()
    ref_bn = nn.BatchNorm2d(40).cuda()
    M, K = 960, 8
    a0 = F.pad(torch.randn(M, 20, device="cuda"), (0, 44)).bfloat16()
    shared_mlp_max(a0, K, [conv], [bn], True)
    z = (a0.float()[:, :20] @ conv.weight.view(40, 20).bfloat16().float().t()).bfloat16().float() + conv.bias
    ref_bn.train()
    ref_bn(z.t().reshape(1, 40, M, 1))
    assert torch.allclose(bn.running_mean, ref_bn.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-4, atol=1e-5)


def test_short_training_run_tracks_fp32_path():
    """Functional check of the benchmarked configuration: 12 optimisation steps on one batch, bf16
    tensor-core MLP vs strict-fp32 MLP from identical initial weights (dropout off).  Whole-network
    gradients are ill-conditioned at random init (nine training-mode BatchNorms; even fp32 CPU vs fp64
    differs by 5e-3), so the criterion is the optimisation trajectory, not element-wise gradients."""
    from maskplanner_b200 import pointnet2_utils as P
    from maskplanner_b200 import synthetic
    from maskplanner_b200.train_step import Trainer
    old = P.get_mlp_precision()
    try:
        curves = {}
        batch = synthetic.make_batch(8, "windows_v2", seed0=5)
        for prec in ("fp32", "bf16"):
            P.set_mlp_precision(prec)
            tr = Trainer("windows_v2", torch.device("cuda", 0), seed=11)
            tr.model.dropout.p = 0.0
            dev_batch = tr.to_device(batch)
            gen = torch.Generator().manual_seed(3)
            losses = []
            for _ in range(12):
                seeds = (torch.randint(0, 5120, (8,), generator=gen), torch.randint(0, 512, (8,), generator=gen))
                losses.append(float(tr.step(dev_batch, seeds).item()))
            curves[prec] = losses
        f, b = curves["fp32"], curves["bf16"]
        assert np.isclose(b[0], f[0], rtol=1e-2)                 # same starting loss
        assert f[-1] < 0.5 * f[0] and b[-1] < 0.5 * b[0]          # both optimise
        assert abs(b[-1] - f[-1]) < 0.25 * f[-1], (f, b)          # and end up in the same place
    finally:
        P.set_mlp_precision(old)


@pytest.mark.parametrize("D,layout", [(0, "pm"), (6, "cm"), (128, "pm"), (128, "cm"), (10, "pm")])
def test_bf16_group_rows_features_first(D, layout):
    """mpb_group_points_bf16: rows [feats | centred xyz | 0-pad] rounded to bf16, and its scatter-add backward."""
    from maskplanner_b200 import pointnet2_utils as P
    from maskplanner_b200.shared_mlp import pad64
    from oracle import torch_oracle as T
    g = torch.Generator().manual_seed(D)
    B, N, S, K = 2, 200, 16, 8
    xyz = torch.rand(B, N, 3, generator=g)
    idx = torch.randint(0, N, (B, S, K), generator=g)
    new_xyz = T.index_points(xyz, torch.randint(0, N, (B, S), generator=g))
    feats = torch.rand(B, N, D, generator=g) if D else None
    fd = None
    if D:
        fd = feats.cuda() if layout == "pm" else feats.permute(0, 2, 1).contiguous().cuda().permute(0, 2, 1)   # channel-major storage
        fd.requires_grad_(True)
    ldo = pad64(3 + D)
    rows = P._GroupPointsBF16.apply(xyz.cuda(), fd, new_xyz.cuda(), idx.cuda(), ldo)
    assert rows.dtype == torch.bfloat16 and tuple(rows.shape) == (B * S * K, ldo)
    want = torch.zeros(B, S, K, ldo)
    if D:
        want[..., :D] = T.index_points(feats, idx)
    want[..., D:D + 3] = T.index_points(xyz, idx) - new_xyz[:, :, None]
    assert torch.equal(rows.float().cpu(), want.reshape(-1, ldo).bfloat16().float())
    if D:
        w = torch.randn(B * S * K, ldo, generator=g).bfloat16()
        (rows.float() * w.cuda().float()).sum().backward()
        fo = feats.clone().requires_grad_(True)
        (T.index_points(fo, idx).reshape(-1, D) * w[:, :D].float()).sum().backward()
        assert torch.allclose(fd.grad.cpu(), fo.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("precision,tol_out,tol_grad", [("bf16", 2e-3, 1e-2), ("fp32", 2e-5, 2e-4)])
@pytest.mark.parametrize("D,radius", [(0, 0.2), (3, 0.2), (0, 0.004)])
def test_narrow_first_layer_matches_materialised_rows(D, radius, precision, tol_out, tol_grad, monkeypatch):
    """SA1-style module (3 or 6 input channels): the on-the-fly first layer (mpb_sa_first_layer[_bwd], rows
    gathered inside the kernel, fused statistics / fused weight gradient) against the materialised path
    (mpb_group_points_bf16 -> tcgen05 GEMM -> bwd_apply -> wgrad GEMM) on the same inputs and weights.  Same bf16
    rounding points, different fp32 summation order: outputs rel 2e-3 (max-abs), every parameter gradient rel-L2 1e-2.
    radius 0.004 leaves most balls with the centre only (rows padded with the first hit)."""
    import copy
    from maskplanner_b200 import pointnet2_utils as P
    from maskplanner_b200 import synthetic
    B, N, S, K = 3, 2048, 128, 16
    cloud = synthetic.make_clouds(B, N, seed0=41 + D).cuda()                    # [B, N, 3]
    xyz = cloud.permute(0, 2, 1).contiguous()
    pts = None
    if D:
        g = torch.Generator().manual_seed(8)
        pts = F.normalize(torch.randn(B, D, N, generator=g), dim=1).cuda()
    torch.manual_seed(12)
    sa = P.PointNetSetAbstraction(S, radius, K, 3 + D, [32, 48, 64], False).cuda().train()
    sa.precision = precision
    sa_ref = copy.deepcopy(sa)
    seed = torch.tensor([7, 700, 1999])
    outs = []
    for mod, narrow in ((sa, "1"), (sa_ref, "0")):
        monkeypatch.setenv("MPB_NARROW_FIRST", narrow)
        from maskplanner_b200 import _cabi
        n0 = _cabi.KERNEL_LAUNCHES
        new_xyz, feat = mod(xyz, pts, seed_idx=seed)
        w = torch.linspace(-1.0, 1.0, feat.numel(), device="cuda").view_as(feat)
        (feat * w).sum().backward()
        outs.append((new_xyz, feat, _cabi.KERNEL_LAUNCHES - n0))
    (x1, f1, l1), (x0, f0, l0) = outs
    assert torch.equal(x1, x0)
    assert l1 < l0                                             # really took the other code path (fewer launches)
    assert float((f1 - f0).abs().max() / f0.abs().max()) < tol_out
    for (n, p1), (_, p0) in zip(sa.named_parameters(), sa_ref.named_parameters()):
        if n.endswith("mlp_convs.0.bias") or ".bias" in n and "convs" in n:
            assert float(p1.grad.abs().max()) == 0.0 and float(p0.grad.abs().max()) == 0.0   # exact zeros under BN
            continue
        assert _rel_l2(p1.grad, p0.grad) < tol_grad, (n, _rel_l2(p1.grad, p0.grad))
    for b1, b0 in zip(sa.mlp_bns, sa_ref.mlp_bns):
        assert torch.allclose(b1.running_mean, b0.running_mean, rtol=1e-3, atol=1e-5)
        assert torch.allclose(b1.running_var, b0.running_var, rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("cout,cin,xyz_last", [(64, 3, True), (64, 6, True), (128, 131, True), (256, 259, False), (24, 35, False)])
def test_pack_weight_matches_torch_ops(cout, cin, xyz_last):
    """mpb_pack_weight_bf16: zero-padded bf16 operand, its transpose, and the xyz-last column permutation."""
    c = _lib()
    lib = c.load()
    g = torch.Generator(device="cuda").manual_seed(cout + cin)
    W = torch.randn(cout, cin, device="cuda", generator=g)
    cout_p, cin_p = (cout + 63) // 64 * 64, (8 if cin <= 8 else (cin + 63) // 64 * 64)
    wp = torch.full((cout_p, cin_p), 7.0, dtype=torch.bfloat16, device="cuda")
    wt = torch.full((cin_p, cout_p), 7.0, dtype=torch.bfloat16, device="cuda")
    c.check(lib.mpb_pack_weight_bf16(c.ptr(W), cout, cin, cout_p, cin_p, int(xyz_last), c.ptr(wp), c.ptr(wt), c.stream_ptr()), "pack")
    want = torch.zeros(cout_p, cin_p, device="cuda")
    src = torch.cat([W[:, 3:], W[:, :3]], dim=1) if (xyz_last and cin > 3) else W
    want[:cout, :cin] = src
    assert torch.equal(wp, want.bfloat16())
    assert torch.equal(wt, want.bfloat16().t().contiguous())


@pytest.mark.parametrize("dt", [0, 1])
@pytest.mark.parametrize("G,K,C", [(4, 128, 1024), (64, 128, 1024), (3, 70, 512), (5, 33, 64), (1000, 16, 64)])
def test_bn_relu_max_small_and_large_group_counts(G, K, C, dt):
    """mpb_bn_relu_max picks the 1024-thread K-split kernel for few groups and the one-thread-per-(group, 8
    channels) kernel otherwise: values, FIRST arg-max (torch.max semantics) and the saved pre-activation must agree
    with torch on both, including ties (bf16 inputs collide often) and all-negative columns (relu -> 0, arg 0)."""
    c = _lib()
    lib = c.load()
    g = torch.Generator(device="cuda").manual_seed(G * K + C)
    Z = (torch.randn(G * K, C, device="cuda", generator=g) * 2).bfloat16()      # bf16-representable values: ties in both dtypes
    if dt == 1:
        Z = Z.float()
    Z[:, 0] = -Z[:, 0].abs() - 1.0                                  # a column that relu kills everywhere
    scale = torch.rand(C, device="cuda", generator=g) + 0.5
    scale[1::3] *= -1.0                                             # decreasing channels: the minimum of z is pooled
    shift = torch.randn(C, device="cuda", generator=g) * 0.1
    shift[0] = 0.0
    out = torch.empty(G, C, device="cuda")
    arg = torch.empty(G, C, dtype=torch.int32, device="cuda")
    zmax = torch.empty(G, C, device="cuda")
    c.check(lib.mpb_bn_relu_max(dt, c.ptr(Z), c.ptr(scale), c.ptr(shift), G, K, C, c.ptr(out), c.ptr(arg), c.ptr(zmax), c.stream_ptr()), "relu_max")
    act = torch.relu((Z.double() * scale.double() + shift.double()).float()).view(G, K, C)   # == fmaf(z, scale, shift)
    want = act.max(dim=1).values
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-7)
    first = (act == want.unsqueeze(1)).int().argmax(dim=1).int()     # first k attaining the maximum
    live = want > 0     # groups the ReLU kills entirely carry no gradient: the many-group kernel reports the arg-max of s*z there
    assert torch.equal(arg[live], first[live])
    assert torch.equal(zmax[live], Z.float().view(G, K, C).gather(1, first.long().unsqueeze(1)).squeeze(1)[live])
    zs = Z.float().view(G, K, C) * torch.where(scale < 0, -1.0, 1.0)
    assert bool(((arg == first) | (arg == zs.argmax(dim=1).int()) | (zs.gather(1, arg.long().unsqueeze(1)).squeeze(1) == zs.max(dim=1).values)).all())


def _float64_stack(a, K, convs, bns, argmax=None, masks=None):
    """The reference's op sequence (pointnet2_utils.py:208-214) in float64, no rounding model: 1x1 conv -> training-mode
    BatchNorm -> ReLU per layer, then the max over the K rows of each group.  The two DISCRETE decisions of the stack can be
    imposed from outside, everything else (products, statistics) staying free: `masks[l]` [M, C_l] replaces the ReLU's own
    sign test, `argmax` [G, C] the pooling selection.  Returns (pooled, list of the float64 run's own ReLU masks)."""
    x, own = a, []
    for l, (conv, bn) in enumerate(zip(convs, bns)):
        cout, cin = conv.weight.shape[:2]
        z = x[:, :cin] @ conv.weight.view(cout, cin).double().t()
        mean, var = z.mean(0), z.var(0, unbiased=False)
        sc = bn.weight.double() / torch.sqrt(var + bn.eps)
        y = z * sc + (bn.bias.double() - mean * sc)
        own.append(y > 0)
        x = torch.where(masks[l] if masks is not None else own[-1], y, torch.zeros_like(y))
    x = x.view(x.shape[0] // K, K, -1)
    if argmax is None:
        return x.max(1)[0], own
    return x.gather(1, argmax.long()[:, None, :]).squeeze(1), own


@pytest.mark.parametrize("G,K,cin,mlp", [(512, 32, 3, [64, 64, 128]), (256, 64, 131, [128, 128, 256]), (8, 128, 259, [256, 512, 1024])])
def test_bf16_stack_against_float64_ground_truth(G, K, cin, mlp):
    """north_star: outputs AND gradients of the bf16 tensor-core path within 1e-2 of the reference arithmetic.  The oracle is the
    reference's own op sequence in FLOAT64 on the same rows and weights (no rounding model of the kernels), at the three model
    shapes (sa1, sa2, sa3).  Measured on B200 (printed by the test):
      * outputs: 3-5e-3 rel-L2 (bound 1e-2);
      * gradients, everything free on both sides: 0.12-0.19.  That is not rounding noise but DISCRETE flips: with 8 mantissa
        bits a few of the 32-128 candidates of a (group, channel) tie at the top, so bf16 and float64 route the channel's
        gradient to different rows (imposing the kernels' arg-max rows alone brings the last layer to 6-8e-3 and the others
        to 5-9e-2), and a pre-activation within 2^-9 of the ReLU kink lands on the other side of it: 0.03-0.2 % of the
        decisions per layer, and a fraction f of flipped mask entries alone is a sqrt(f) relative error (0.2 % -> 4.5e-2).
        No implementation that stores bf16 activations avoids it; it is recorded, and bounded through the flip FRACTION
        (< 0.5 % of the ReLU decisions, asserted), not through the gradient norm;
      * gradients with the bf16 path's discrete decisions (arg-max rows, ReLU masks) imposed on the float64 run, everything
        else -- products, batch statistics -- free: the arithmetic error proper: 4-8e-3 for every parameter and the input,
        bound 1e-2 (asserted)."""
    from maskplanner_b200.shared_mlp import pad64, shared_mlp_max
    torch.manual_seed(1234 + G)
    convs, bns, c = nn.ModuleList(), nn.ModuleList(), cin
    for co in mlp:
        convs.append(nn.Conv2d(c, co, 1))
        bns.append(nn.BatchNorm2d(co))
        c = co
    convs.cuda(), bns.cuda()
    for bn in bns:
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.uniform_(-0.3, 0.3)
    M, L = G * K, len(mlp)
    a0 = F.pad(torch.randn(M, cin, device="cuda"), (0, pad64(cin) - cin)).bfloat16()     # the rows both sides see (bf16-exact)
    wout = torch.randn(G, mlp[-1], device="cuda")
    params = list(convs.parameters()) + list(bns.parameters())
    names = ["a0"] + ["conv." + n for n, _ in convs.named_parameters()] + ["bn." + n for n, _ in bns.named_parameters()]
    a1 = a0.clone().requires_grad_(True)
    out1 = shared_mlp_max(a1, K, convs, bns, True, mode="bf16")
    saved = out1.grad_fn.saved_tensors                                                    # SharedMLPMax: arg-max, Z_l, (scale, shift, ..)_l
    argmax = saved[0][:, :mlp[-1]]
    k_masks = [(saved[1 + l].float() * saved[1 + L + l][0] + saved[1 + L + l][1] > 0)[:, :mlp[l]] for l in range(L)]
    (out1 * wout).sum().backward()
    g1 = [a1.grad.float()] + [p.grad.clone() for p in params]
    report, own_masks = {}, None
    for tag, am, mk in (("free", None, None), ("pool imposed", argmax, None), ("pool+relu imposed", argmax, k_masks)):
        for p in params:
            p.grad = None
        a2 = a0.double().requires_grad_(True)
        out2, own = _float64_stack(a2, K, convs, bns, am, mk)
        own_masks = own_masks or own
        (out2 * wout.double()).sum().backward()
        g2 = [a2.grad] + [(p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for p in params]
        report[tag] = (_rel_l2(out1, out2), {n: _rel_l2(x, y) for n, x, y in zip(names, g1, g2)
                                             if not (n.startswith("conv.") and n.endswith("bias"))})
    flips = [float((a != b).float().mean()) for a, b in zip(k_masks, own_masks)]
    for tag, (e_out, errs) in report.items():
        print("bf16 vs float64 (%s): out %.2e, grads %s" % (tag, e_out, {n: "%.1e" % e for n, e in errs.items()}))
    print("ReLU decisions that differ from the float64 run, per layer:", ["%.2e" % f for f in flips])
    assert all(r[0] < 1e-2 for r in report.values()), report
    assert max(flips) < 5e-3, flips
    for n, e in report["pool+relu imposed"][1].items():
        assert e < 1e-2, (n, e)
