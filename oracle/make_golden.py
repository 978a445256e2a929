"""Freezes golden vectors from the REAL reference -- TEST INFRASTRUCTURE, run in the build container.

    python -B -m oracle.make_golden

Imports /root/reference (oracle/ref_loader.py), runs the reference's own functions on seeded inputs,
asserts that oracle/torch_oracle.py and oracle/c/mp_oracle.c reproduce them (bit-for-bit for the
encoder, exactly for the chamfer wrapper logic over the shared knn shim) and writes small fixtures
to tests/golden/.  The GPU box has no /root/reference: tests there read only these files.
"""
import os
import sys

import numpy as np
import torch

from maskplanner_b200 import synthetic
from oracle import c_oracle as C
from oracle import ref_loader
from oracle import torch_oracle as T

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_fps(R, xyz, npoint, seed):
    """Run the reference FPS with a chosen seed vector by staging the CPU generator (its :77 draw)."""
    B, N, _ = xyz.shape
    # find a generator state that reproduces `seed`: simplest is to monkeypatch randint for the call
    orig = torch.randint
    try:
        torch.randint = lambda *a, **k: seed.clone()
        return R.farthest_point_sample(xyz, npoint)
    finally:
        torch.randint = orig


def boundary_cloud():
    """Lattice with spacing 0.1: many points at EXACTLY distance 0.2 / 0.4 from lattice queries, plus
    duplicates and ties for FPS."""
    g = torch.arange(-5, 6, dtype=torch.float32) * 0.1
    pts = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)  # 1331 points
    dup = pts[::7]
    return torch.cat([pts, dup], 0)[None].contiguous()  # [1, 1522, 3]


def encoder_fixtures(R):
    cases = {}
    torch.manual_seed(0)
    specs = [  # name, cloud, npoint, (radius, nsample) list
        ("cube_small", synthetic.make_clouds(3, 700, seed0=11, kind="cube"), 64, [(0.2, 32), (0.4, 64), (0.3, 16)]),
        ("cuboid_small", synthetic.make_clouds(2, 1024, seed0=21, kind="cuboid"), 128, [(0.2, 32), (0.4, 64)]),
        ("lattice", boundary_cloud(), 200, [(0.2, 32), (0.4, 64), (0.1, 8), (0.3, 16)]),
        ("tiny", synthetic.make_clouds(2, 5, seed0=31, kind="cube"), 9, [(0.2, 4), (5.0, 5)]),
        ("one_point", synthetic.make_clouds(1, 1, seed0=41, kind="cube"), 3, [(0.2, 1)]),
    ]
    for name, xyz, npoint, balls in specs:
        B, N, _ = xyz.shape
        seed = torch.randint(0, N, (B,))
        ref_idx = _ref_fps(R, xyz, npoint, seed)
        assert torch.equal(ref_idx, T.farthest_point_sample(xyz, npoint, seed)), name
        assert np.array_equal(ref_idx.numpy(), C.fps(xyz, npoint, seed)), name
        new_xyz = R.index_points(xyz, ref_idx)
        cases[name + "/xyz"] = xyz.numpy()
        cases[name + "/seed"] = seed.numpy()
        cases[name + "/fps"] = ref_idx.numpy().astype(np.int32)
        for r, k in balls:
            q = R.query_ball_point(r, k, xyz, new_xyz)
            assert torch.equal(q, T.query_ball_point(r, k, xyz, new_xyz)), (name, r, k)
            assert np.array_equal(q.numpy(), C.ball_query(r, k, xyz, new_xyz)), (name, r, k)
            cases["%s/ball_r%g_k%d" % (name, r, k)] = q.numpy().astype(np.int32)
        # grouped tensor with synthetic features (xyz-first concat, centred)
        feats = torch.randn(B, N, 5, generator=torch.Generator().manual_seed(7))
        r, k = balls[0]
        orig = torch.randint
        try:
            torch.randint = lambda *a, **kw: seed.clone()
            nx, npts = R.sample_and_group(npoint, r, k, xyz, feats)
        finally:
            torch.randint = orig
        tx, tp = T.sample_and_group(npoint, r, k, xyz, feats, seed_idx=seed)
        assert torch.equal(nx, tx) and torch.equal(npts, tp), name
        cases[name + "/feats"] = feats.numpy()
        cases[name + "/grouped"] = npts.numpy()
    np.savez_compressed(os.path.join(OUT, "encoder_small.npz"), **cases)
    print("encoder_small.npz:", len(cases), "arrays")


def encoder_model_shapes(R):
    """Model / micro-benchmark shapes: inputs are regenerated from seeds, only indices are stored."""
    cases = {}
    for name, B, kind, npoint, r, k in [("sa1_cuboid", 2, "cuboid", 512, 0.2, 32), ("mu_cube", 2, "cube", 1024, 0.2, 32),
                                         ("mu_cuboid", 1, "cuboid", 1024, 0.2, 32)]:
        xyz = synthetic.make_clouds(B, 5120, seed0=1000, kind=kind)
        seed = torch.randint(0, 5120, (B,), generator=torch.Generator().manual_seed(3))
        ref_idx = _ref_fps(R, xyz, npoint, seed)
        assert np.array_equal(ref_idx.numpy(), C.fps(xyz, npoint, seed)), name
        new_xyz = R.index_points(xyz, ref_idx)
        q = R.query_ball_point(r, k, xyz, new_xyz)
        assert np.array_equal(q.numpy(), C.ball_query(r, k, xyz, new_xyz)), name
        cases[name + "/seed"] = seed.numpy()
        cases[name + "/fps"] = ref_idx.numpy().astype(np.int16)
        cases[name + "/ball"] = q.numpy().astype(np.int16)
        if name == "sa1_cuboid":  # SA2 geometry on SA1's output positions
            seed2 = torch.randint(0, 512, (B,), generator=torch.Generator().manual_seed(4))
            idx2 = _ref_fps(R, new_xyz, 128, seed2)
            assert np.array_equal(idx2.numpy(), C.fps(new_xyz, 128, seed2))
            nx2 = R.index_points(new_xyz, idx2)
            q2 = R.query_ball_point(0.4, 64, new_xyz, nx2)
            assert np.array_equal(q2.numpy(), C.ball_query(0.4, 64, new_xyz, nx2))
            cases["sa2/seed"] = seed2.numpy()
            cases["sa2/fps"] = idx2.numpy().astype(np.int16)
            cases["sa2/ball"] = q2.numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "encoder_model_shapes.npz"), **cases)
    print("encoder_model_shapes.npz:", len(cases), "arrays")


def sa_module_fixture(R):
    """A small PointNetSetAbstraction (train and eval mode) + a group-all layer, weights included."""
    torch.manual_seed(123)
    B, N = 2, 300
    xyz = synthetic.make_clouds(B, N, seed0=51, kind="cube").permute(0, 2, 1).contiguous()    # [B,3,N]
    feats = torch.randn(B, 6, N)
    cases = {"xyz": xyz.numpy(), "feats": feats.numpy()}
    sa = R.PointNetSetAbstraction(npoint=40, radius=0.45, nsample=12, in_channel=6 + 3, mlp=[16, 24, 32], group_all=False)
    sa_all = R.PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=32 + 3, mlp=[32, 48], group_all=True)
    for m in (sa, sa_all):
        for bn in m.mlp_bns:
            bn.weight.data.uniform_(0.5, 1.5)
            bn.bias.data.uniform_(-0.3, 0.3)
    for k, v in sa.state_dict().items():
        cases["sa.init/" + k] = v.numpy().copy()
    for k, v in sa_all.state_dict().items():
        cases["sa_all.init/" + k] = v.numpy().copy()
    mine = T.PointNetSetAbstraction(40, 0.45, 12, 9, [16, 24, 32], False)
    mine_all = T.PointNetSetAbstraction(None, None, None, 35, [32, 48], True)
    mine.load_state_dict(sa.state_dict())
    mine_all.load_state_dict(sa_all.state_dict())
    seed = torch.randint(0, N, (B,))
    cases["seed"] = seed.numpy()
    for mode in ("train", "eval"):
        for m in (sa, sa_all, mine, mine_all):
            m.train(mode == "train")
        x1 = xyz.clone()
        f1 = feats.clone().requires_grad_(True)
        orig = torch.randint
        try:
            torch.randint = lambda *a, **kw: seed.clone()
            nx, nf = sa(x1, f1)
        finally:
            torch.randint = orig
        gx, gf = sa_all(nx, nf)
        loss = (gf ** 2).sum() + nf.sum()
        grads = torch.autograd.grad(loss, [f1] + list(sa.parameters()) + list(sa_all.parameters()))
        f2 = feats.clone().requires_grad_(True)
        mx, mf = mine(xyz.clone(), f2, seed_idx=seed)
        mgx, mgf = mine_all(mx, mf)
        assert torch.equal(nx, mx) and torch.equal(nf, mf) and torch.equal(gf, mgf), mode
        cases[mode + "/new_xyz"] = nx.detach().numpy()
        cases[mode + "/new_points"] = nf.detach().numpy()
        cases[mode + "/global"] = gf.detach().numpy()
        cases[mode + "/grad_feats"] = grads[0].numpy()
        names = ["sa." + k for k, _ in sa.named_parameters()] + ["sa_all." + k for k, _ in sa_all.named_parameters()]
        for n, g in zip(names, grads[1:]):
            cases[mode + "/grad/" + n] = g.numpy()
        if mode == "train":
            for k, v in sa.state_dict().items():
                cases["sa.after_train/" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "sa_module_small.npz"), **cases)
    print("sa_module_small.npz:", len(cases), "arrays")


def chamfer_fixtures(RC):
    import itertools
    cases = {}
    g = torch.Generator().manual_seed(99)
    N, P1, P2 = 3, 37, 45
    for D in (3, 6, 24):
        x = torch.randn(N, P1, D, generator=g)
        y = torch.randn(N, P2, D, generator=g)
        ypad = y.clone()
        for b, l in enumerate([45, 20, 31]):
            ypad[b, l:] = -100
        cases["D%d/x" % D] = x.numpy()
        cases["D%d/y" % D] = y.numpy()
        cases["D%d/ypad" % D] = ypad.numpy()
        for padded, asym, rev, pr, br in itertools.product([False, True], [False, True], [False, True],
                                                           ["mean", "sum", None], ["mean", None]):
            if asym and rev:
                continue
            if br is not None and pr is None:
                continue
            if pr is None and not (asym or rev):
                continue  # P1 != P2: the reference raises on cham_x + cham_y
            yy = ypad if padded else y
            kw = dict(padded=padded, asymmetric=asym, reverse_asymmetric=rev, point_reduction=pr, batch_reduction=br,
                      return_matching=True)
            x1 = x.clone().requires_grad_(True)
            y1 = yy.clone().requires_grad_(True)
            d, _, xi, yi = RC.chamfer_distance(x1, y1, **kw)
            x2 = x.clone().requires_grad_(True)
            y2 = yy.clone().requires_grad_(True)
            d2, _, xi2, yi2 = T.chamfer_distance(x2, y2, **kw)
            assert torch.equal(d, d2) and torch.equal(xi, xi2) and torch.equal(yi, yi2), kw
            d.sum().backward()
            d2.sum().backward()
            assert torch.allclose(x1.grad, x2.grad, rtol=1e-6, atol=1e-7)
            key = "D%d/p%d_a%d_r%d_%s_%s" % (D, padded, asym, rev, pr, br)
            cases[key + "/dist"] = d.detach().numpy()
            cases[key + "/xi"] = xi.numpy().astype(np.int16)
            cases[key + "/yi"] = yi.numpy().astype(np.int16)
            cases[key + "/gx"] = x1.grad.numpy()
            cases[key + "/gy"] = y1.grad.numpy()
    # explicit lengths + weights + normals (D=3)
    x, y = torch.from_numpy(cases["D3/x"]), torch.from_numpy(cases["D3/y"])
    xl, yl, w = torch.tensor([37, 10, 25]), torch.tensor([45, 45, 7]), torch.tensor([1.0, 0.5, 2.0])
    xn, yn = torch.randn(N, P1, 3, generator=g), torch.randn(N, P2, 3, generator=g)
    d, dn = RC.chamfer_distance(x, y, x_lengths=xl.clone(), y_lengths=yl.clone(), weights=w, x_normals=xn, y_normals=yn)
    d2, dn2 = T.chamfer_distance(x, y, x_lengths=xl.clone(), y_lengths=yl.clone(), weights=w, x_normals=xn, y_normals=yn)
    assert torch.equal(d, d2) and torch.allclose(dn, dn2)
    cases.update({"lw/xl": xl.numpy(), "lw/yl": yl.numpy(), "lw/w": w.numpy(), "lw/xn": xn.numpy(), "lw/yn": yn.numpy(),
                  "lw/dist": d.numpy(), "lw/normals": dn.numpy()})
    np.savez_compressed(os.path.join(OUT, "chamfer_small.npz"), **cases)
    print("chamfer_small.npz:", len(cases), "arrays")


def msg_fp_fixture(R):
    """PointNetSetAbstractionMsg and PointNetFeaturePropagation (off the MaskPlanner path): frozen straight
    from the reference modules (no oracle restatement; the GPU tests compare the drop-in modules with these files)."""
    torch.manual_seed(321)
    B, N = 2, 400
    xyz = synthetic.make_clouds(B, N, seed0=61, kind="cube").permute(0, 2, 1).contiguous()
    feats = torch.randn(B, 5, N)
    cases = {"xyz": xyz.numpy(), "feats": feats.numpy()}
    msg = R.PointNetSetAbstractionMsg(48, [0.3, 0.6], [8, 16], 5, [[16, 32], [24, 40]])
    fp = R.PointNetFeaturePropagation(72 + 5, [32, 16])
    for k, v in msg.state_dict().items():
        cases["msg/" + k] = v.numpy().copy()
    for k, v in fp.state_dict().items():
        cases["fp/" + k] = v.numpy().copy()
    seed = torch.randint(0, N, (B,))
    cases["seed"] = seed.numpy()
    msg.train(), fp.train()
    f1 = feats.clone().requires_grad_(True)
    orig = torch.randint
    try:
        torch.randint = lambda *a, **kw: seed.clone()
        nx, nf = msg(xyz, f1)
    finally:
        torch.randint = orig
    up = fp(xyz, nx, f1, nf)                       # propagate the 72 multi-scale channels back to all N points
    loss = (up ** 2).sum()
    g = torch.autograd.grad(loss, [f1] + list(msg.parameters()))
    cases["msg_new_xyz"], cases["msg_new_points"], cases["fp_out"] = nx.detach().numpy(), nf.detach().numpy(), up.detach().numpy()
    cases["grad_feats"] = g[0].numpy()
    for (n, _), gr in zip(msg.named_parameters(), g[1:]):
        cases["grad/msg." + n] = gr.numpy()
    np.savez_compressed(os.path.join(OUT, "msg_fp_small.npz"), **cases)
    print("msg_fp_small.npz:", len(cases), "arrays")


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def reference_loss_config():
    """Resolved values of config=[maskplanner,<cat>,longx_v2] that loss_handler.py reads
    (asymm_chamfer_v9.yaml, default.yaml; mask weights at their post-delay targets, delayMasksLoss.yaml:5-6)."""
    return _Cfg(lambda_points=4, extra_data=["orientnorm"], per_segment_confidence=False, smooth_target_stroke_masks=False,
                weight_asymm_segment_chamfer=1.0, weight_reverse_asymm_point_chamfer=100, weight_reverse_asymm_segment_chamfer=0.01,
                explicit_weight_stroke_masks=1.0, explicit_weight_stroke_masks_confidence=100., explicit_no_stroke_weight=1.0,
                weight_asymm_v6_chamfer_with_stroke_masks=1.0, soft_attraction=False)


def step_fixture():
    """Whole model + loss + one Adam step: the REAL reference model and LossHandler (CPU, cuda calls shimmed)
    against oracle/step_oracle.py -- asserted bit-identical, then frozen as scalars / small samples."""
    from oracle import step_oracle as SO
    SSG = ref_loader.pointnet2_cls_ssg()
    LH = ref_loader.loss_handler()
    cases = {}
    B = 2
    batch = synthetic.make_batch(B, "windows_v2", seed0=0)
    torch.manual_seed(0)
    ref = SSG.PointNet2Regressor_StrokeMasks(out_vectors=449, outdim=12, outdim_orient=12, weight_orient=0.25,
                                             hidden_size=[1024, 1024], pred_stroke_masks=True, n_stroke_masks=22,
                                             mask_confidence_scores=True, segment_confidence_scores=False)
    mine = SO.Regressor(449, n_stroke_masks=22)
    mine.load_state_dict(ref.state_dict())
    lh = LH.LossHandler(["asymm_v6_chamfer_with_stroke_masks"], reference_loss_config())
    opt_r = torch.optim.Adam(ref.parameters(), lr=1e-3)
    opt_m = torch.optim.Adam(mine.parameters(), lr=1e-3)
    cloud = batch["point_cloud"].permute(0, 2, 1).float()
    # eval-mode forward at the INITIAL weights, frozen as strided samples of every output (weights after an
    # Adam step are machine dependent: its first step is sign-like and amplifies last-bit gradient noise)
    ref.eval(), mine.eval()
    torch.manual_seed(12)
    with torch.no_grad():
        a = ref(cloud)
    torch.manual_seed(12)
    e1 = torch.randint(0, 5120, (B,), dtype=torch.long)
    e2 = torch.randint(0, 512, (B,), dtype=torch.long)
    with torch.no_grad():
        b = mine(cloud, (e1, e2))
    for name, x, y in zip(("traj_pred", "masks", "scores"), a[:3], b[:3]):
        assert torch.equal(x, y), name
        cases["eval/" + name + "_sample"] = x.reshape(-1)[::97].numpy()
        cases["eval/" + name + "_sum"] = np.float64(x.double().sum())
    cases["eval/seeds1"], cases["eval/seeds2"] = e1.numpy(), e2.numpy()
    ref.train(), mine.train()
    torch.manual_seed(11)
    ref.zero_grad()
    pred, masks, scores, seg = ref(cloud)
    with ref_loader.cpu_cuda_shim():
        loss_r, _ = lh.compute(y_pred=pred, y=batch["traj"].clone(), pred_stroke_masks=masks, mask_scores=scores, seg_logits=seg,
                               stroke_ids=batch["stroke_ids"], traj_as_pc=batch["traj_as_pc"].clone())
    loss_r.backward()
    opt_r.step()
    torch.manual_seed(11)
    loss_m = SO.train_step(mine, opt_m, batch)            # draws the two FPS seeds and the dropout masks in the same order
    assert float(loss_r) == loss_m, (float(loss_r), loss_m)
    for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
        assert torch.equal(p, q), n
    cases["train/loss"] = np.float64(loss_m)
    torch.manual_seed(11)
    s1 = torch.randint(0, 5120, (B,), dtype=torch.long)
    s2 = torch.randint(0, 512, (B,), dtype=torch.long)
    cases["seeds1"], cases["seeds2"] = s1.numpy(), s2.numpy()
    # loss terms for fixed predictions (no model involved)
    g = torch.Generator().manual_seed(0)
    pred = synthetic.noisy_predictions(batch["traj"], 449, seed=1)
    masks, scores = torch.randn(B, 22, 449, generator=g), torch.randn(B, 22, generator=g)
    with ref_loader.cpu_cuda_shim():
        lr_, _ = lh.compute(y_pred=pred, y=batch["traj"].clone(), pred_stroke_masks=masks, mask_scores=scores, seg_logits=None,
                            stroke_ids=batch["stroke_ids"], traj_as_pc=batch["traj_as_pc"].clone())
    lm, terms = SO.asymm_v6_loss(pred, batch["traj"].clone(), masks, scores, batch["stroke_ids"], batch["traj_as_pc"].clone(),
                                 return_terms=True)
    assert torch.equal(lr_, lm)
    cases["loss/total"] = np.float64(lm)
    for k in ("asymm_segment", "reverse_point", "reverse_segment", "masks"):
        cases["loss/" + k] = np.float64(terms[k])
    cases["loss/match"] = terms["match"].numpy().astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "step_small.npz"), **cases)
    print("step_small.npz:", len(cases), "arrays")


def main():
    if not ref_loader.available():
        sys.exit("reference tree not found at %s" % ref_loader.REF)
    os.makedirs(OUT, exist_ok=True)
    R = ref_loader.pointnet2_utils()
    RC = ref_loader.pytorch3d_chamfer()
    encoder_fixtures(R)
    encoder_model_shapes(R)
    sa_module_fixture(R)
    chamfer_fixtures(RC)
    msg_fp_fixture(R)
    step_fixture()


if __name__ == "__main__":
    main()
