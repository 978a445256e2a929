"""GPU parity (run on the B200 box with -m gpu): the drop-in encoder ops, called through the C ABI,
against the committed golden vectors and the CPU oracle.  Indices and grouped tensors: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as C
from oracle import torch_oracle as T

pytestmark = pytest.mark.gpu

ENC_CASES = ["cube_small", "cuboid_small", "lattice", "tiny", "one_point"]
GROUP_BALL = {"cube_small": (0.2, 32), "cuboid_small": (0.2, 32), "lattice": (0.2, 32), "tiny": (0.2, 4), "one_point": (0.2, 1)}


@pytest.fixture(scope="module")
def P():
    from maskplanner_b200 import pointnet2_utils
    return pointnet2_utils


def _balls(g, name):
    out = []
    for k in g.files:
        if k.startswith(name + "/ball_r"):
            r, kk = k.split("/ball_r")[1].split("_k")
            out.append((float(r), int(kk), g[k]))
    return out


@pytest.mark.parametrize("name", ENC_CASES)
def test_fps_golden(P, golden, name):
    g = golden("encoder_small.npz")
    xyz = torch.from_numpy(g[name + "/xyz"]).cuda()
    want = g[name + "/fps"]
    got = P.farthest_point_sample(xyz, want.shape[1], seed_idx=torch.from_numpy(g[name + "/seed"]))
    assert got.dtype == torch.int64 and tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("name", ENC_CASES)
def test_ball_query_golden(P, golden, name):
    g = golden("encoder_small.npz")
    xyz = torch.from_numpy(g[name + "/xyz"]).cuda()
    new_xyz = P.index_points(xyz, torch.from_numpy(g[name + "/fps"]).long().cuda())
    for r, k, want in _balls(g, name):
        got = P.query_ball_point(r, k, xyz, new_xyz)
        assert got.dtype == torch.int64
        assert np.array_equal(got.cpu().numpy(), want), (r, k)


@pytest.mark.parametrize("name", ENC_CASES)
def test_sample_and_group_golden(P, golden, name):
    g = golden("encoder_small.npz")
    xyz = torch.from_numpy(g[name + "/xyz"]).cuda()
    feats = torch.from_numpy(g[name + "/feats"]).cuda()
    r, k = GROUP_BALL[name]
    new_xyz, grouped, gxyz, fps_idx = P.sample_and_group(g[name + "/fps"].shape[1], r, k, xyz, feats, returnfps=True,
                                                        seed_idx=torch.from_numpy(g[name + "/seed"]))
    assert np.array_equal(fps_idx.cpu().numpy(), g[name + "/fps"])
    assert np.array_equal(grouped.cpu().numpy(), g[name + "/grouped"])       # gather + one fp32 subtract: exact
    assert np.array_equal(new_xyz.cpu().numpy(), g[name + "/xyz"][np.arange(xyz.shape[0])[:, None], g[name + "/fps"]])


def test_model_and_microbench_shapes_golden(P, golden):
    from maskplanner_b200 import synthetic
    g = golden("encoder_model_shapes.npz")
    for name, B, kind, npoint in [("sa1_cuboid", 2, "cuboid", 512), ("mu_cube", 2, "cube", 1024), ("mu_cuboid", 1, "cuboid", 1024)]:
        xyz = synthetic.make_clouds(B, 5120, seed0=1000, kind=kind).cuda()
        idx = P.farthest_point_sample(xyz, npoint, seed_idx=torch.from_numpy(g[name + "/seed"]))
        assert np.array_equal(idx.cpu().numpy(), g[name + "/fps"].astype(np.int64)), name
        new_xyz = P.index_points(xyz, idx)
        ball = P.query_ball_point(0.2, 32, xyz, new_xyz)
        assert np.array_equal(ball.cpu().numpy(), g[name + "/ball"].astype(np.int64)), name
        if name == "sa1_cuboid":
            idx2 = P.farthest_point_sample(new_xyz, 128, seed_idx=torch.from_numpy(g["sa2/seed"]))
            assert np.array_equal(idx2.cpu().numpy(), g["sa2/fps"].astype(np.int64))
            ball2 = P.query_ball_point(0.4, 64, new_xyz, P.index_points(new_xyz, idx2))
            assert np.array_equal(ball2.cpu().numpy(), g["sa2/ball"].astype(np.int64))


def test_fps_draws_its_seed_like_the_reference(P):
    """models/pointnet2_utils.py:77: one CPU-generator randint per call -> same stream consumption."""
    xyz = torch.rand(4, 600, 3).cuda()
    torch.manual_seed(77)
    got = P.farthest_point_sample(xyz, 50)
    after = torch.rand(1)
    torch.manual_seed(77)
    seed = torch.randint(0, 600, (4,), dtype=torch.long)
    assert torch.equal(after, torch.rand(1))
    assert np.array_equal(got.cpu().numpy(), C.fps(xyz.cpu(), 50, seed))


def test_strided_views_need_no_copy(P):
    """The model feeds permuted views (reference :196); kernels take element strides."""
    base = (torch.rand(3, 3, 1000, generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()   # [B,3,N] physical
    xyz = base.permute(0, 2, 1)                                                                   # [B,N,3] view
    assert not xyz.is_contiguous()
    seed = torch.tensor([5, 999, 0])
    got = P.farthest_point_sample(xyz, 100, seed_idx=seed)
    want = C.fps(xyz.cpu().contiguous(), 100, seed)
    assert np.array_equal(got.cpu().numpy(), want)
    new_xyz = P.index_points(xyz, got)
    ball = P.query_ball_point(0.3, 16, xyz, new_xyz)
    assert np.array_equal(ball.cpu().numpy(), C.ball_query(0.3, 16, xyz.cpu().contiguous(), new_xyz.cpu()))
    feats = torch.rand(3, 7, 1000).cuda().permute(0, 2, 1)                                        # channel-major features
    grouped = P.group_points(xyz, feats, new_xyz, ball)
    want_g = torch.cat([T.index_points(xyz.cpu(), ball.cpu()) - new_xyz.cpu()[:, :, None], T.index_points(feats.cpu(), ball.cpu())], -1)
    assert torch.equal(grouped.cpu(), want_g)


@pytest.mark.parametrize("N,npoint", [(8192, 64), (8193, 48), (20000, 64), (100000, 96), (131072, 24), (140000, 12)])
def test_fps_large_clouds_cluster_and_streaming_paths(P, N, npoint):
    """N > 8192 runs one thread-block cluster per cloud (DSMEM arg-max); N > 131072 the streaming fallback."""
    xyz = (torch.rand(2, N, 3, generator=torch.Generator().manual_seed(N)) * 2 - 1)
    seed = torch.tensor([N - 1, N // 3])
    got = P.farthest_point_sample(xyz.cuda(), npoint, seed_idx=seed)
    assert np.array_equal(got.cpu().numpy(), C.fps(xyz, npoint, seed))


def test_fps_ties_and_exhausted_cloud(P):
    """Duplicates: once every distinct point is taken all distances are 0 and torch.max returns index 0."""
    pts = torch.tensor([[[0., 0, 0], [1, 0, 0], [1, 0, 0], [0, 0, 0], [0, 1, 0]]])
    seed = torch.tensor([3])
    got = P.farthest_point_sample(pts.cuda(), 8, seed_idx=seed)
    assert np.array_equal(got.cpu().numpy(), C.fps(pts, 8, seed))
    assert got[0, -1].item() == 0


def test_ball_query_padding_and_empty_ball(P):
    xyz = torch.tensor([[[0., 0, 0], [0.1, 0, 0], [5, 5, 5], [0.05, 0, 0]]]).cuda()
    q = torch.tensor([[[0., 0, 0], [9, 9, 9]]]).cuda()
    got = P.query_ball_point(0.2, 4, xyz, q).cpu().numpy()
    assert got[0, 0].tolist() == [0, 1, 3, 0]          # ascending hits, padded with the first
    assert got[0, 1].tolist() == [4, 4, 4, 4]          # empty ball -> N, exactly like the reference
    assert np.array_equal(got, C.ball_query(0.2, 4, xyz.cpu(), q.cpu()))


def test_ball_query_microbench_shape_vs_oracle(P):
    from maskplanner_b200 import synthetic
    xyz = synthetic.make_clouds(4, 5120, seed0=2000, kind="cube")
    seed = torch.arange(4) * 100
    idx = C.fps(xyz, 1024, seed)
    new_xyz = T.index_points(xyz, torch.from_numpy(idx))
    got = P.query_ball_point(0.2, 32, xyz.cuda(), new_xyz.cuda())
    assert np.array_equal(got.cpu().numpy(), C.ball_query(0.2, 32, xyz, new_xyz))


def test_square_distance_bit_exact(P):
    a = (torch.rand(2, 65, 3) * 2 - 1)
    b = (torch.rand(2, 300, 3) * 2 - 1)
    got = P.square_distance(a.cuda(), b.cuda())
    assert np.array_equal(got.cpu().numpy(), C.square_distance(a, b))


@pytest.mark.parametrize("k", [1, 8, 16, 32, 50])
def test_knn_group_matches_oracle(P, k):
    xyz = torch.rand(2, 3000, 3, generator=torch.Generator().manual_seed(k)) * 2 - 1
    q = xyz[:, ::30].contiguous()
    idx, d = P.knn_group(k, xyz.cuda(), q.cuda(), return_dist=True)
    widx, wd = C.knn_group(xyz, q, k)
    assert np.array_equal(d.cpu().numpy(), wd)                     # distance multiset (sorted) is exact
    assert np.array_equal(idx.cpu().numpy(), widx)                 # and the lowest-index tie-break too


def test_index_points_and_group_backward_match_autograd(P):
    g = torch.Generator().manual_seed(0)
    B, N, S, K, D = 2, 50, 7, 5, 6
    xyz = torch.rand(B, N, 3, generator=g)
    feats = torch.rand(B, N, D, generator=g)
    idx = torch.randint(0, N, (B, S, K), generator=g)
    fidx = torch.randint(0, N, (B, S), generator=g)
    w = torch.rand(B, S, K, 3 + D, generator=g)

    def run(ix_fn, grp_fn, dev):
        x = xyz.detach().clone().to(dev).requires_grad_(True)
        f = feats.detach().clone().to(dev).requires_grad_(True)
        new_xyz = ix_fn(x, fidx.to(dev))
        out = grp_fn(x, f, new_xyz, idx.to(dev))
        (out * w.to(dev)).sum().backward()
        return out.detach().cpu(), x.grad.cpu(), f.grad.cpu()

    def ref_group(x, f, new_xyz, ix):
        return torch.cat([T.index_points(x, ix) - new_xyz[:, :, None], T.index_points(f, ix)], -1)

    o1, gx1, gf1 = run(T.index_points, ref_group, "cpu")
    o2, gx2, gf2 = run(P.index_points, P.group_points, "cuda")
    assert torch.equal(o1, o2)
    assert torch.allclose(gx1, gx2, rtol=1e-5, atol=1e-6) and torch.allclose(gf1, gf2, rtol=1e-5, atol=1e-6)


def _load_sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}


def _rel(a, b):
    """max |a-b| relative to the largest magnitude of the reference tensor."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_set_abstraction_module_tensor_core_path_golden(P, golden, mode):
    """bf16 tensor-core MLP (tcgen05 GEMMs, bf16 activations, fp32 statistics) against the same fixture.
    Tolerance from BASELINE.json north_star for bf16: rel 1e-2 (forward), gradients 3e-2 of their scale."""
    g = golden("sa_module_small.npz")
    sa = P.PointNetSetAbstraction(40, 0.45, 12, 9, [16, 24, 32], False)
    sa_all = P.PointNetSetAbstraction(None, None, None, 35, [32, 48], True)
    sa.precision = sa_all.precision = "bf16"
    sa.load_state_dict(_load_sd(g, "sa.init/"))
    sa_all.load_state_dict(_load_sd(g, "sa_all.init/"))
    sa.cuda(), sa_all.cuda()
    xyz = torch.from_numpy(g["xyz"]).cuda()
    if mode == "eval":
        sa.load_state_dict(_load_sd(g, "sa.after_train/"))
        sa.eval()
        with torch.no_grad():
            nx, nf = sa(xyz, torch.from_numpy(g["feats"]).cuda(), seed_idx=torch.from_numpy(g["seed"]))
        assert np.array_equal(nx.cpu().numpy(), g["eval/new_xyz"])
        assert _rel(nf.cpu().numpy(), g["eval/new_points"]) < 1e-2
        return
    feats = torch.from_numpy(g["feats"]).cuda().requires_grad_(True)
    nx, nf = sa(xyz, feats, seed_idx=torch.from_numpy(g["seed"]))
    assert np.array_equal(nx.detach().cpu().numpy(), g["train/new_xyz"])          # geometry stays bit-exact
    assert tuple(nf.shape) == g["train/new_points"].shape
    assert _rel(nf.detach().cpu().numpy(), g["train/new_points"]) < 1e-2
    gx, gf = sa_all(nx, nf)
    assert _rel(gf.detach().cpu().numpy(), g["train/global"]) < 2e-2
    loss = (gf ** 2).sum() + nf.sum()
    grads = torch.autograd.grad(loss, [feats] + list(sa.parameters()))
    # gradients: a bf16-sized perturbation flips individual ReLU / arg-max decisions, so compare
    # directions (element-wise kernel parity is pinned in tests/test_gpu_mlp.py against a same-rounding emulation)
    def cos(a, b):
        a, b = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
        return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
    assert cos(grads[0].cpu().numpy(), g["train/grad_feats"]) > 0.9
    for (n, _), gr in zip(sa.named_parameters(), grads[1:]):
        want = g["train/grad/sa." + n]
        if "convs" in n and n.endswith("bias"):
            assert float(gr.abs().max()) == 0.0        # exact: training-mode BN removes the conv bias
            continue
        assert cos(gr.cpu().numpy(), want) > 0.9, n
    for k, v in sa.state_dict().items():              # running statistics / num_batches_tracked
        assert np.allclose(v.cpu().numpy(), g["sa.after_train/" + k], rtol=1e-2, atol=1e-3), k


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_set_abstraction_module_golden(P, golden, mode):
    """fp32 tolerance from BASELINE.json north_star: rel 1e-4 (atol scaled to the output range)."""
    g = golden("sa_module_small.npz")
    sa = P.PointNetSetAbstraction(40, 0.45, 12, 9, [16, 24, 32], False)
    sa_all = P.PointNetSetAbstraction(None, None, None, 35, [32, 48], True)
    sa.precision = sa_all.precision = "fp32"
    sa.load_state_dict(_load_sd(g, "sa.init/"))
    sa_all.load_state_dict(_load_sd(g, "sa_all.init/"))
    if mode == "eval":
        sa.load_state_dict(_load_sd(g, "sa.after_train/"))
        sa.eval()
    sa.cuda(), sa_all.cuda()
    xyz = torch.from_numpy(g["xyz"]).cuda()
    feats = torch.from_numpy(g["feats"]).cuda().requires_grad_(True)
    nx, nf = sa(xyz, feats, seed_idx=torch.from_numpy(g["seed"]))
    assert tuple(nx.shape) == g[mode + "/new_xyz"].shape and tuple(nf.shape) == g[mode + "/new_points"].shape
    assert np.array_equal(nx.detach().cpu().numpy(), g[mode + "/new_xyz"])
    assert np.allclose(nf.detach().cpu().numpy(), g[mode + "/new_points"], rtol=1e-4, atol=1e-5)
    if mode == "train":
        gx, gf = sa_all(nx, nf)
        assert np.allclose(gf.detach().cpu().numpy(), g["train/global"], rtol=1e-4, atol=1e-5)
        loss = (gf ** 2).sum() + nf.sum()
        grads = torch.autograd.grad(loss, [feats] + list(sa.parameters()))
        scale = np.abs(g["train/grad_feats"]).max()
        assert np.allclose(grads[0].cpu().numpy(), g["train/grad_feats"], rtol=1e-3, atol=1e-4 * scale)
        gmax = max(np.abs(g["train/grad/sa." + n]).max() for n, _ in sa.named_parameters())
        for (n, _), gr in zip(sa.named_parameters(), grads[1:]):
            want = g["train/grad/sa." + n]
            # conv biases feed a training-mode BatchNorm: their true gradient is 0 and both sides hold
            # rounding noise, so the absolute floor is tied to the layer-wide gradient scale
            assert np.allclose(gr.cpu().numpy(), want, rtol=1e-3, atol=1e-5 * gmax), n
        for k, v in sa.state_dict().items():
            assert np.allclose(v.cpu().numpy(), g["sa.after_train/" + k], rtol=1e-4, atol=1e-6), k


def test_state_dict_keys_match_reference_layout(P):
    sa = P.PointNetSetAbstraction(512, 0.2, 32, 3, [64, 64, 128], False)
    keys = set(sa.state_dict())
    for i, (ci, co) in enumerate([(3, 64), (64, 64), (64, 128)]):
        assert tuple(sa.state_dict()["mlp_convs.%d.weight" % i].shape) == (co, ci, 1, 1)
        for s in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            assert "mlp_bns.%d.%s" % (i, s) in keys


def test_cpu_tensors_are_refused_loudly(P):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.farthest_point_sample(torch.rand(1, 10, 3), 2)
