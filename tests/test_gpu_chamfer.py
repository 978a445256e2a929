"""GPU parity for the chamfer drop-in (through the C ABI) against golden vectors and the oracle.
Tolerance (BASELINE.json north_star): values rel 1e-4 in fp32; indices equal except at distance ties."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as C
from oracle import torch_oracle as T

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-6


@pytest.fixture(scope="module")
def CH():
    from maskplanner_b200 import pytorch3d_chamfer
    return pytorch3d_chamfer


def _keys(g, D):
    return sorted({k.rsplit("/", 1)[0] for k in g.files if k.startswith("D%d/p" % D)})


@pytest.mark.parametrize("D", [3, 6, 24])
def test_chamfer_option_surface_golden(CH, golden, D):
    g = golden("chamfer_small.npz")
    for key in _keys(g, D):
        p, a, r, pr, br = key.split("/")[1].split("_")
        padded, asym, rev = p == "p1", a == "a1", r == "r1"
        pr = None if pr == "None" else pr
        br = None if br == "None" else br
        x = torch.from_numpy(g["D%d/x" % D]).cuda().requires_grad_(True)
        y = torch.from_numpy(g["D%d/ypad" % D] if padded else g["D%d/y" % D]).cuda().requires_grad_(True)
        d, nrm, xi, yi = CH.chamfer_distance(x, y, padded=padded, asymmetric=asym, reverse_asymmetric=rev, point_reduction=pr,
                                             batch_reduction=br, return_matching=True)
        assert nrm is None and xi.dtype == torch.int64
        assert np.allclose(d.detach().cpu().numpy(), g[key + "/dist"], rtol=RTOL, atol=ATOL), key
        assert np.array_equal(xi.cpu().numpy(), g[key + "/xi"]) and np.array_equal(yi.cpu().numpy(), g[key + "/yi"]), key
        d.sum().backward()
        assert np.allclose(x.grad.cpu().numpy(), g[key + "/gx"], rtol=RTOL, atol=1e-5), key
        assert np.allclose(y.grad.cpu().numpy(), g[key + "/gy"], rtol=RTOL, atol=1e-5), key


def test_lengths_weights_normals_golden(CH, golden):
    g = golden("chamfer_small.npz")
    c = lambda k: torch.from_numpy(g[k]).cuda()
    d, dn = CH.chamfer_distance(c("D3/x"), c("D3/y"), x_lengths=c("lw/xl"), y_lengths=c("lw/yl"), weights=c("lw/w"),
                                x_normals=c("lw/xn"), y_normals=c("lw/yn"))
    assert np.allclose(d.cpu().numpy(), g["lw/dist"], rtol=RTOL) and np.allclose(dn.cpu().numpy(), g["lw/normals"], rtol=RTOL)


def _maskplanner_like(B, P1, G, D, seed):
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(B, G, D, generator=g) * 0.3
    lens = torch.randint(G // 2, G + 1, (B,), generator=g)
    lens[0] = G
    for b in range(B):
        y[b, lens[b]:] = -100
    pick = torch.randint(0, G // 2, (B, P1), generator=g)
    x = torch.gather(y, 1, pick[:, :, None].expand(B, P1, D)) + 0.05 * torch.randn(B, P1, D, generator=g)
    return x, y, lens


@pytest.mark.parametrize("P1,G,D", [(999, 986, 24), (3996, 2959, 6), (449, 449, 24), (1796, 1350, 6), (500, 700, 3)])
def test_loss_handler_call_patterns_vs_oracle(CH, P1, G, D):
    """The three calls of loss_handler.py:604-645 on MaskPlanner shapes, values + matching + gradients."""
    B = 4
    x, y, lens = _maskplanner_like(B, P1, G, D, seed=P1)
    for kw in (dict(padded=True, asymmetric=True, return_matching=True, point_reduction=None, batch_reduction=None),
               dict(padded=True, reverse_asymmetric=True), dict(padded=True)):
        xo, yo = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        want = T.chamfer_distance(xo, yo, **kw)
        xg, yg = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
        got = CH.chamfer_distance(xg, yg, **kw)
        assert np.allclose(got[0].detach().cpu().numpy(), want[0].detach().numpy(), rtol=RTOL, atol=ATOL)
        if kw.get("return_matching"):
            # indices may differ only where two candidates tie within fp32 rounding: compare the distances they select
            for gi, wi, q, t in ((got[2], want[2], x, y), (got[3], want[3], y, x)):
                gi, wi = gi.cpu(), wi
                diff = gi != wi
                if diff.any():
                    dg = ((q - torch.gather(t, 1, gi[:, :, None].expand(-1, -1, D))) ** 2).sum(-1)
                    dw = ((q - torch.gather(t, 1, wi[:, :, None].expand(-1, -1, D))) ** 2).sum(-1)
                    assert torch.allclose(dg[diff], dw[diff], rtol=1e-5)
                assert diff.float().mean() < 1e-3
        want[0].sum().backward()
        got[0].sum().backward()
        assert np.allclose(xg.grad.cpu().numpy(), xo.grad.numpy(), rtol=1e-3, atol=1e-5)
        assert np.allclose(yg.grad.cpu().numpy(), yo.grad.numpy(), rtol=1e-3, atol=1e-5)


def test_padded_lengths_on_device(CH):
    y = torch.randn(5, 40, 6)
    y[1, 10:] = -100
    y[3, 0:] = -100
    y[4, 39:] = -100
    got = CH.padded_lengths(y.cuda(), None, False).cpu()
    assert got.tolist() == [40, 10, 40, 0, 39]
    # caller-supplied lengths are only overwritten when some sample is padded (reference :140)
    user = torch.tensor([7, 7, 7, 7, 7]).cuda()
    assert CH.padded_lengths(torch.randn(5, 40, 6).cuda(), user.clone(), True).cpu().tolist() == [7] * 5
    assert CH.padded_lengths(y.cuda(), user.clone(), True).cpu().tolist() == [40, 10, 40, 0, 39]


@pytest.mark.parametrize("D", [2, 5, 16, 33])
def test_generic_dimension_path(CH, D):
    g = torch.Generator().manual_seed(D)
    x, y = torch.randn(2, 130, D, generator=g), torch.randn(2, 257, D, generator=g)
    d, _, xi, yi = CH.chamfer_distance(x.cuda(), y.cuda(), return_matching=True, point_reduction=None, asymmetric=True,
                                       batch_reduction=None)
    wd, wi = C.knn(x, y, K=1)
    assert np.allclose(d.cpu().numpy(), wd[..., 0], rtol=RTOL, atol=ATOL) and np.array_equal(xi.cpu().numpy(), wi[..., 0])
    wd2, wi2 = C.knn(y, x, K=1)
    assert np.array_equal(yi.cpu().numpy(), wi2[..., 0])


def test_k2_attraction_branch_vs_oracle(CH):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 64, 6, generator=g)
    y = x + 0.01 * torch.randn(3, 64, 6, generator=g)
    want = T.chamfer_distance(x, y, avoid_in_sequence_collapsing=True, point_reduction=None, batch_reduction=None)[0]
    got = CH.chamfer_distance(x.cuda(), y.cuda(), avoid_in_sequence_collapsing=True, point_reduction=None, batch_reduction=None)[0]
    assert np.allclose(got.cpu().numpy(), want.numpy(), rtol=RTOL, atol=ATOL)
    d, i = CH.knn_points(x.cuda(), y.cuda(), K=2)[:2]
    wd, wi = C.knn(x, y, K=2)
    assert np.array_equal(i.cpu().numpy(), wi) and np.allclose(d.cpu().numpy(), wd, rtol=RTOL, atol=ATOL)


def test_min_centroids_and_velocities_branches(CH):
    g = torch.Generator().manual_seed(6)
    x, y = torch.randn(2, 50, 6, generator=g), torch.randn(2, 50, 6, generator=g)
    for kw in (dict(min_centroids=True), dict(velocities=True)):
        want = T.chamfer_distance(x, y, **kw)[0]
        got = CH.chamfer_distance(x.cuda(), y.cuda(), **kw)[0]
        assert np.allclose(got.cpu().numpy(), want.numpy(), rtol=RTOL), kw


def test_error_behaviour_matches_reference(CH):
    x, y = torch.randn(2, 5, 3).cuda(), torch.randn(2, 6, 3).cuda()
    with pytest.raises(ValueError, match="batch_reduction"):
        CH.chamfer_distance(x, y, batch_reduction="max")
    with pytest.raises(ValueError, match="point_reduction"):
        CH.chamfer_distance(x, y, point_reduction=None)
    with pytest.raises(ValueError, match="correct shape"):
        CH.chamfer_distance(x, torch.randn(2, 6, 4).cuda())
    with pytest.raises(ValueError, match="shape \\(N, P, D\\)"):
        CH.chamfer_distance(x[0], y)
    with pytest.raises(ValueError, match="negative"):
        CH.chamfer_distance(x, y, weights=torch.tensor([1.0, -1.0]).cuda())
    z = CH.chamfer_distance(x, y, weights=torch.zeros(2).cuda())
    assert float(z[0]) == 0.0


@pytest.mark.parametrize("P,D", [(8192, 3), (4096, 6), (2048, 24)])
def test_large_sweep_properties(CH, P, D):
    """Size-independent properties at sweep sizes: the returned distance is the distance AT the returned
    index, no target beats it (checked on a random subset), symmetric = asymmetric + reverse."""
    B = 8
    g = torch.Generator().manual_seed(P)
    x, y = torch.randn(B, P, D, generator=g).cuda(), torch.randn(B, P, D, generator=g).cuda()
    dx, _, xi, yi = CH.chamfer_distance(x, y, asymmetric=True, return_matching=True, point_reduction=None, batch_reduction=None)
    at = ((x - torch.gather(y, 1, xi[:, :, None].expand(-1, -1, D))) ** 2).sum(-1)
    assert torch.allclose(dx, at, rtol=1e-5, atol=1e-6)
    sub = torch.randint(0, P, (256,), generator=g).cuda()
    cand = ((x[:, :, None, :] - y[:, None, sub, :]) ** 2).sum(-1).min(-1)[0]
    assert (dx <= cand * (1 + 1e-5) + 1e-6).all()
    a = CH.chamfer_distance(x, y, asymmetric=True)[0]
    r = CH.chamfer_distance(x, y, reverse_asymmetric=True)[0]
    s = CH.chamfer_distance(x, y)[0]
    assert torch.allclose(a + r, s, rtol=1e-5)
    # full check against the oracle for one batch element
    wd, wi = C.knn(x[:1].cpu(), y[:1].cpu(), K=1)
    assert np.allclose(dx[:1].cpu().numpy(), wd[..., 0], rtol=RTOL, atol=ATOL)
    assert (xi[:1].cpu().numpy() != wi[..., 0]).mean() < 1e-3


def test_empty_and_degenerate_inputs(CH):
    x = torch.randn(2, 4, 3).cuda()
    y = torch.full((2, 3, 3), -100.0).cuda()          # every GT row is padding -> y_lengths = 0
    d, _, xi, yi = CH.chamfer_distance(x, y, padded=True, asymmetric=True, return_matching=True, point_reduction=None, batch_reduction=None)
    assert (d == 0).all() and (xi == 0).all() and (yi == 0).all()
    d0 = CH.chamfer_distance(torch.randn(0, 4, 3).cuda(), torch.randn(0, 3, 3).cuda(), batch_reduction=None)[0]
    assert d0.numel() == 0
