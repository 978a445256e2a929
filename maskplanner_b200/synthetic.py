"""Synthetic PaintNet-shaped inputs -- the PaintNet dataset is not shipped.

Shapes and padding conventions follow the reference's loader/collate
(utils/dataset/paintnet_ODv1.py:214-218, 289-294, 735-747; utils/pointcloud.py:343-354):
clouds [B,5120,3] fp32 roughly inside the unit ball; GT poses [B,T,6] = position + 0.25 * unit normal;
GT segments = sliding windows of lambda=4 poses with stride 3 inside each stroke -> [B,G,24];
batches padded to the longest sample with -100 rows and stroke id -1 (float dtype, as the loader does).
Everything is generated on the CPU from per-sample seeded generators, so every rank / test / bench
arm sees bit-identical data.
"""
import torch

LAMBDA_POINTS = 4          # configs/maskplanner/asymm_chamfer_v9.yaml:7
OVERLAPPING = 1            # :8
WEIGHT_ORIENT = 0.25       # :6
PAD = -100.0               # paintnet_ODv1.py:740-744
PC_POINTS = 5120           # configs/maskplanner/default.yaml:25

CATEGORIES = {
    # n_pred_traj_points, max_n_strokes (configs/maskplanner/{cuboids,windows}_v2.yaml)
    "cuboids_v2": dict(n_pred_traj_points=3000, max_n_strokes=6),
    "windows_v2": dict(n_pred_traj_points=1350, max_n_strokes=22),
}


def out_vectors(n_pred_traj_points, lam=LAMBDA_POINTS, overlap=OVERLAPPING):
    """models/__init__.py:309: number of predicted segments (999 cuboids, 449 windows)."""
    return (n_pred_traj_points - lam) // (lam - overlap) + 1


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def cuboid_surface_cloud(n_points, g):
    """Points uniform on the surface of an axis-aligned cuboid, half-extents ~ U(0.2, 0.7) per axis."""
    h = 0.2 + 0.5 * torch.rand(3, generator=g)
    areas = torch.stack([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
    face_axis = torch.multinomial(areas / areas.sum(), n_points, replacement=True, generator=g)
    p = (torch.rand(n_points, 3, generator=g) * 2 - 1) * h
    sign = torch.randint(0, 2, (n_points,), generator=g).float() * 2 - 1
    p[torch.arange(n_points), face_axis] = sign * h[face_axis]
    return p.float(), h


def make_clouds(B, n_points=PC_POINTS, seed0=1000, kind="cuboid"):
    """[B, n_points, 3] fp32.  kind='cuboid' (surface-like density) or 'cube' (uniform in [-1,1]^3, worst case)."""
    out = []
    for i in range(B):
        g = _gen(seed0 + i)
        if kind == "cuboid":
            out.append(cuboid_surface_cloud(n_points, g)[0])
        else:
            out.append((torch.rand(n_points, 3, generator=g) * 2 - 1).float())
    return torch.stack(out)


def make_trajectories(B, category="windows_v2", seed0=5000):
    """GT for one batch: dict(traj [B,G,24], traj_as_pc [B,T,6], stroke_ids [B,G] float, n_strokes list)."""
    cfg = CATEGORIES[category]
    lam, stride = LAMBDA_POINTS, LAMBDA_POINTS - OVERLAPPING
    segs, poses, sids, nstrokes = [], [], [], []
    for i in range(B):
        g = _gen(seed0 + i)
        ns = int(torch.randint(1, cfg["max_n_strokes"] + 1, (1,), generator=g))
        n_pts = int((0.5 + 0.5 * torch.rand(1, generator=g)) * cfg["n_pred_traj_points"])
        # split n_pts over the strokes, every stroke long enough for at least one segment
        base = max(lam, n_pts // ns)
        lens = [base] * ns
        h = 0.2 + 0.5 * torch.rand(3, generator=g)
        s_list, p_list, id_list = [], [], []
        for s, m in enumerate(lens):
            start = (torch.rand(3, generator=g) * 2 - 1) * h
            step = torch.randn(3, generator=g)
            step = 0.05 * step / step.norm()                      # equal_spaced_points_distance = 0.05
            pos = start[None] + torch.arange(m)[:, None] * step[None] + 0.002 * torch.randn(m, 3, generator=g)
            nrm = torch.randn(3, generator=g)
            nrm = (nrm / nrm.norm())[None].expand(m, 3) * WEIGHT_ORIENT
            pose = torch.cat([pos, nrm], dim=1).float()           # [m, 6]
            p_list.append(pose)
            n_seg = (m - lam) // stride + 1
            win = torch.stack([pose[k * stride:k * stride + lam].reshape(-1) for k in range(n_seg)])
            s_list.append(win)
            id_list.append(torch.full((n_seg,), float(s)))
        segs.append(torch.cat(s_list))
        poses.append(torch.cat(p_list))
        sids.append(torch.cat(id_list))
        nstrokes.append(ns)
    G = max(s.shape[0] for s in segs)
    T = max(p.shape[0] for p in poses)
    traj = torch.full((B, G, 6 * lam), PAD)
    traj_as_pc = torch.full((B, T, 6), PAD)
    stroke_ids = torch.full((B, G), -1.0)
    for i in range(B):
        traj[i, :segs[i].shape[0]] = segs[i]
        traj_as_pc[i, :poses[i].shape[0]] = poses[i]
        stroke_ids[i, :sids[i].shape[0]] = sids[i]
    return dict(traj=traj, traj_as_pc=traj_as_pc, stroke_ids=stroke_ids, n_strokes=nstrokes)


def make_batch(B, category="windows_v2", seed0=0, cloud_kind="cuboid"):
    """One training batch as the reference's collate would deliver it (train_maskplanner.py:186-193)."""
    d = make_trajectories(B, category, seed0=5000 + seed0)
    d["point_cloud"] = make_clouds(B, PC_POINTS, seed0=1000 + seed0, kind=cloud_kind)
    return d


def noisy_predictions(traj, n_out, seed=0, sigma=0.05):
    """Stand-alone chamfer benches: predictions = GT rows (resampled to n_out) + N(0, sigma^2) noise."""
    g = _gen(seed)
    B, G, D = traj.shape
    out = torch.empty(B, n_out, D)
    for b in range(B):
        valid = int((traj[b, :, 0] != PAD).sum())
        pick = torch.randint(0, max(valid, 1), (n_out,), generator=g)
        out[b] = traj[b, pick] + sigma * torch.randn(n_out, D, generator=g)
    return out
