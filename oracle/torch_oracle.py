"""CPU restatement of the hot path in plain torch ops -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Encoder functions follow the reference's op order so that, on CPU, they are bit-identical to it
(oracle/make_golden.py asserts this against the imported reference).  All file:line citations are
relative to /root/reference.
"""
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import c_oracle

# --------------------------------------------------------------------------------------------
# models/pointnet2_utils.py
# --------------------------------------------------------------------------------------------


def square_distance(src, dst):
    """models/pointnet2_utils.py:21-42: -2*src@dst^T, then += |src|^2, then += |dst|^2 (that order)."""
    B, N, _ = src.shape
    M = dst.shape[1]
    d = torch.matmul(src, dst.transpose(1, 2)) * -2
    d += (src ** 2).sum(-1).view(B, N, 1)
    d += (dst ** 2).sum(-1).view(B, 1, M)
    return d


def index_points(points, idx):
    """models/pointnet2_utils.py:45-62: points[b, idx[b, ...], :]."""
    B = points.shape[0]
    shape = [B] + [1] * (idx.dim() - 1)
    bidx = torch.arange(B, dtype=torch.long, device=points.device).view(shape).expand_as(idx)
    return points[bidx, idx, :]


def draw_fps_seed(B, N, device=None):
    """models/pointnet2_utils.py:77: one CPU-generator randint(0, N, (B,)) per FPS call (then `.to(device)`)."""
    return torch.randint(0, N, (B,), dtype=torch.long).to(device or "cpu")


def farthest_point_sample(xyz, npoint, seed_idx=None):
    """models/pointnet2_utils.py:65-86.  `seed_idx` replaces the :77 draw when given."""
    B, N, _ = xyz.shape
    dev = xyz.device                                                   # :74 (every allocation follows the input's device)
    picked = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    nearest = torch.ones(B, N, device=dev) * 1e10                      # :76
    cur = draw_fps_seed(B, N, dev) if seed_idx is None else seed_idx.clone().long().to(dev)
    rows = torch.arange(B, dtype=torch.long, device=dev)
    for i in range(npoint):                                            # :79
        picked[:, i] = cur                                             # :80
        c = xyz[rows, cur, :].view(B, 1, 3)                            # :81
        d = torch.sum((xyz - c) ** 2, -1)                              # :82
        nearest = torch.where(d < nearest, d, nearest)                 # :83-84 (strict <)
        cur = torch.max(nearest, -1)[1]                                # :85 (first max index)
    return picked


def query_ball_point(radius, nsample, xyz, new_xyz):
    """models/pointnet2_utils.py:89-109, same sort-based construction as the reference."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    gi = torch.arange(N, dtype=torch.long, device=xyz.device).view(1, 1, N).repeat(B, S, 1)    # :102
    d = square_distance(new_xyz, xyz)                                        # :103
    gi[d > radius ** 2] = N                                                  # :104
    gi = gi.sort(dim=-1)[0][:, :, :nsample]                                  # :105
    first = gi[:, :, :1].expand(B, S, nsample)                               # :106
    return torch.where(gi == N, first, gi)                                   # :107-108


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, full_points=None, seed_idx=None):
    """models/pointnet2_utils.py:112-148."""
    B, N, C = xyz.shape
    fps_idx = farthest_point_sample(xyz, npoint, seed_idx)                   # :130
    new_xyz = index_points(xyz, fps_idx)                                     # :131
    idx = query_ball_point(radius, nsample, xyz, new_xyz)                    # :132
    grouped_xyz = index_points(xyz, idx)                                     # :133
    centred = grouped_xyz - new_xyz.view(B, npoint, 1, C)                    # :134
    if points is not None:
        new_points = torch.cat([centred, index_points(points, idx)], dim=-1)  # :136-138 (xyz first)
    elif full_points is not None:
        new_points = index_points(full_points, idx)                          # :139-141
    else:
        new_points = centred
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """models/pointnet2_utils.py:151-168: one group holding every point, not centred."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device)
    g = xyz.view(B, 1, N, C)
    if points is not None:
        g = torch.cat([g, points.view(B, 1, N, -1)], dim=-1)
    return new_xyz, g


class PointNetSetAbstraction(nn.Module):
    """models/pointnet2_utils.py:171-216.  Same attribute / parameter names (state_dict compatible)."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        c = in_channel
        for co in mlp:
            self.mlp_convs.append(nn.Conv2d(c, co, 1))
            self.mlp_bns.append(nn.BatchNorm2d(co))
            c = co

    def forward(self, xyz, points, full_points=None, seed_idx=None):
        xyz = xyz.permute(0, 2, 1)                                           # :196
        points = points.permute(0, 2, 1) if points is not None else None
        full_points = full_points.permute(0, 2, 1) if full_points is not None else None
        if self.group_all:
            new_xyz, g = sample_and_group_all(xyz, points)                   # :203
        else:
            new_xyz, g = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points,
                                          full_points=full_points, seed_idx=seed_idx)   # :205
        g = g.permute(0, 3, 2, 1)                                            # :208 -> [B,C,K,S]
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            g = F.relu(bn(conv(g)))                                          # :210-212
        return new_xyz.permute(0, 2, 1), torch.max(g, 2)[0]                  # :214-215


# --------------------------------------------------------------------------------------------
# pytorch3d.ops.knn (third-party, not vendored): restated semantics, see oracle/__init__.py
# --------------------------------------------------------------------------------------------
KNN = namedtuple("KNN", "dists idx knn")


class _KnnFn(torch.autograd.Function):
    """knn_points forward (no grad tracking, C loop) + analytic backward, as pytorch3d does:
    grad_p1[n,i] += 2*g[n,i,k]*(p1[n,i]-p2[n,idx]);  grad_p2[n,idx] -= the same."""

    @staticmethod
    def forward(ctx, p1, p2, len1, len2, K):
        dev = p1.device
        if p1.dtype == torch.float64 or p1.is_cuda:
            # float64 "ground truth" variant (tests use it to separate conditioning from kernel error) and the
            # CUDA stand-in for pytorch3d's knn kernel (bench.py's same-box reference arm): brute force in torch,
            # direct-form distances, same candidate/length/tie rules (stable sort / first minimum => lowest index)
            d, i = [], []
            step = max(1, (64 << 20) // max(1, p1.shape[1] * p2.shape[1] * p1.shape[2]))    # <= 256 MB of differences
            for b0 in range(0, p1.shape[0], step):
                q, t, l1, l2 = p1[b0:b0 + step], p2[b0:b0 + step], len1[b0:b0 + step], len2[b0:b0 + step]
                full = ((q[:, :, None, :] - t[:, None, :, :]) ** 2).sum(-1)
                full = full.masked_fill(torch.arange(t.shape[1], device=dev)[None, None, :] >= l2[:, None, None], float("inf"))
                if K == 1:
                    dd, ii = full.min(dim=-1, keepdim=True)
                else:
                    dd, ii = torch.sort(full, dim=-1, stable=True)
                    dd, ii = dd[:, :, :K].clone(), ii[:, :, :K].clone()
                dead = (torch.arange(q.shape[1], device=dev)[None, :, None] >= l1[:, None, None]) | torch.isinf(dd)
                d.append(dd.masked_fill(dead, 0.0))
                i.append(ii.masked_fill(dead, 0))
            d, i = torch.cat(d), torch.cat(i)
        else:
            d, i = c_oracle.knn(p1, p2, len1, len2, K=K, use_fma=True)
            d, i = torch.from_numpy(d), torch.from_numpy(i)
        ctx.save_for_backward(p1, p2, len1, len2, i)
        ctx.mark_non_differentiable(i)
        return d, i

    @staticmethod
    def backward(ctx, gd, _gi):
        p1, p2, len1, len2, idx = ctx.saved_tensors
        N, P1, D = p1.shape
        K = idx.shape[2]
        # pytorch3d's backward kernel: only (p1_idx < lengths1[n] and k < lengths2[n]) contribute
        dev = p1.device
        valid = ((torch.arange(P1, device=dev)[None, :, None] < len1[:, None, None])
                 & (torch.arange(K, device=dev)[None, None, :] < len2[:, None, None])).to(p1.dtype)[..., None]
        nb = p2[torch.arange(N, device=dev)[:, None, None], idx]                              # [N,P1,K,D]
        diff = (p1[:, :, None, :] - nb) * (2.0 * gd[..., None]) * valid
        g1 = diff.sum(2)
        g2 = torch.zeros_like(p2)
        g2.scatter_add_(1, idx.reshape(N, P1 * K, 1).expand(N, P1 * K, D), -diff.reshape(N, P1 * K, D))
        return g1, g2, None, None, None


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
    """pytorch3d.ops.knn.knn_points (call sites pytorch3d_chamfer.py:182-183, 205-206, 257-258)."""
    assert norm == 2
    N, P1, _ = p1.shape
    P2 = p2.shape[1]
    if lengths1 is None:
        lengths1 = torch.full((N,), P1, dtype=torch.int64, device=p1.device)
    if lengths2 is None:
        lengths2 = torch.full((N,), P2, dtype=torch.int64, device=p1.device)
    keep = torch.float64 if p1.dtype == torch.float64 else torch.float32
    d, i = _KnnFn.apply(p1.contiguous().to(keep), p2.contiguous().to(keep), lengths1.long(), lengths2.long(), K)
    nn_pts = knn_gather(p2, i, lengths2) if return_nn else None
    return KNN(d, i, nn_pts)


def knn_gather(x, idx, lengths=None):
    """pytorch3d.ops.knn.knn_gather (call site pytorch3d_chamfer.py:274-275): x[n, idx[n,l,k], :],
    zero-filled where k >= lengths[n]."""
    N, M, U = x.shape
    _, L, K = idx.shape
    out = x[torch.arange(N, device=x.device)[:, None, None], idx]
    if lengths is not None:
        short = lengths[:, None] <= torch.arange(K, device=x.device)[None]
        if short.any():
            out = out.masked_fill(short[:, None, :, None].expand(N, L, K, U), 0.0)
    return out


# --------------------------------------------------------------------------------------------
# pytorch3d_chamfer.py
# --------------------------------------------------------------------------------------------


def padded_lengths(y, y_lengths, sentinel=-100):
    """pytorch3d_chamfer.py:138-149.  If any y[b, :, 0] equals the sentinel, EVERY sample's length is
    rewritten: first sentinel column for padded samples, P2 for the others; otherwise untouched."""
    hit = y[:, :, 0] == sentinel
    if not hit.any():
        return y_lengths
    P2 = y.shape[1]
    first = torch.where(hit.any(1), hit.float().argmax(1), torch.full((y.shape[0],), P2, dtype=torch.long, device=y.device))
    y_lengths[:] = first.to(y_lengths.dtype)
    return y_lengths


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction="mean", point_reduction="mean", velocities=False, min_centroids=False,
                     padded=False, avoid_in_sequence_collapsing=False, soft_attraction=False,
                     asymmetric=False, reverse_asymmetric=False, return_matching=False):
    """pytorch3d_chamfer.py:76-344, restated (tensor inputs only; Pointclouds objects are a
    pytorch3d type the reference never passes)."""
    if not soft_attraction:                                                       # :124-125, :16-30
        if batch_reduction is not None and batch_reduction not in ("mean", "sum"):
            raise ValueError('batch_reduction must be one of ["mean", "sum"] or None')
        if batch_reduction is not None and point_reduction not in ("mean", "sum"):
            raise ValueError('point_reduction must be one of ["mean", "sum"] if batch_reduction is not None')

    def _prep(pts, lengths, normals):                                             # :38-73
        if not torch.is_tensor(pts):
            raise ValueError("The input pointclouds should be either Pointclouds objects or torch.Tensor "
                             "of shape (minibatch, num_points, 3).")
        if pts.ndim != 3:
            raise ValueError("Expected points to be of shape (N, P, D)")
        if lengths is not None and (lengths.ndim != 1 or lengths.shape[0] != pts.shape[0]):
            raise ValueError("Expected lengths to be of shape (N,)")
        if lengths is None:
            lengths = torch.full((pts.shape[0],), pts.shape[1], dtype=torch.int64, device=pts.device)
        if normals is not None and normals.ndim != 3:
            raise ValueError("Expected normals to be of shape (N, P, 3")
        return pts, lengths, normals

    x, x_lengths, x_normals = _prep(x, x_lengths, x_normals)
    y, y_lengths, y_normals = _prep(y, y_lengths, y_normals)
    with_normals = x_normals is not None and y_normals is not None
    N, P1, D = x.shape
    P2 = y.shape[1]
    if padded:
        y_lengths = padded_lengths(y, y_lengths)                                  # :138-149
    x_ragged = bool((x_lengths != P1).any())                                      # :152-153
    y_ragged = bool((y_lengths != P2).any())
    x_mask = torch.arange(P1, device=x.device)[None] >= x_lengths[:, None]        # :154-159
    y_mask = torch.arange(P2, device=x.device)[None] >= y_lengths[:, None]
    if y.shape[0] != N or y.shape[2] != D:
        raise ValueError("y does not have the correct shape.")                    # :161-162
    if weights is not None:                                                       # :163-176
        if weights.size(0) != N:
            raise ValueError("weights must be of shape (N,).")
        if not (weights >= 0).all():
            raise ValueError("weights cannot be negative.")
        if weights.sum() == 0.0:
            w = weights.view(N, 1)
            z = (x.sum((1, 2)) * w)
            if batch_reduction in ("mean", "sum"):
                return z.sum() * 0.0, z.sum() * 0.0
            return z * 0.0, z * 0.0

    cn_x = x.new_zeros(())
    cn_y = x.new_zeros(())
    x_nn = y_nn = None
    if velocities:                                                                # :180-198
        assert D == 6, 'Velocities is True but traj does not contain velocities'
        xi = knn_points(x[:, :, :3], y[:, :, :3], lengths1=x_lengths, lengths2=y_lengths, K=1).idx
        yi = knn_points(y[:, :, :3], x[:, :, :3], lengths1=y_lengths, lengths2=x_lengths, K=1).idx
        cham_x = torch.linalg.norm(x - y[torch.arange(N, device=x.device)[:, None], xi[..., 0]], dim=-1).square()
        cham_y = torch.linalg.norm(y - x[torch.arange(N, device=x.device)[:, None], yi[..., 0]], dim=-1).square()
    elif avoid_in_sequence_collapsing:                                            # :200-239
        assert P1 == P2
        seq = torch.arange(P1, device=x.device)
        x_nn = knn_points(x, y, lengths1=x_lengths, lengths2=y_lengths, K=2)
        y_nn = knn_points(y, x, lengths1=y_lengths, lengths2=x_lengths, K=2)
        x_self = x_nn.idx[:, :, 0] == seq[None]
        y_self = y_nn.idx[:, :, 0] == seq[None]
        if not soft_attraction:
            cham_x = torch.where(x_self, x_nn.dists[:, :, 1], x_nn.dists[:, :, 0]).sum(1)
            cham_y = torch.where(y_self, y_nn.dists[:, :, 1], y_nn.dists[:, :, 0]).sum(1)
        else:
            assert point_reduction is None and batch_reduction is None
            cham_x = torch.stack([x_nn.dists[b, ~x_self[b], 0].mean() for b in range(N)]).mean()
            cham_y = torch.stack([y_nn.dists[b, ~y_self[b], 0].mean() for b in range(N)]).mean()
    else:                                                                         # :241-261
        if min_centroids:
            assert P1 == P2
            assert D % 3 == 0
            lam = D // 3
            y = y.view(N, P1, lam, 3).mean(dim=-2)
            x = x.view(N, P1, lam, 3).mean(dim=-2)
        x_nn = knn_points(x, y, lengths1=x_lengths, lengths2=y_lengths, K=1)
        y_nn = knn_points(y, x, lengths1=y_lengths, lengths2=x_lengths, K=1)
        cham_x = x_nn.dists[..., 0]
        cham_y = y_nn.dists[..., 0]

    if x_ragged:                                                                  # :263-266
        cham_x = cham_x.masked_fill(x_mask, 0.0)
    if y_ragged:
        cham_y = cham_y.masked_fill(y_mask, 0.0)
    if weights is not None:                                                       # :268-270
        cham_x = cham_x * weights.view(N, 1)
        cham_y = cham_y * weights.view(N, 1)

    if with_normals:                                                              # :272-291
        xn_near = knn_gather(y_normals, x_nn.idx, y_lengths)[..., 0, :]
        yn_near = knn_gather(x_normals, y_nn.idx, x_lengths)[..., 0, :]
        cn_x = 1 - torch.abs(F.cosine_similarity(x_normals, xn_near, dim=2, eps=1e-6))
        cn_y = 1 - torch.abs(F.cosine_similarity(y_normals, yn_near, dim=2, eps=1e-6))
        if x_ragged:
            cn_x = cn_x.masked_fill(x_mask, 0.0)
        if y_ragged:
            cn_y = cn_y.masked_fill(y_mask, 0.0)
        if weights is not None:
            cn_x = cn_x * weights.view(N, 1)
            cn_y = cn_y * weights.view(N, 1)

    if point_reduction is not None and not avoid_in_sequence_collapsing:          # :295-308
        cham_x, cham_y = cham_x.sum(1), cham_y.sum(1)
        if with_normals:
            cn_x, cn_y = cn_x.sum(1), cn_y.sum(1)
        if point_reduction == "mean":
            cham_x, cham_y = cham_x / x_lengths, cham_y / y_lengths
            if with_normals:
                cn_x, cn_y = cn_x / x_lengths, cn_y / y_lengths

    if batch_reduction is not None:                                               # :312-326
        cham_x, cham_y = cham_x.sum(), cham_y.sum()
        if with_normals:
            cn_x, cn_y = cn_x.sum(), cn_y.sum()
        if batch_reduction == "mean":
            div = weights.sum() if weights is not None else N
            cham_x, cham_y = cham_x / div, cham_y / div
            if with_normals:
                cn_x, cn_y = cn_x / div, cn_y / div

    if asymmetric:                                                                # :329-334
        dist = cham_x
    elif reverse_asymmetric:
        dist = cham_y
    else:
        dist = cham_x + cham_y
    normals_out = cn_x + cn_y if with_normals else None
    if return_matching:                                                           # :338-342
        return dist, normals_out, x_nn.idx.flatten(1, 2), y_nn.idx.flatten(1, 2)
    return dist, normals_out
