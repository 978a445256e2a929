"""Drop-in for the reference's ``models/pointnet2_utils.py`` on B200 (sm_100a).

Same names, argument order, shapes, dtypes and error behaviour as the reference module
(/root/reference/models/pointnet2_utils.py); every function enqueues hand-written CUDA kernels from
libmaskplanner_b200.so on the current torch stream through the C ABI (include/maskplanner_b200.h).
CUDA tensors only -- there is no CPU or pure-torch fallback for the hot ops.

    square_distance          :21-42     index_points           :45-62
    farthest_point_sample    :65-86     query_ball_point       :89-109
    sample_and_group         :112-148   sample_and_group_all   :151-168
    PointNetSetAbstraction   :171-216
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from ._cabi import check, ptr, require_cuda, stream_ptr


def _f32(t, name):
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32 (the reference path is fp32); got %s" % (name, t.dtype))
    return t


def _strides3(t):
    return t.stride(0), t.stride(1), t.stride(2)


def _check_out(out, shape, device):
    if tuple(out.shape) != tuple(shape) or out.dtype != torch.int64 or out.device != device or not out.is_contiguous():
        raise ValueError("out must be a contiguous int64 tensor of shape %s on %s" % (tuple(shape), device))


# ---------------------------------------------------------------------------------------------
# a2  square_distance
# ---------------------------------------------------------------------------------------------
def square_distance(src, dst):
    """Pairwise squared distance, expanded form (reference :21-42).  src [B,N,C], dst [B,M,C] -> [B,N,M].

    C == 3 (the only case on the hot path) runs the bit-exact kernel; other C (feature-space
    distances in the off-path feature-propagation module) use the same three torch ops as the
    reference."""
    B, N, C = src.shape
    M = dst.shape[1]
    if C != 3 or src.requires_grad or dst.requires_grad:
        d = -2 * torch.matmul(src, dst.permute(0, 2, 1))
        d += torch.sum(src ** 2, -1).view(B, N, 1)
        d += torch.sum(dst ** 2, -1).view(B, 1, M)
        return d
    require_cuda(src, dst)
    s, d = _f32(src, "src").contiguous(), _f32(dst, "dst").contiguous()
    out = torch.empty(B, N, M, dtype=torch.float32, device=src.device)
    check(_cabi.load().mpb_square_distance_f32(ptr(s), ptr(d), B, N, M, ptr(out), stream_ptr()), "mpb_square_distance_f32")
    return out


# ---------------------------------------------------------------------------------------------
# a4  index_points
# ---------------------------------------------------------------------------------------------
class _IndexPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        B, N, C = points.shape
        idx_c = idx.contiguous()
        M = idx_c.numel() // max(B, 1)
        out = torch.empty(tuple(idx.shape) + (C,), dtype=points.dtype, device=points.device)
        sb, sn, sc = _strides3(points)
        check(_cabi.load().mpb_index_points_f32(ptr(points), sb, sn, sc, B, N, C, ptr(idx_c), M, ptr(out), stream_ptr()),
              "mpb_index_points_f32")
        ctx.save_for_backward(idx_c)
        ctx.shape = (B, N, C)
        return out

    @staticmethod
    def backward(ctx, go):
        (idx_c,) = ctx.saved_tensors
        B, N, C = ctx.shape
        go = go.contiguous()
        gp = torch.zeros(B, N, C, dtype=go.dtype, device=go.device)
        check(_cabi.load().mpb_index_points_bwd_f32(ptr(go), ptr(idx_c), B, N, C, idx_c.numel() // max(B, 1), ptr(gp),
                                                    stream_ptr()), "mpb_index_points_bwd_f32")
        return gp, None


def index_points(points, idx):
    """points [B,N,C] (any strides), idx [B,S] or [B,S,K] int64 -> [B,S,C] / [B,S,K,C] (reference :45-62)."""
    require_cuda(points, idx)
    _f32(points, "points")
    if idx.dtype != torch.int64:
        idx = idx.long()
    return _IndexPoints.apply(points, idx)


# ---------------------------------------------------------------------------------------------
# a1  farthest_point_sample
# ---------------------------------------------------------------------------------------------
def draw_fps_seed(B, N, device):
    """The reference's seed draw (:77): ONE torch.randint(0, N, (B,)) from the CPU generator, then
    moved to the device -- identical RNG-stream consumption, so seeded runs pick the same seeds."""
    return torch.randint(0, N, (B,), dtype=torch.long).to(device)


def farthest_point_sample(xyz, npoint, seed_idx=None, out=None):
    """xyz [B,N,3] f32 -> centroids [B,npoint] int64, bit-exact vs the reference (:65-86).

    `seed_idx` (extension, default None = reference behaviour) supplies the first index per cloud
    instead of drawing it; `out` (extension): a preallocated contiguous [B,npoint] int64 result buffer."""
    require_cuda(xyz)
    _f32(xyz, "xyz")
    B, N, C = xyz.shape
    assert C == 3, "farthest_point_sample expects [B, N, 3]"
    seed = draw_fps_seed(B, N, xyz.device) if seed_idx is None else seed_idx.to(device=xyz.device, dtype=torch.long).contiguous()
    if out is None:
        out = torch.empty(B, npoint, dtype=torch.long, device=xyz.device)
    else:
        _check_out(out, (B, npoint), xyz.device)
    lib = _cabi.load()
    ws_bytes = lib.mpb_fps_workspace_bytes(B, N)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=xyz.device) if ws_bytes else None
    sb, sn, sc = _strides3(xyz)
    check(lib.mpb_fps_f32(ptr(xyz), sb, sn, sc, B, N, ptr(seed), npoint, ptr(out), ptr(ws), stream_ptr()), "mpb_fps_f32")
    return out


# ---------------------------------------------------------------------------------------------
# a3  query_ball_point (+ kNN grouping for the stress configuration)
# ---------------------------------------------------------------------------------------------
def query_ball_point(radius, nsample, xyz, new_xyz, out=None):
    """xyz [B,N,3], new_xyz [B,S,3] -> group_idx [B,S,nsample] int64, bit-exact vs the reference (:89-109).
    `out` (extension): a preallocated contiguous [B,S,nsample] int64 result buffer."""
    require_cuda(xyz, new_xyz)
    _f32(xyz, "xyz"), _f32(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    if out is None:
        out = torch.empty(B, S, nsample, dtype=torch.long, device=xyz.device)
    else:
        _check_out(out, (B, S, nsample), xyz.device)
    r2 = float(np.float32(radius ** 2))  # torch compares fp32 tensors against fp32(radius ** 2) (:104)
    check(_cabi.load().mpb_ball_query_f32(ptr(xyz), *_strides3(xyz), ptr(new_xyz), *_strides3(new_xyz), B, N, S, r2,
                                          nsample, ptr(out), stream_ptr()), "mpb_ball_query_f32")
    return out


def knn_group(k, xyz, new_xyz, return_dist=False):
    """k nearest points (expanded-form distance, ascending, lowest index on ties) -> [B,S,k] int64.
    Stress-configuration grouping; the reference has no such function."""
    require_cuda(xyz, new_xyz)
    _f32(xyz, "xyz"), _f32(new_xyz, "new_xyz")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty(B, S, k, dtype=torch.long, device=xyz.device)
    dist = torch.empty(B, S, k, dtype=torch.float32, device=xyz.device) if return_dist else None
    check(_cabi.load().mpb_knn_group_f32(ptr(xyz), *_strides3(xyz), ptr(new_xyz), *_strides3(new_xyz), B, N, S, k,
                                         ptr(out), ptr(dist), stream_ptr()), "mpb_knn_group_f32")
    return (out, dist) if return_dist else out


# ---------------------------------------------------------------------------------------------
# a5  sample_and_group: fused gather + centre + concat
# ---------------------------------------------------------------------------------------------
class _GroupPoints(torch.autograd.Function):
    """out[b,s,k] = cat(xyz[b,idx] - new_xyz[b,s], feats[b,idx]) in one kernel (reference :133-138)."""

    @staticmethod
    def forward(ctx, xyz, feats, new_xyz, idx, ldo):
        B, N, _ = xyz.shape
        _, S, K = idx.shape
        D = 0 if feats is None else feats.shape[2]
        ldo = max(ldo, 3 + D)
        new_c = new_xyz.contiguous()
        idx_c = idx.contiguous()
        out = torch.empty(B, S, K, ldo, dtype=torch.float32, device=xyz.device)
        fs = _strides3(feats) if feats is not None else (0, 0, 0)
        check(_cabi.load().mpb_group_points_f32(ptr(xyz), *_strides3(xyz), ptr(feats), *fs, ptr(new_c), ptr(idx_c),
                                                B, N, S, K, D, ldo, ptr(out), stream_ptr()), "mpb_group_points_f32")
        ctx.save_for_backward(idx_c)
        ctx.dims = (B, N, S, K, D, ldo)
        ctx.has_feats = feats is not None
        return out

    @staticmethod
    def backward(ctx, go):
        (idx_c,) = ctx.saved_tensors
        B, N, S, K, D, ldo = ctx.dims
        go = go.contiguous()
        need_xyz, need_f, need_new = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and ctx.has_feats, ctx.needs_input_grad[2]
        gf = torch.zeros(B, N, D, dtype=torch.float32, device=go.device) if need_f else None
        gx = torch.zeros(B, N, 3, dtype=torch.float32, device=go.device) if need_xyz else None
        gn = torch.zeros(B, S, 3, dtype=torch.float32, device=go.device) if need_new else None
        if need_f or need_xyz or need_new:
            check(_cabi.load().mpb_group_points_bwd_f32(ptr(go), ldo, ptr(idx_c), B, N, S, K, D, ptr(gf), ptr(gx), ptr(gn),
                                                        stream_ptr()), "mpb_group_points_bwd_f32")
        return gx, gf, gn, None, None


def group_points(xyz, points, new_xyz, idx, ldo=0):
    """Fused form of reference :133-138.  `ldo` > 3+D zero-pads the last dimension (GEMM K padding)."""
    return _GroupPoints.apply(xyz, points, new_xyz, idx, ldo)


class _GroupPointsBF16(torch.autograd.Function):
    """Same gather + centre + concat, emitted directly as the tensor-core GEMM's A operand:
    bf16 rows [B*S*K, ldo] with ldo = pad64(3 + D).  Gradient flows to the features only (the
    coordinates are never differentiated on the MaskPlanner path)."""

    @staticmethod
    def forward(ctx, xyz, feats, new_xyz, idx, ldo):
        B, N, _ = xyz.shape
        _, S, K = idx.shape
        D = 0 if feats is None else feats.shape[2]
        new_c = new_xyz.contiguous()
        idx_c = idx.contiguous()
        out = torch.empty(B * S * K, ldo, dtype=torch.bfloat16, device=xyz.device)
        fs = _strides3(feats) if feats is not None else (0, 0, 0)
        check(_cabi.load().mpb_group_points_bf16(ptr(xyz), *_strides3(xyz), ptr(feats), *fs, ptr(new_c), ptr(idx_c),
                                                 B, N, S, K, D, ldo, ptr(out), stream_ptr()), "mpb_group_points_bf16")
        ctx.save_for_backward(idx_c)
        ctx.dims = (B, N, S, K, D, ldo)
        return out

    @staticmethod
    def backward(ctx, go):
        (idx_c,) = ctx.saved_tensors
        B, N, S, K, D, ldo = ctx.dims
        if not (ctx.needs_input_grad[1] and D > 0):
            return None, None, None, None, None
        go = go.contiguous()
        gf = torch.zeros(B, N, D, dtype=torch.float32, device=go.device)
        check(_cabi.load().mpb_group_points_bwd_bf16(ptr(go), ldo, ptr(idx_c), B, N, S, K, D, ptr(gf), stream_ptr()),
              "mpb_group_points_bwd_bf16")
        return None, gf, None, None, None


GROUP_ALL_KERNEL = os.environ.get("MPB_GROUP_ALL_KERNEL", "1") == "1"
_IDENTITY_GROUPS = {}


def _identity_group(B, N, device):
    """idx [B,1,N] = 0..N-1 for every sample (sample_and_group_all as a grouping with one all-points neighbourhood)."""
    key = (B, N, str(device))
    idx = _IDENTITY_GROUPS.get(key)
    if idx is None:
        idx = torch.arange(N, dtype=torch.long, device=device).expand(B, 1, N).contiguous()
        if not (idx.is_cuda and torch.cuda.is_current_stream_capturing()):   # a tensor born inside a capture lives in the graph's private pool
            _IDENTITY_GROUPS[key] = idx
    return idx


# Arithmetic of the shared MLP (reference :210-212); every mode runs the hand-written tcgen05 GEMMs + BatchNorm
# kernels of maskplanner_b200.shared_mlp:
#   "bf16"  bf16 activations/operands, fp32 accumulation and statistics                      tolerance rel 1e-2
#   "tf32"  fp32 activations, one TF32 pass per product (what cuDNN does for the reference
#           on a GPU with torch's default allow_tf32)                                         tolerance ~1e-3
#   "fp32"  fp32 activations, 3xTF32 (hi/lo split of both operands): the reference-precision
#           mode, what the parity fixtures frozen from the CPU reference are checked with     tolerance rel 1e-4
_MLP_PRECISION = "bf16"
PRECISIONS = ("bf16", "tf32", "fp32")


def set_mlp_precision(mode):
    """Select the shared-MLP arithmetic for PointNetSetAbstraction modules that do not override it."""
    global _MLP_PRECISION
    if mode not in PRECISIONS:
        raise ValueError("precision must be one of %s" % (PRECISIONS,))
    _MLP_PRECISION = mode


def get_mlp_precision():
    return _MLP_PRECISION


def sampling_plan(xyz, specs, seeds=None, out=None):
    """The index half of a chain of set-abstraction layers, which depends on the cloud alone (reference :130-132 of
    every layer: FPS -> centroid gather -> ball query; the next layer samples the previous layer's centroids).

    xyz [B,N,3]; specs = [(npoint, radius, nsample), ...]; seeds = per-layer FPS seed indices (None: drawn like the
    reference).  Returns [(fps_idx [B,S] i64, new_xyz [B,S,3], idx [B,S,K] i64), ...].  Nothing here needs the features,
    so a training loop can compute the plan of batch i+1 while batch i is still in flight (Trainer, pipeline_sampling).
    `out`: optional preallocated plan (same shapes) to write into (fixed addresses for CUDA-graph replay)."""
    plan = []
    cur = xyz
    for l, (npoint, radius, nsample) in enumerate(specs):
        seed = None if seeds is None else seeds[l]
        if out is not None:      # the index kernels write straight into the caller's buffers
            fps_idx = farthest_point_sample(cur, npoint, seed, out=out[l][0])
            new_xyz = out[l][1].copy_(index_points(cur, fps_idx))
            idx = query_ball_point(radius, nsample, cur, new_xyz, out=out[l][2])
        else:
            fps_idx = farthest_point_sample(cur, npoint, seed)
            new_xyz = index_points(cur, fps_idx)
            idx = query_ball_point(radius, nsample, cur, new_xyz)
        plan.append((fps_idx, new_xyz, idx))
        cur = new_xyz
    return plan


def empty_sampling_plan(B, specs, device):
    """Preallocated buffers with the shapes sampling_plan() returns."""
    return [(torch.zeros(B, S, dtype=torch.long, device=device), torch.zeros(B, S, 3, dtype=torch.float32, device=device),
             torch.zeros(B, S, K, dtype=torch.long, device=device)) for (S, _r, K) in specs]


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, full_points=None, seed_idx=None):
    """Reference :112-148.  xyz [B,N,3], points [B,N,D]|None -> new_xyz [B,S,3], new_points [B,S,K,3+D]
    (or [B,S,K,C_full] with `full_points`; 4-tuple with `returnfps`)."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint, seed_idx)       # :130
    new_xyz = index_points(xyz, fps_idx)                          # :131
    idx = query_ball_point(radius, nsample, xyz, new_xyz)         # :132
    if points is not None:
        new_points = group_points(xyz, points, new_xyz, idx)      # :133-138 fused
    elif full_points is not None:
        new_points = index_points(full_points, idx)               # :139-141
    else:
        new_points = group_points(xyz, None, new_xyz, idx)        # :143
    if returnfps:
        return new_xyz, new_points, index_points(xyz, idx), fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """Reference :151-168: new_xyz = zeros [B,1,3]; new_points = cat(xyz, points) as one group [B,1,N,3+D]."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device)
    grouped_xyz = xyz.reshape(B, 1, N, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz, points.reshape(B, 1, N, -1)], dim=-1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points


# ---------------------------------------------------------------------------------------------
# a7  PointNetSetAbstraction
# ---------------------------------------------------------------------------------------------
class PointNetSetAbstraction(nn.Module):
    """Reference :171-216.  Same constructor, attributes and parameter names/shapes
    (``mlp_convs.{i}.weight [Cout,Cin,1,1]``, ``.bias``, ``mlp_bns.{i}.*``) so state_dicts interchange.

    forward(xyz [B,3,N], points [B,D,N]|None, full_points=None) -> (new_xyz [B,3,S], new_points [B,D',S])
    """

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.group_all = group_all
        # The reference returns a contiguous [B,C',S] tensor.  Internally features are produced
        # position-major ([B,S,C']); a caller that immediately feeds the next SA layer (which
        # permutes back, reference :196-198) can set this to False and skip the transpose copy.
        self.contiguous_output = True
        self.precision = None   # None -> module-level default (set_mlp_precision); or "bf16" / "tf32" / "fp32"

    def forward(self, xyz, points, full_points=None, seed_idx=None, sampling=None):
        """`sampling` (extension): a precomputed (fps_idx, new_xyz, idx) triple of this layer from sampling_plan()."""
        xyz = xyz.permute(0, 2, 1)                                                      # :196
        if points is not None:
            points = points.permute(0, 2, 1)
        if full_points is not None:
            full_points = full_points.permute(0, 2, 1)
        if self.training or not torch.is_grad_enabled():
            return self._forward_tensor_core(xyz, points, full_points, seed_idx, sampling)
        assert sampling is None, "precomputed sampling is only wired into the tensor-core path"
        # eval mode WITH autograd (not on the training or inference path): stock torch ops below
        if self.group_all:
            new_xyz, new_points = sample_and_group_all(xyz, points)                     # :203
        else:
            new_xyz, new_points = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points,
                                                   full_points=full_points, seed_idx=seed_idx)   # :205
        new_points = self._shared_mlp_max(new_points)                                   # :208-214
        return new_xyz.permute(0, 2, 1), new_points

    def _shared_mlp_max(self, grouped):
        """grouped [B,S,K,C] -> [B,C',S]: per-layer relu(bn(conv1x1)) then max over K (reference :208-214).

        The grouped tensor is already position-major, i.e. the channels-last image [B,C,S,K]; a 1x1
        conv, BatchNorm and the max over K do not care about the order of the two spatial axes."""
        x = grouped  # [B,S,K,C], rows = neighbourhood positions
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            x = F.linear(x, conv.weight.view(conv.out_channels, -1), conv.bias)     # 1x1 conv == row-wise GEMM (strict fp32)
            x = F.relu(bn(x.permute(0, 3, 1, 2))).permute(0, 2, 3, 1)               # BN over all B*S*K rows per channel
        out = torch.max(x, 2)[0].permute(0, 2, 1)                                    # [B,S,C'] -> [B,C',S]
        return out.contiguous() if self.contiguous_output else out

    def _forward_tensor_core(self, xyz, points, full_points, seed_idx, sampling=None):
        """Reference :203-215 with the grouped tensor produced directly as GEMM rows (bf16 or fp32, by precision)
        and the MLP on tcgen05 (maskplanner_b200.shared_mlp).  xyz [B,N,3], points [B,N,D]|None (position-major)."""
        from .shared_mlp import MODES, narrow_rows_supported, pad64, shared_mlp_max
        require_cuda(xyz)
        mode = self.precision or _MLP_PRECISION
        row_dtype = MODES[mode][2]
        B, N, C = xyz.shape
        a0 = None
        if self.group_all:                                                              # :151-168
            new_xyz = torch.zeros(B, 1, C, device=xyz.device)
            S, K = 1, N
            if GROUP_ALL_KERNEL and mode == "bf16" and points is not None and C == 3 and not xyz.requires_grad:
                # one group = every point, centred on the origin: the grouping kernel with the identity index emits the
                # padded bf16 rows in ONE launch (torch: cat + pad (fill + copy) + cast, and three copies in the backward)
                a0 = _GroupPointsBF16.apply(xyz, points, new_xyz, _identity_group(B, N, xyz.device), pad64(3 + points.shape[2]))
                rows = None
            else:
                rows = xyz if points is None else torch.cat([xyz, points], dim=-1)
        else:
            S, K = self.npoint, self.nsample
            if sampling is not None:
                fps_idx, new_xyz, idx = sampling
                assert tuple(idx.shape) == (B, S, K), "sampling plan does not match this layer"
            else:
                fps_idx = farthest_point_sample(xyz, S, seed_idx)                       # :130
                new_xyz = index_points(xyz, fps_idx)                                    # :131
                idx = query_ball_point(self.radius, K, xyz, new_xyz)                    # :132
            rows = index_points(full_points, idx) if (points is None and full_points is not None) else None
        xyz_last = False
        if a0 is not None:
            xyz_last = True
        elif rows is not None:   # group-all / full_points: plain rows, padded (and rounded to bf16 in that mode)
            w = rows.shape[-1]
            a0 = F.pad(rows.reshape(B * S * K, w), (0, pad64(w) - w)).to(row_dtype)
        elif narrow_rows_supported(points, K) and len(self.mlp_convs) >= 2:
            # <= 8 channels per grouped row (SA1): the first layer gathers its rows on the fly, nothing is materialised
            a0 = (xyz, points, new_xyz.contiguous(), idx.contiguous())                  # :133-138 fused into :210
            xyz_last = True
        else:
            D = 0 if points is None else points.shape[2]
            if mode == "bf16":   # features first, centred xyz last: aligned 16-byte feature copies
                a0 = _GroupPointsBF16.apply(xyz, points, new_xyz, idx, pad64(3 + D))    # :133-138
                xyz_last = True
            else:                # fp32 rows in the reference's own column order (xyz first), zero padded
                a0 = group_points(xyz, points, new_xyz, idx, ldo=pad64(3 + D)).view(B * S * K, pad64(3 + D))
        pooled = shared_mlp_max(a0, K, self.mlp_convs, self.mlp_bns, self.training, xyz_last=xyz_last, mode=mode)   # :208-214
        out = pooled.view(B, S, -1).permute(0, 2, 1)
        return new_xyz.permute(0, 2, 1), (out.contiguous() if self.contiguous_output else out)


# ---------------------------------------------------------------------------------------------
# Remaining surface of the reference file: multi-scale grouping and feature propagation.
# Neither is on the MaskPlanner path (Msg is unused, FP only appears in a NotImplementedError class), but
# other backbones of the reference import them; they reuse the same kernels.
# ---------------------------------------------------------------------------------------------
def _torch_mlp_max(rows, convs, bns):
    """rows [B,S,K,C] -> [B,C',S] with strict-fp32 library GEMMs (reference :267-271)."""
    x = rows
    for conv, bn in zip(convs, bns):
        x = F.linear(x, conv.weight.view(conv.out_channels, -1), conv.bias)
        x = F.relu(bn(x.permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    return torch.max(x, 2)[0].permute(0, 2, 1).contiguous()


class PointNetSetAbstractionMsg(nn.Module):
    """Reference :219-276: one FPS, then per radius ball query -> [feats | centred xyz] -> MLP -> max, concatenated.
    Same constructor, attributes and parameter names (``conv_blocks.{i}.{j}``, ``bn_blocks.{i}.{j}``)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint = npoint
        self.radius_list = radius_list
        self.nsample_list = nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for mlp in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last_channel = in_channel + 3
            for out_channel in mlp:
                convs.append(nn.Conv2d(last_channel, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last_channel = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
        self.precision = None

    def forward(self, xyz, points, seed_idx=None):
        from .shared_mlp import pad64, shared_mlp_max
        xyz = xyz.permute(0, 2, 1)
        if points is not None:
            points = points.permute(0, 2, 1)
        B, N, C = xyz.shape
        S = self.npoint
        new_xyz = index_points(xyz, farthest_point_sample(xyz, S, seed_idx))          # :253
        tensor_core = self.training or not torch.is_grad_enabled()
        mode = self.precision or _MLP_PRECISION
        D = 0 if points is None else points.shape[2]
        outs = []
        for i, radius in enumerate(self.radius_list):
            K = self.nsample_list[i]
            idx = query_ball_point(radius, K, xyz, new_xyz)                           # :257
            if tensor_core:
                # the reference concatenates [feats, centred xyz] here (:262-263): exactly the bf16 row layout
                if mode == "bf16":
                    a0 = _GroupPointsBF16.apply(xyz, points, new_xyz, idx, pad64(3 + D))
                else:   # fp32 rows [feats | centred xyz | 0] in the reference's order
                    centred = group_points(xyz, None, new_xyz, idx)
                    rows = centred if points is None else torch.cat([index_points(points, idx), centred], dim=-1)
                    a0 = F.pad(rows.reshape(B * S * K, 3 + D), (0, pad64(3 + D) - (3 + D)))
                pooled = shared_mlp_max(a0, K, self.conv_blocks[i], self.bn_blocks[i], self.training, xyz_last=False, mode=mode)
                outs.append(pooled.view(B, S, -1).permute(0, 2, 1))
            else:
                centred = group_points(xyz, None, new_xyz, idx)                       # :258-259
                rows = centred if points is None else torch.cat([index_points(points, idx), centred], dim=-1)
                outs.append(_torch_mlp_max(rows, self.conv_blocks[i], self.bn_blocks[i]))
        return new_xyz.permute(0, 2, 1), torch.cat(outs, dim=1)                       # :274-276


class PointNetFeaturePropagation(nn.Module):
    """Reference :279-329: inverse-distance interpolation from the 3 nearest sampled points, concat, Conv1d MLP.
    The 3-NN search is the kNN kernel (expanded-form distances, ascending); the MLP over [B,C,N] stays a library op."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last_channel = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        xyz1 = xyz1.permute(0, 2, 1)
        xyz2 = xyz2.permute(0, 2, 1)
        points2 = points2.permute(0, 2, 1)
        B, N, C = xyz1.shape
        S = xyz2.shape[1]
        if S == 1:
            interpolated = points2.repeat(1, N, 1)                                     # :307-308
        else:
            k = min(3, S)
            idx, dists = knn_group(k, xyz2, xyz1, return_dist=True)                     # :310-312 (sort + first 3)
            recip = 1.0 / (dists + 1e-8)                                                # :314
            weight = recip / torch.sum(recip, dim=2, keepdim=True)                      # :315-316
            interpolated = torch.sum(index_points(points2, idx) * weight.view(B, N, k, 1), dim=2)   # :317
        if points1 is not None:
            new_points = torch.cat([points1.permute(0, 2, 1), interpolated], dim=-1)    # :319-321
        else:
            new_points = interpolated
        x = new_points                                                                  # [B,N,C] rows
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):                              # :326-328
            x = F.linear(x, conv.weight.squeeze(-1), conv.bias)                         # Conv1d(k=1) as a strict-fp32 row GEMM
            x = F.relu(bn(x.permute(0, 2, 1))).permute(0, 2, 1)
        return x.permute(0, 2, 1).contiguous()
