"""The MaskPlanner regressor the SA encoder drops into (caller side of the hot path).

Mirror of the reference's ``PointNet2Regressor_StrokeMasks`` (models/pointnet2_cls_ssg.py:233-344):
same constructor arguments, same sub-module / parameter names (``sa1..sa3``, ``fc1..fc3``, ``bn1``,
``bn2``, ``fc_normals``, ``sm_fc1..3``, ``sm_bn1..2``, ``mask_conf_out``, ``seg_conf_*``) so reference
state_dicts load unchanged, same output tuple.  The three set-abstraction layers are the B200
drop-ins from ``maskplanner_b200.pointnet2_utils``; the heads keep their stock ``nn.Linear`` / ``BatchNorm1d`` modules as
parameter containers, but in the MaskPlanner configuration their arithmetic runs as one autograd node over hand-written
kernels (``maskplanner_b200.heads``: weight-streaming tcgen05 GEMMs + row-local BatchNorm1d/ReLU/dropout kernels,
SURVEY.md 8f-3).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import synthetic
from .pointnet2_utils import PointNetSetAbstraction
from .streams import Fork


class PointNet2Regressor_StrokeMasks(nn.Module):
    def __init__(self, outdim=3, outdim_orient=3, weight_orient=1., normal_channel=False, out_vectors=1500,
                 hidden_size=(1024, 1024), inputdim=None, pred_stroke_masks=False, n_stroke_masks=None,
                 mask_confidence_scores=False, segment_confidence_scores=False):
        super().__init__()
        self.outdim = outdim
        self.outdim_orient = outdim_orient
        self.out_vectors = out_vectors
        self.weight_orient = weight_orient
        self.pred_stroke_masks = pred_stroke_masks
        self.n_stroke_masks = n_stroke_masks
        self.mask_confidence_scores = mask_confidence_scores
        self.segment_confidence_scores = segment_confidence_scores
        self.normal_channel = normal_channel
        in_channel = inputdim if inputdim is not None else (6 if normal_channel else 3)
        # encoder geometry: pointnet2_cls_ssg.py:266-268
        self.sa1 = PointNetSetAbstraction(npoint=512, radius=0.2, nsample=32, in_channel=in_channel, mlp=[64, 64, 128], group_all=False)
        self.sa2 = PointNetSetAbstraction(npoint=128, radius=0.4, nsample=64, in_channel=128 + 3, mlp=[128, 128, 256], group_all=False)
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=256 + 3, mlp=[256, 512, 1024], group_all=True)
        # sa1/sa2 feed the next SA layer directly: skip the [B,S,C] -> [B,C,S] transpose copy
        self.sa1.contiguous_output = False
        self.sa2.contiguous_output = False
        h0, h1 = hidden_size
        self.fc1 = nn.Linear(1024, h0)
        self.fc2 = nn.Linear(h0, h1)
        self.fc3 = nn.Linear(h1, out_vectors * outdim)
        self.dropout = nn.Dropout(p=0.3)
        self.bn1 = nn.BatchNorm1d(h0)
        self.bn2 = nn.BatchNorm1d(h1)
        if outdim_orient > 0:
            self.fc_normals = nn.Linear(h1, out_vectors * outdim_orient)
            self.tanh = nn.Tanh()
        if segment_confidence_scores:
            self.seg_conf_fc1 = nn.Linear(1024, h0)
            self.seg_conf_fc2 = nn.Linear(h0, h1)
            self.seg_conf_out = nn.Linear(h1, out_vectors)
        if pred_stroke_masks:
            self.sm_fc1 = nn.Linear(1024, h0)
            self.sm_fc2 = nn.Linear(h0, h1)
            self.sm_fc3 = nn.Linear(h1, out_vectors * n_stroke_masks)
            self.sm_bn1 = nn.BatchNorm1d(h0)
            self.sm_bn2 = nn.BatchNorm1d(h1)
            if mask_confidence_scores:
                self.mask_conf_out = nn.Linear(h1, n_stroke_masks)
        import os
        self.fused_heads = os.environ.get("MPB_FUSED_HEADS", "1") == "1"
        self._heads = None

    def sampling_specs(self):
        return [(self.sa1.npoint, self.sa1.radius, self.sa1.nsample), (self.sa2.npoint, self.sa2.radius, self.sa2.nsample)]

    def sampling_plan(self, xyz, fps_seeds=None, out=None):
        """FPS / centroid / ball-query indices of sa1 and sa2 for a cloud xyz [B,3,N] (they depend on the coordinates only)."""
        from .pointnet2_utils import sampling_plan
        return sampling_plan(xyz[:, :3, :].permute(0, 2, 1), self.sampling_specs(), fps_seeds, out=out)

    def encode(self, xyz, fps_seeds=None, plan=None):
        """xyz [B,3,N] -> global feature [B,1024] (pointnet2_cls_ssg.py:297-309)."""
        B = xyz.shape[0]
        norm = xyz[:, 3:, :] if self.normal_channel else None
        if self.normal_channel:
            xyz = xyz[:, :3, :]
        s1, s2 = fps_seeds if fps_seeds is not None else (None, None)
        p1, p2 = plan if plan is not None else (None, None)
        l1_xyz, l1_points = self.sa1(xyz, norm, seed_idx=s1, sampling=p1)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, seed_idx=s2, sampling=p2)
        _, l3_points = self.sa3(l2_xyz, l2_points)
        return l3_points.reshape(B, 1024)

    def forward(self, xyz, fps_seeds=None, plan=None, after_encode=None):
        """plan: precomputed sampling_plan() of this cloud; after_encode: optional callable run between the encoder and
        the heads (the Trainer forks the NEXT batch's sampling plan onto a side stream there)."""
        B = xyz.shape[0]
        feat = self.encode(xyz, fps_seeds, plan)
        if after_encode is not None:
            after_encode()
        if self.fused_heads and feat.is_cuda and B <= 128:
            # MaskPlanner configuration: both heads as one autograd node over hand-written kernels (maskplanner_b200.heads)
            from .heads import FusedHeads
            from .pointnet2_utils import get_mlp_precision
            if self._heads is None:
                if not FusedHeads.supported(self):
                    self.fused_heads = False
                    return self._forward_heads_torch(feat, B)
                self._heads = FusedHeads(self)
            out, masks, mask_scores = self._heads(feat, self.sa1.precision or get_mlp_precision())
            return out, masks, mask_scores, None
        return self._forward_heads_torch(feat, B)

    def _forward_heads_torch(self, feat, B):
        """The heads as stock torch modules (library GEMMs): configurations the fused path does not cover
        (per-segment confidence scores, no stroke masks, batches above 128) and A/B comparisons in the tests."""
        h = self.dropout(F.relu(self.bn1(self.fc1(feat))))                       # :310
        h = self.dropout(F.relu(self.bn2(self.fc2(h))))                          # :311
        seg = self.fc3(h)                                                        # :312

        seg_conf = None
        if self.segment_confidence_scores:                                       # :314-319
            c = self.dropout(F.relu(self.seg_conf_fc1(feat)))
            c = self.dropout(F.relu(self.seg_conf_fc2(c)))
            seg_conf = torch.sigmoid(self.seg_conf_out(c))

        masks, mask_scores = None, None
        fork = None
        if self.pred_stroke_masks:                                               # :321-329
            # the stroke-mask head only shares `feat` with the pose head: same host order (dropout RNG consumption as
            # in the reference), but issued on a side stream so the two chains of M = batch GEMMs overlap
            with Fork(feat) as fork:
                m = self.dropout(F.relu(self.sm_bn1(self.sm_fc1(feat))))
                m = self.dropout(F.relu(self.sm_bn2(self.sm_fc2(m))))
                masks = self.sm_fc3(m).view(B, self.n_stroke_masks, -1)
                if self.mask_confidence_scores:
                    mask_scores = self.mask_conf_out(m)

        if self.outdim_orient > 0:                                               # :332-339
            nrm = self.tanh(self.fc_normals(h)).view(B, -1, 3)
            nrm = F.normalize(nrm, dim=-1) * self.weight_orient
            out = torch.cat((seg.view(B, -1, 3), nrm), dim=-1).view(B, self.out_vectors, -1)
        else:
            out = seg.view(B, self.out_vectors, self.outdim)
        if fork is not None:
            fork.join(masks, mask_scores)
        return out, masks, mask_scores, seg_conf


def maskplanner_model(category="windows_v2", per_segment_confidence=False):
    """get_model(which='pointnet2_strokemasks', io_type='MaskPlanner') with the resolved values of
    config=[maskplanner,<category>,longx_v2] (models/__init__.py:111-122, :297-325): 6-d poses
    (3 position + 3 orientation dims), lambda = 4 -> 12 translational + 12 orientation dims per segment."""
    cfg = synthetic.CATEGORIES[category]
    lam = synthetic.LAMBDA_POINTS
    return PointNet2Regressor_StrokeMasks(
        out_vectors=synthetic.out_vectors(cfg["n_pred_traj_points"]), outdim=3 * lam, outdim_orient=3 * lam,
        weight_orient=synthetic.WEIGHT_ORIENT, hidden_size=(1024, 1024), pred_stroke_masks=True,
        n_stroke_masks=cfg["max_n_strokes"], mask_confidence_scores=True, segment_confidence_scores=per_segment_confidence)
