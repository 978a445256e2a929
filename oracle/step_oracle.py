"""CPU restatement of the MaskPlanner model + loss + training-step body -- TEST INFRASTRUCTURE.

Restates, with plain CPU torch ops and the reference's own per-sample structure:
  * models/pointnet2_cls_ssg.py:233-344   PointNet2Regressor_StrokeMasks (over oracle SA layers)
  * loss_handler.py:596-666, 816-967      asymm_v6 chamfer terms + Hungarian-matched stroke-mask loss
  * train_maskplanner.py:183-227          the step body (zero_grad, forward, loss, backward, Adam step)
It is the checker for tests/ and the timed CPU arm of bench.py (`--impl reference`, `cpu_baseline`).
Validated against the imported reference model by oracle/make_golden.py / tests (when the tree is mounted).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from . import torch_oracle as T


class Regressor(nn.Module):
    """models/pointnet2_cls_ssg.py:233-344 with identical parameter names."""

    def __init__(self, out_vectors, outdim=12, outdim_orient=12, weight_orient=0.25, hidden_size=(1024, 1024),
                 n_stroke_masks=22):
        super().__init__()
        self.out_vectors, self.outdim, self.outdim_orient = out_vectors, outdim, outdim_orient
        self.weight_orient, self.n_stroke_masks = weight_orient, n_stroke_masks
        self.sa1 = T.PointNetSetAbstraction(512, 0.2, 32, 3, [64, 64, 128], False)             # :266
        self.sa2 = T.PointNetSetAbstraction(128, 0.4, 64, 128 + 3, [128, 128, 256], False)     # :267
        self.sa3 = T.PointNetSetAbstraction(None, None, None, 256 + 3, [256, 512, 1024], True)  # :268
        h0, h1 = hidden_size
        self.fc1, self.fc2, self.fc3 = nn.Linear(1024, h0), nn.Linear(h0, h1), nn.Linear(h1, out_vectors * outdim)
        self.dropout = nn.Dropout(p=0.3)
        self.bn1, self.bn2 = nn.BatchNorm1d(h0), nn.BatchNorm1d(h1)
        self.fc_normals = nn.Linear(h1, out_vectors * outdim_orient)
        self.sm_fc1, self.sm_fc2 = nn.Linear(1024, h0), nn.Linear(h0, h1)
        self.sm_fc3 = nn.Linear(h1, out_vectors * n_stroke_masks)
        self.sm_bn1, self.sm_bn2 = nn.BatchNorm1d(h0), nn.BatchNorm1d(h1)
        self.mask_conf_out = nn.Linear(h1, n_stroke_masks)

    def encode(self, xyz, fps_seeds=None):
        s1, s2 = fps_seeds if fps_seeds is not None else (None, None)
        l1_xyz, l1_points = self.sa1(xyz, None, seed_idx=s1)                                  # :305
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, seed_idx=s2)                           # :306
        _, l3_points = self.sa3(l2_xyz, l2_points)                                             # :307
        return l3_points.view(xyz.shape[0], 1024)                                              # :309

    def forward(self, xyz, fps_seeds=None):
        B = xyz.shape[0]
        g = self.encode(xyz, fps_seeds)
        x = self.dropout(F.relu(self.bn1(self.fc1(g))))                                        # :310
        final = self.dropout(F.relu(self.bn2(self.fc2(x))))                                    # :311
        seg = self.fc3(final)                                                                  # :312
        m = self.dropout(F.relu(self.sm_bn1(self.sm_fc1(g))))                                  # :323
        m = self.dropout(F.relu(self.sm_bn2(self.sm_fc2(m))))                                  # :324
        masks = self.sm_fc3(m).view(B, self.n_stroke_masks, -1)                                # :325-326
        scores = self.mask_conf_out(m)                                                         # :329
        nrm = torch.tanh(self.fc_normals(final)).view(B, -1, 3)                                # :333-334
        nrm = F.normalize(nrm, dim=-1) * self.weight_orient                                    # :335-336
        out = torch.cat((seg.view(B, -1, 3), nrm), dim=-1).view(B, self.out_vectors, -1)       # :337-339
        return out, masks, scores, None


def _masks_from_ids(ids):
    """loss_handler.py:938-967: one binary mask per distinct stroke id (ascending), id -1 skipped."""
    rows = [(ids == s).int() for s in torch.unique(ids) if s != -1]
    return torch.stack(rows)


def stroke_masks_loss(match, pred_masks, scores, stroke_ids, w_masks=1.0, w_conf=100.0, no_stroke_weight=1.0):
    """loss_handler.py:816-935, smooth_targets=False."""
    tgt_ids = stroke_ids.to(pred_masks.device).gather(dim=1, index=match)                      # :838 (`stroke_ids.cuda()`)
    tgt_masks = [_masks_from_ids(t) for t in tgt_ids]                                          # :847-848
    assert not torch.any(tgt_ids == -1)                                                        # :852
    B, P, S = pred_masks.shape
    pairs = []
    with torch.no_grad():
        for pm, tm in zip(pred_masks, tgt_masks):                                              # :862
            nt = tm.shape[0]
            a = pm.repeat_interleave(nt, dim=0)                                                # :867
            b = tm.repeat(P, 1).to(pm.dtype)                                                  # :868
            cost = F.binary_cross_entropy_with_logits(a, b, reduction="none").sum(-1).view(P, nt)   # :871-873
            pairs.append(linear_sum_assignment(cost.cpu().numpy()))                            # :875 (per-sample .cpu())
    dev = pred_masks.device
    bi = torch.cat([torch.full((len(r),), i, dtype=torch.int64) for i, (r, _) in enumerate(pairs)]).to(dev)
    pi = torch.cat([torch.as_tensor(r, dtype=torch.int64) for r, _ in pairs]).to(dev)
    ti = torch.cat([torch.as_tensor(c, dtype=torch.int64) for _, c in pairs]).to(dev)
    matched_pred = pred_masks[bi, pi]                                                          # :886
    matched_tgt = torch.stack([tgt_masks[b][t] for b, t in zip(bi.tolist(), ti.tolist())]).to(pred_masks.dtype)    # :896-902
    mask_loss = F.binary_cross_entropy_with_logits(matched_pred, matched_tgt, reduction="none").sum(-1).mean()   # :906
    tgt_scores = torch.zeros(scores.shape, dtype=scores.dtype, device=dev)                     # :920-921
    tgt_scores[bi, pi] = 1.0
    w = no_stroke_weight * torch.ones(scores.shape, dtype=scores.dtype, device=dev)            # :924-925
    w[bi, pi] = 1.0
    conf = F.binary_cross_entropy_with_logits(scores, tgt_scores, reduction="none", weight=w).mean()   # :930
    return w_masks * mask_loss + w_conf * conf, (bi, pi, ti)


def asymm_v6_loss(y_pred, y, pred_masks, scores, stroke_ids, traj_as_pc, w=(1.0, 100.0, 0.01), w_masks=1.0, w_conf=100.0,
                  return_terms=False):
    """loss_handler.py:596-666."""
    d1, _, match, _ = T.chamfer_distance(y_pred, y, padded=True, asymmetric=True, return_matching=True,
                                         point_reduction=None, batch_reduction=None)           # :604-610
    t1 = 100 * d1.mean()                                                                       # :611
    B = y_pred.shape[0]
    t2 = 100 * T.chamfer_distance(y_pred.reshape(B, -1, 6), traj_as_pc, padded=True, reverse_asymmetric=True)[0]   # :631-636
    t3 = 100 * T.chamfer_distance(y_pred, y, padded=True, reverse_asymmetric=True)[0]          # :642-645
    masks, assignment = stroke_masks_loss(match, pred_masks, scores, stroke_ids, w_masks, w_conf)   # :650-656
    loss = w[0] * t1 + w[1] * t2 + w[2] * t3 + masks                                           # :660-664
    if return_terms:
        return loss, dict(asymm_segment=t1.detach(), reverse_point=t2.detach(), reverse_segment=t3.detach(),
                          masks=masks.detach(), match=match, assignment=assignment)
    return loss


def train_step(model, opt, batch, fps_seeds=None):
    """train_maskplanner.py:183-227 for one batch (dict from maskplanner_b200.synthetic.make_batch)."""
    model.zero_grad()
    dev = next(model.parameters()).device
    cloud = batch["point_cloud"].permute(0, 2, 1).float().to(dev)                              # :207-208
    pred, masks, scores, _ = model(cloud, fps_seeds)                                           # :210
    loss = asymm_v6_loss(pred, batch["traj"].float().clone().to(dev), masks, scores, batch["stroke_ids"].to(dev),
                         batch["traj_as_pc"].float().clone().to(dev))                          # :212-218 (:629 `.to('cuda')`)
    loss.backward()                                                                            # :220
    opt.step()                                                                                 # :221
    return float(loss.item())                                                                  # :223
