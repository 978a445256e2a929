"""The MaskPlanner training loss around the chamfer drop-in (caller side of the hot path).

``asymm_v6_chamfer_with_stroke_masks`` mirrors ``LossHandler.get_asymm_v6_chamfer_with_stroke_masks``
(loss_handler.py:596-666) with the resolved weights of config=[maskplanner,<cat>,longx_v2]
(configs/maskplanner/asymm_chamfer_v9.yaml:10-14):

    1. 100 * mean over (B, out_segments) of the pred->GT segment chamfer, no reduction, with matching  (:604-611)
    2. 100 * GT->pred POINT chamfer on the poses (reverse_asymmetric, mean/mean)                          (:631-636)
    3. 100 * GT->pred SEGMENT chamfer (reverse_asymmetric, mean/mean)                                    (:642-645)
    4. stroke-mask loss: Hungarian-matched BCE on masks + weighted BCE on mask confidences              (:816-935)

Three implementations, selected by ``fused``:

* ``fused=True`` (default, the training path): ONE autograd node over the fused kernels of ``csrc/loss.cu``
  (``_FusedAsymmV6``): a length scan of both ground-truth tensors, the segment nearest-neighbour search in both
  directions (terms 1 + 3 share it) with the pose search (term 2) on a side stream, all B x P x T mask cost matrices,
  the per-sample Hungarian assignment on the device (``mpb_lap_f32``, one warp per sample), a single-CTA loss-value
  kernel off the critical path, and three backward launches.  No host synchronisation, CUDA-graph capturable, the five
  schedulable weights read from device memory.
* ``fused="nn"``: terms 1 and 3 from one nearest-neighbour launch, everything around it stock torch ops (round 1's path;
  kept for A/B runs and as a second implementation the parity tests compare against).
* ``fused=False``: the reference's three ``chamfer_distance`` calls literally, with ``matcher="host"`` the reference's own
  scipy solver from one device->host copy (parity tests).

The mask loss needs a linear assignment per sample (<= 22 x 22).  The reference builds each cost matrix with Python loops
and moves it to the host one sample at a time (:860-875).  BCE-with-logits against a binary target is
softplus(x) - x*y, so cost[p,t] = sum_i softplus(x[p,i]) - sum_{i in stroke t} x[p,i] for every pair at once, and the
matched BCE of the loss (:886-906) is exactly the selected cost entry.
"""
import os
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

from . import pytorch3d_chamfer as CH
from .streams import Fork


@dataclass
class LossConfig:
    """Resolved loss weights (asymm_chamfer_v9.yaml:10-14, default.yaml explicit_* entries)."""
    weight_asymm_segment_chamfer: float = 1.0
    weight_reverse_asymm_point_chamfer: float = 100.0
    weight_reverse_asymm_segment_chamfer: float = 0.01
    explicit_weight_stroke_masks: float = 1.0            # target value after delayMasksLoss (delayMasksLoss.yaml:5)
    explicit_weight_stroke_masks_confidence: float = 100.0
    explicit_no_stroke_weight: float = 1.0
    pose_dim: int = 6                                    # get_dim_traj_points(['orientnorm'])
    # Data parallel only: normalise the matched-pair BCE (:906, `.mean()` over all matched pairs of the batch) by the
    # GLOBAL pair count (one scalar all-reduce) so that the rank-averaged gradient equals the single-process one;
    # False = each rank takes the mean over its own shard (plain DDP behaviour).
    mask_loss_global_mean: bool = False


def chamfer_terms_13(y_pred, y, cfg, fused=True):
    """Terms 1 and 3 (loss_handler.py:604-645).  Returns (term1, term3, nn_distance [B,P1], pred_to_gt_match [B,P1])."""
    B = y_pred.shape[0]
    if fused:
        x = y_pred.float().contiguous()
        yy = y.float().contiguous()
        x_len = torch.full((B,), x.shape[1], dtype=torch.int64, device=x.device)
        y_len = CH.padded_lengths(yy, None, False)
        d_x, match, d_y, _ = CH._ChamferNN.apply(x, yy, x_len, y_len, True, True)
        term1 = 100 * d_x.mean()
        term3 = 100 * (d_y.sum(1) / y_len).sum() / B
    else:
        d_x, _, match, _ = CH.chamfer_distance(y_pred, y, padded=True, asymmetric=True, return_matching=True,
                                               point_reduction=None, batch_reduction=None)
        term1 = 100 * d_x.mean()
        term3 = 100 * CH.chamfer_distance(y_pred, y, padded=True, reverse_asymmetric=True)[0]
    return term1, term3, d_x, match


def chamfer_term2(y_pred, traj_as_pc, cfg):
    """Term 2: ground-truth poses -> predicted poses (loss_handler.py:628-636)."""
    points_pred = y_pred.reshape(y_pred.shape[0], -1, cfg.pose_dim)
    return 100 * CH.chamfer_distance(points_pred, traj_as_pc, padded=True, reverse_asymmetric=True)[0]


def chamfer_terms(y_pred, y, traj_as_pc, cfg, fused=True):
    """Terms 1-3.  Returns (term1, term2, term3, nn_distance [B,P1], pred_to_gt_match [B,P1] int64)."""
    term1, term3, d_x, match = chamfer_terms_13(y_pred, y, cfg, fused=fused)
    term2 = chamfer_term2(y_pred, traj_as_pc, cfg)
    return term1, term2, term3, d_x, match


def mask_cost_matrices(pred_stroke_masks, target_ids, n_ids):
    """cost[b,p,t] = sum_i BCEWithLogits(pred[b,p,i], [target_ids[b,i] == t]) and present[b,t] (stroke t
    owns at least one predicted segment).  Equivalent to loss_handler.py:865-873 for every sample at once."""
    classes = torch.arange(n_ids, device=target_ids.device)
    onehot = (target_ids[:, :, None] == classes[None, None, :]).to(pred_stroke_masks.dtype)   # [B, S, T] (no value check => no sync)
    sp = F.softplus(pred_stroke_masks).sum(-1, keepdim=True)                     # [B, P, 1]
    cost = sp - torch.bmm(pred_stroke_masks, onehot)                             # [B, P, T]
    present = onehot.sum(1) > 0                                                  # [B, T]
    return cost, present, onehot


def hungarian_host(cost, present):
    """scipy linear_sum_assignment per sample on the present columns (loss_handler.py:875).
    Returns (batch_idx, pred_idx, target_id) int64 CPU tensors, concatenated over the batch."""
    from scipy.optimize import linear_sum_assignment
    c = cost.detach().cpu().numpy()
    pres = present.cpu().numpy()
    bi, pi, ti = [], [], []
    for b in range(c.shape[0]):
        cols = np.nonzero(pres[b])[0]
        r, k = linear_sum_assignment(c[b][:, cols])
        bi.append(np.full(len(r), b, dtype=np.int64))
        pi.append(r.astype(np.int64))
        ti.append(cols[k].astype(np.int64))
    return (torch.from_numpy(np.concatenate(bi)), torch.from_numpy(np.concatenate(pi)), torch.from_numpy(np.concatenate(ti)))


def hungarian_device(cost, present):
    """The same assignment on the device: one warp per sample (mpb_lap_f32), no host synchronisation.
    Returns row [B, T] int64: predicted mask matched to target stroke t, -1 where the stroke is absent."""
    from . import _cabi
    B, P, T = cost.shape
    cost = cost.detach().float().contiguous()
    pres = present.to(torch.uint8).contiguous()
    row = torch.empty(B, T, dtype=torch.int64, device=cost.device)
    _cabi.check(_cabi.load().mpb_lap_f32(_cabi.ptr(cost), _cabi.ptr(pres), B, P, T, _cabi.ptr(row), _cabi.stream_ptr()), "mpb_lap_f32")
    return row


def validate_stroke_ids(stroke_ids, n_pred_masks):
    """Host-side precondition of the fixed-shape mask loss: every real stroke id lies in [0, n_pred_masks) and the
    padding id is -1 (utils/dataset/paintnet_ODv1.py:746).  The reference builds its masks with torch.unique over
    arbitrary id values (loss_handler.py:938-967) and then requires n_strokes <= n_pred_masks for the assignment;
    here the ids index a one-hot of width n_pred_masks, so a larger id would silently drop its segments.  Called on
    the HOST copy of the batch (Trainer.to_device / pad_batch): no device synchronisation."""
    ids = stroke_ids.detach()
    if ids.numel() == 0:
        return
    lo, hi = float(ids.min()), float(ids.max())
    if hi >= n_pred_masks or lo < -1 or not bool((ids == ids.round()).all()):
        raise ValueError("stroke ids must be integers in [-1, %d) (-1 = padding); got range [%g, %g]" % (n_pred_masks, lo, hi))


def stroke_masks_loss(pred_to_gt_match, pred_stroke_masks, scores, stroke_ids, cfg, matcher="device", check_ids=False):
    """loss_handler.py:816-935 with smooth_targets=False (binary masks, BCE).

    matcher="device" (default): batched on-device assignment and fixed-shape masked reductions -- no
    host round trip, CUDA-graph capturable.  matcher="host": scipy on the host from one D2H copy (the
    reference's own solver; used by the parity tests).
    check_ids=True re-creates the reference's sanity asserts (:852-854: no predicted segment matched to the padding
    id, ids inside the mask budget) with one device->host synchronisation; never set inside a captured step -- the
    Trainer validates the host batch instead (validate_stroke_ids)."""
    dev = pred_stroke_masks.device
    B, n_pred_masks, out_segments = pred_stroke_masks.shape
    ids = stroke_ids.to(dev).gather(1, pred_to_gt_match)                         # :838  [B, out_segments], float
    ids = ids.long()                                                             # the -1 padding id is never matched (:852)
    n_ids = n_pred_masks                                                         # ids < max_n_strokes == n_pred_masks
    if check_ids and not torch.cuda.is_current_stream_capturing():
        assert not bool((ids == -1).any()), "no predicted segment should be associated with the fake stroke id -1"   # :852
        assert int(ids.max()) < n_ids, "stroke id %d outside the mask budget %d" % (int(ids.max()), n_ids)
    if matcher == "device":
        with torch.no_grad():
            # ids are NOT clamped: a (never expected) -1 matches no class and leaves an all-zero one-hot row
            cost, present, onehot = mask_cost_matrices(pred_stroke_masks, ids, n_ids)
            row = hungarian_device(cost, present)                                # [B, T]  (:860-877)
            pres_f = present.to(pred_stroke_masks.dtype)
            row_c = row.clamp_min(0)
        # matched pairs as dense [B, T] slots, absent strokes masked out (the reference stacks them, :886-902)
        sel = pred_stroke_masks.gather(1, row_c[:, :, None].expand(B, n_ids, out_segments))          # pred mask matched to stroke t
        bce = F.binary_cross_entropy_with_logits(sel, onehot.transpose(1, 2), reduction="none").sum(-1)   # [B, T]
        n_pairs = pres_f.sum()
        if cfg.mask_loss_global_mean and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            n_glob = n_pairs.detach().clone()
            torch.distributed.all_reduce(n_glob)
            n_pairs = n_glob / torch.distributed.get_world_size()                # rank-mean of (sum_r / this) = global sum / global count
        mask_loss = (bce * pres_f).sum() / n_pairs                               # :906  mean over all matched pairs in the batch
        target_scores = torch.zeros_like(scores).scatter_add_(1, row_c, pres_f)  # :920-921 (each mask matched at most once)
        weights = cfg.explicit_no_stroke_weight + (1.0 - cfg.explicit_no_stroke_weight) * target_scores   # :924-925
        conf_loss = F.binary_cross_entropy_with_logits(scores, target_scores, reduction="none", weight=weights).mean()   # :930
        return cfg.explicit_weight_stroke_masks * mask_loss + cfg.explicit_weight_stroke_masks_confidence * conf_loss
    with torch.no_grad():
        cost, present, onehot = mask_cost_matrices(pred_stroke_masks, ids, n_ids)
        b_idx, p_idx, t_idx = hungarian_host(cost, present)                      # :860-877
        b_idx, p_idx, t_idx = b_idx.to(dev), p_idx.to(dev), t_idx.to(dev)
    matched_pred = pred_stroke_masks[b_idx, p_idx]                               # :886  [M, out_segments]
    matched_tgt = onehot.transpose(1, 2)[b_idx, t_idx]                           # :902  [M, out_segments]
    mask_loss = F.binary_cross_entropy_with_logits(matched_pred, matched_tgt, reduction="none").sum(-1).mean()   # :906
    target_scores = torch.zeros_like(scores)                                     # :920-921
    target_scores[b_idx, p_idx] = 1.0
    weights = torch.full_like(scores, cfg.explicit_no_stroke_weight)             # :924-925
    weights[b_idx, p_idx] = 1.0
    conf_loss = F.binary_cross_entropy_with_logits(scores, target_scores, reduction="none", weight=weights).mean()   # :930
    return cfg.explicit_weight_stroke_masks * mask_loss + cfg.explicit_weight_stroke_masks_confidence * conf_loss


SCHEDULABLE = ("weight_asymm_segment_chamfer", "weight_reverse_asymm_point_chamfer", "weight_reverse_asymm_segment_chamfer",
               "explicit_weight_stroke_masks", "explicit_weight_stroke_masks_confidence")


class DeviceLossWeights:
    """The five schedulable loss weights (delayMasksLoss, PSACDScheduler: train_maskplanner.py:186-199 rewrite them
    between epochs) as ONE device tensor read by the step's kernels, so a step captured in a CUDA graph follows the
    schedule: `sync(cfg)` pushes changed host values (outside capture); inside the loss they are 0-d tensor views."""

    def __init__(self, device):
        self.t = torch.zeros(len(SCHEDULABLE), dtype=torch.float32, device=device)
        self._host = None

    def sync(self, cfg):
        vals = tuple(float(getattr(cfg, k)) for k in SCHEDULABLE)
        if vals != self._host:
            self.t.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
            self._host = vals

    def __getitem__(self, name):
        return self.t[SCHEDULABLE.index(name)]


class _FusedAsymmV6(torch.autograd.Function):
    """The whole loss as one autograd node over the fused kernels of csrc/loss.cu (six launches forward, three
    backward, no torch glue): lengths -> [segment NN -> mask costs -> Hungarian] next to [pose NN on a side stream] ->
    loss value.  Returns (loss 0-d, terms [8]); gradients flow to y_pred, the mask logits and the mask scores."""

    @staticmethod
    def forward(ctx, y_pred, masks, scores, y, stroke_ids, traj_as_pc, w5, no_stroke_w, pose_dim, global_mean):
        from . import _cabi
        from ._cabi import check, ptr, stream_ptr
        lib = _cabi.load()
        dev = y_pred.device
        x = y_pred.detach().float().contiguous()
        yy = y.detach().float().contiguous()
        pc = traj_as_pc.detach().float().contiguous()
        mk = masks.detach().float().contiguous()
        sc = scores.detach().float().contiguous()
        sid = stroke_ids.detach().float().contiguous()
        B, P1, D = x.shape
        P2, P3, NM = yy.shape[1], pc.shape[1], mk.shape[1]
        assert D % pose_dim == 0 and pc.shape[2] == pose_dim and mk.shape[2] == P1 and sid.shape[1] == P2
        i64 = dict(dtype=torch.int64, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        len_y, len_y2 = torch.empty(B, **i64), torch.empty(B, **i64)
        check(lib.mpb_loss_lengths_f32(ptr(yy), P2, D, ptr(pc), P3, pose_dim, B, CH.PAD_SENTINEL, ptr(len_y), ptr(len_y2), stream_ptr()),
              "mpb_loss_lengths_f32")
        d_x, idx_x = torch.empty(B, P1, **f32), torch.empty(B, P1, **i64)
        d_y, idx_y = torch.empty(B, P2, **f32), torch.empty(B, P2, **i64)
        d_y2, idx_y2 = torch.empty(B, P3, **f32), torch.empty(B, P3, **i64)
        # term 2 (ground-truth poses -> predicted poses) feeds nothing but the loss value: side stream, next to the
        # segment search -> cost matrices -> Hungarian chain
        with Fork(x, pc, len_y2, d_y2, idx_y2) as fork:
            check(lib.mpb_chamfer_nn_f32(ptr(x), ptr(pc), B, P1 * (D // pose_dim), P3, pose_dim, None, ptr(len_y2), None, None,
                                         ptr(d_y2), ptr(idx_y2), stream_ptr()), "mpb_chamfer_nn_f32")
        check(lib.mpb_chamfer_nn_f32(ptr(x), ptr(yy), B, P1, P2, D, None, ptr(len_y), ptr(d_x), ptr(idx_x), ptr(d_y), ptr(idx_y),
                                     stream_ptr()), "mpb_chamfer_nn_f32")
        cost = torch.empty(B, NM, NM, **f32)
        present = torch.empty(B, NM, dtype=torch.uint8, device=dev)
        ids = torch.empty(B, P1, dtype=torch.int32, device=dev)
        check(lib.mpb_mask_cost_f32(ptr(mk), ptr(sid), ptr(idx_x), B, NM, P1, P2, ptr(cost), ptr(present), ptr(ids), stream_ptr()),
              "mpb_mask_cost_f32")
        row = torch.empty(B, NM, **i64)
        check(lib.mpb_lap_f32(ptr(cost), ptr(present), B, NM, NM, ptr(row), stream_ptr()), "mpb_lap_f32")
        n_pairs = None
        if global_mean and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            n_pairs = present.sum().float().reshape(1)
            torch.distributed.all_reduce(n_pairs)
            n_pairs = n_pairs / torch.distributed.get_world_size()       # rank-mean of (sum_r / this) = global sum / global count
        fork.join(d_y2, idx_y2)
        loss = torch.empty((), **f32)
        terms = torch.empty(8, **f32)
        # the VALUE of the loss feeds nothing downstream (the backward kernels recompute what they need): its single-CTA
        # reduction runs on the side stream, off the step's critical path; whoever reads the loss joins that stream first
        with Fork(d_x, d_y, d_y2, len_y, len_y2, cost, present, row, sc, w5, n_pairs, loss, terms, slot=3) as vfork:
            check(lib.mpb_asymm_v6_loss_value_f32(ptr(d_x), ptr(d_y), ptr(len_y), ptr(d_y2), ptr(len_y2), ptr(cost), ptr(present),
                                                  ptr(row), ptr(sc), ptr(w5), ptr(n_pairs), float(no_stroke_w), B, P1, P2, P3, NM,
                                                  ptr(loss), ptr(terms), stream_ptr()), "mpb_asymm_v6_loss_value_f32")
        _PENDING_VALUE.append(vfork)
        ctx.n_pairs = n_pairs
        ctx.save_for_backward(x, yy, pc, mk, sc, idx_x, idx_y, len_y, idx_y2, len_y2, ids, present, row, w5)
        ctx.dims = (B, P1, P2, D, P3, pose_dim, NM, float(no_stroke_w))
        ctx.shapes = (y_pred.shape, masks.shape, scores.shape)
        ctx.mark_non_differentiable(terms)
        ctx.set_materialize_grads(False)       # no zero-filled gradient tensor for `terms` in front of the backward kernels
        ctx.match = idx_x
        return loss, terms

    @staticmethod
    def backward(ctx, g_loss, _g_terms):
        from . import _cabi
        from ._cabi import check, ptr, stream_ptr
        while BACKWARD_START_HOOKS:
            BACKWARD_START_HOOKS.pop()()
        if g_loss is None:
            return (None,) * 10
        x, yy, pc, mk, sc, idx_x, idx_y, len_y, idx_y2, len_y2, ids, present, row, w5 = ctx.saved_tensors
        B, P1, P2, D, P3, D2, NM, nsw = ctx.dims
        g = g_loss.detach().float().reshape(1).contiguous()
        gp, gm, gs = torch.empty_like(x), torch.empty_like(mk), torch.empty_like(sc)
        check(_cabi.load().mpb_asymm_v6_loss_bwd_f32(ptr(x), ptr(yy), ptr(pc), ptr(mk), ptr(sc), ptr(idx_x), ptr(idx_y), ptr(len_y),
                                                     ptr(idx_y2), ptr(len_y2), ptr(ids), ptr(present), ptr(row), ptr(w5), ptr(ctx.n_pairs), ptr(g),
                                                     nsw, B, P1, P2, D, P3, D2, NM, ptr(gp), ptr(gm), ptr(gs), stream_ptr()),
              "mpb_asymm_v6_loss_bwd_f32", launches=3)
        s0, s1, s2 = ctx.shapes
        return gp.view(s0), gm.view(s1), gs.view(s2), None, None, None, None, None, None, None


_PENDING_VALUE = []      # side-stream forks of loss-value kernels not yet joined
BACKWARD_START_HOOKS = []    # callables run once at the start of the fused loss's backward (the first node of the step's backward)


def join_loss_value():
    """Make the current stream wait for every loss-value kernel issued so far (call before reading a loss returned by
    the fused path on another stream's schedule; asymm_v6_chamfer_with_stroke_masks(join_value=True) does it itself)."""
    while _PENDING_VALUE:
        _PENDING_VALUE.pop().join()


_W5_CACHE = {}


def _weights_tensor(cfg, weights, device):
    """The five schedulable weights as a device array: the Trainer's DeviceLossWeights tensor, or a cached upload of the
    host values of `cfg` (eager callers)."""
    if weights is not None:
        return weights.t
    vals = tuple(float(getattr(cfg, k)) for k in SCHEDULABLE)
    key = (vals, str(device))
    t = _W5_CACHE.get(key)
    if t is None:
        if len(_W5_CACHE) > 64:
            _W5_CACHE.clear()
        t = _W5_CACHE[key] = torch.tensor(vals, dtype=torch.float32, device=device)
    return t


def asymm_v6_chamfer_with_stroke_masks(y_pred, y, pred_stroke_masks, mask_scores, stroke_ids, traj_as_pc, cfg=None,
                                       fused=True, return_terms=False, matcher="device", weights=None, join_value=True):
    """The whole training loss (loss_handler.py:596-666); per_segment_confidence is False in the MaskPlanner config.
    weights: optional DeviceLossWeights overriding cfg's five schedulable weights (CUDA-graph replays).
    fused=True (default): the fused kernels of csrc/loss.cu (_FusedAsymmV6) when the device matcher is used;
    fused="nn": terms 1 + 3 share one nearest-neighbour launch, everything around it stock torch ops (the round-1 path,
    kept for A/B runs and as a second implementation the parity tests compare against); fused=False: the reference's
    three chamfer_distance calls literally."""
    cfg = cfg or LossConfig()
    if fused is True and matcher == "device" and y_pred.is_cuda and pred_stroke_masks.shape[1] <= 32 \
            and os.environ.get("MPB_FUSED_LOSS", "1") == "1":
        # the training path: one autograd node over the fused kernels (csrc/loss.cu)
        loss, terms = _FusedAsymmV6.apply(y_pred, pred_stroke_masks, mask_scores, y, stroke_ids.to(y_pred.device), traj_as_pc,
                                          _weights_tensor(cfg, weights, y_pred.device), cfg.explicit_no_stroke_weight, cfg.pose_dim,
                                          bool(cfg.mask_loss_global_mean))
        if join_value:        # join_value=False: the caller (Trainer) joins after backward / the optimizer step instead
            join_loss_value()
        if return_terms:
            return loss, dict(asymm_segment=terms[0], reverse_point=terms[1], reverse_segment=terms[2], masks=terms[3])
        return loss
    if weights is not None:
        import copy
        cfg = copy.copy(cfg)
        for k in SCHEDULABLE:
            setattr(cfg, k, weights[k])
    t1, t3, _, match = chamfer_terms_13(y_pred, y, cfg, fused=bool(fused))
    # term 2 (a second nearest-neighbour search) does not feed the mask loss (cost matrices -> Hungarian solver ->
    # matched BCE/dice): the two run side by side (maskplanner_b200/streams.py)
    with Fork(y_pred, traj_as_pc) as fork:
        t2 = chamfer_term2(y_pred, traj_as_pc, cfg)
    masks = stroke_masks_loss(match, pred_stroke_masks, mask_scores, stroke_ids, cfg, matcher=matcher)
    fork.join(t2)
    loss = (cfg.weight_asymm_segment_chamfer * t1 + cfg.weight_reverse_asymm_point_chamfer * t2
            + cfg.weight_reverse_asymm_segment_chamfer * t3 + masks)
    if return_terms:
        return loss, dict(asymm_segment=t1.detach(), reverse_point=t2.detach(), reverse_segment=t3.detach(), masks=masks.detach())
    return loss
