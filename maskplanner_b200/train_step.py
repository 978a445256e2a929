"""The MaskPlanner training-step body on B200, batch-sharded data parallel (one process per GPU).

Mirrors the step of the reference loop (train_maskplanner.py:183-227): zero_grad -> permute + H2D of
the batch (:207-208) -> model forward (:210) -> loss (:212-218) -> backward (:220) -> Adam step (:221)
-> loss.item() (:223).  The reference is single-process; the data-parallel plumbing added here is the
only collective of the path: every rank owns whole samples (a cloud or a chamfer problem
never spans GPUs) and the flat fp32 gradient is all-reduced (averaged) once per step with NCCL over
NVLink.  Parameter gradients are views into ONE flat buffer, so the all-reduce needs no packing
copies, and the head gradients (>90 % of the parameters, produced first by backward) are reduced on a
side stream while the encoder backward is still running.

No step of the body synchronises with the host (the reference's padding scan, per-sample Hungarian
and .cpu() calls are all on the device here), so with ``use_graph=True`` the whole step -- forward,
loss, backward, all-reduce, Adam -- is captured once into a CUDA graph and replayed; per call the host
issues one staging launch (``mpb_stage_batch``: the batch into the graph's fixed buffers, GT rows padded
with the loader's sentinels), one seed copy and the replay.  The FPS / ball-query plan of the NEXT batch
(``step(..., next_batch=...)``) is computed on a side stream inside the current step (sampling depends on
the cloud alone), so no step starts with a 0.2 ms FPS that nothing can overlap; ``step_from_host_async``
keeps the pinned H2D copies two calls ahead and hands the loss back as a ``PendingLoss``.

BatchNorm semantics: replica semantics (each rank normalises over its own samples), the standard DDP
behaviour; single-process-equivalent statistics would need SyncBN (not built).  `Trainer.sync_bn_buffers()` averages
the running statistics over the ranks before a checkpoint; `LossConfig.mask_loss_global_mean` makes the stroke-mask
loss's mean over matched pairs (loss_handler.py:906) a mean over the GLOBAL batch instead of a mean of per-rank means.
"""
import os

import torch
import torch.distributed as dist

from . import gradsink
from . import loss as L
from . import optim, regressor, streams, synthetic
from .pointnet2_utils import draw_fps_seed


class FlatGradBuckets:
    """All parameter .grad tensors as views into one flat fp32 buffer, split into two buckets:
    `heads` (everything outside sa1/sa2/sa3) and `encoder`."""

    def __init__(self, model):
        named = list(model.named_parameters())
        heads = [p for n, p in named if not n.startswith(("sa1.", "sa2.", "sa3."))]
        enc = [p for n, p in named if n.startswith(("sa1.", "sa2.", "sa3."))]
        self.params = heads + enc
        pad4 = lambda n: (n + 3) // 4 * 4            # every view starts 16-byte aligned (vector path of the Adam kernel)
        n_heads = sum(pad4(p.numel()) for p in heads)
        total = n_heads + sum(pad4(p.numel()) for p in enc)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.heads = self.flat[:n_heads]
        self.encoder = self.flat[n_heads:]
        off = 0
        self.views = []
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            self.views.append(p.grad)
            off += pad4(p.numel())
        self.head_params = heads

    def zero(self):
        self.flat.zero_()

    def register_sinks(self):
        """Let the library's backward kernels write each gradient straight into its view (maskplanner_b200.gradsink):
        no per-step memset, no AccumulateGrad read-modify-write passes."""
        gradsink.clear()
        for p, v in zip(self.params, self.views):
            gradsink.register(p, v)


def shard_range(global_batch, rank, world_size):
    """Samples [lo, hi) owned by `rank`: contiguous, whole samples, sizes differ by at most one."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_mean_(flat, world_size, group=None):
    """In-place mean over ranks of a flat gradient bucket (SUM then scale: works on NCCL and gloo)."""
    if world_size > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world_size)
    return flat


def all_reduce_sum_(flat, world_size, group=None):
    """In-place SUM over ranks; the 1/world_size of the average is folded into the optimizer kernel (grad_scale)."""
    if world_size > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def pad_batch(batch, max_segments, max_poses):
    """Pad the GT tensors to fixed maxima with the loader's own sentinels (-100 rows, stroke id -1;
    utils/dataset/paintnet_ODv1.py:738-747).  Extra sentinel rows change nothing: every consumer derives
    the per-sample length from the first sentinel row.  Fixed shapes are what CUDA-graph replay needs."""
    def pad(t, n, value):
        if t.shape[1] == n:
            return t
        assert t.shape[1] < n, "batch exceeds the configured maximum (%d > %d)" % (t.shape[1], n)
        out = t.new_full((t.shape[0], n) + tuple(t.shape[2:]), value)
        out[:, :t.shape[1]] = t
        return out
    out = dict(batch)
    out["traj"] = pad(batch["traj"], max_segments, synthetic.PAD)
    out["stroke_ids"] = pad(batch["stroke_ids"], max_segments, -1.0)
    out["traj_as_pc"] = pad(batch["traj_as_pc"], max_poses, synthetic.PAD)
    return out


class Trainer:
    """model + Adam(lr=1e-3) (train_maskplanner.py:159) + loss, optionally data parallel / CUDA-graphed."""

    KEYS = ("point_cloud", "traj", "traj_as_pc", "stroke_ids")

    def __init__(self, category="windows_v2", device=None, lr=1e-3, seed=0, loss_cfg=None, world_size=1, fused_loss=True,
                 use_graph=False, graph_warmup_steps=2, heads_tf32=None, pipeline_sampling=None):
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        torch.manual_seed(seed)                       # identical initial weights on every rank
        self.model = regressor.maskplanner_model(category).to(self.device)
        self.model.train()
        self.loss_cfg = loss_cfg or L.LossConfig()
        self.world_size = world_size
        self.fused_loss = fused_loss
        # Heads (nn.Linear with M = batch size) stay library GEMMs.  With the bf16 tensor-core
        # encoder they run on TF32 tensor cores (10-bit mantissa, finer than the encoder's bf16 activations);
        # with the strict-fp32 encoder they stay strict fp32 so that mode keeps the reference's arithmetic.
        from .pointnet2_utils import get_mlp_precision
        self.heads_tf32 = get_mlp_precision() in ("bf16", "tf32") if heads_tf32 is None else bool(heads_tf32)
        # The heads' M = batch GEMMs are library calls either way; cuBLASLt's heuristics are preferred for them (measured
        # 3.95 vs 3.97 ms per step; a single run showed 3.87 and was not reproduced).  MPB_BLAS=cublas|cublaslt|default
        # overrides.  This is a process-wide torch setting (torch.backends.cuda.preferred_blas_library), applied once.
        self.heads_blas = os.environ.get("MPB_BLAS", "cublaslt")
        if self.heads_blas != "default" and self.device.type == "cuda":
            torch.backends.cuda.preferred_blas_library(self.heads_blas)
        cfg = synthetic.CATEGORIES[category]
        self.max_segments = synthetic.out_vectors(cfg["n_pred_traj_points"])   # GT segments never exceed the prediction budget
        self.max_poses = cfg["n_pred_traj_points"]
        # Data parallel: gradients live in one flat buffer (all-reduce without packing).  Single GPU: no collective, so
        # gradients are left unset between steps and autograd hands its freshly computed tensors over as .grad -- this
        # saves one read-modify-write accumulation kernel per parameter (~80 launches) and the flat-buffer memset.
        self.buckets = FlatGradBuckets(self.model) if world_size > 1 else None
        # Direct gradient placement (DP): every gradient is written by its producing kernel into the flat buffer, the
        # collective sums it and the 1/W of the average rides in the Adam kernel.  Needs the fused heads (all parameters
        # then come from this library's autograd nodes) and this library's Adam; MPB_DIRECT_GRADS=0 restores the round-1
        # plumbing (memset + AccumulateGrad + scaling passes) for A/B runs.
        self.direct_grads = (world_size > 1 and self.model.fused_heads and os.environ.get("MPB_DIRECT_GRADS", "1") == "1"
                             and os.environ.get("MPB_TORCH_ADAM", "0") != "1")
        if self.direct_grads:
            self.buckets.register_sinks()
        else:
            gradsink.clear()
        # one-launch Adam (csrc/adam.cu); MPB_TORCH_ADAM=1 swaps torch's fused implementation back in for A/B runs
        if os.environ.get("MPB_TORCH_ADAM", "0") == "1":
            self.opt = torch.optim.Adam(self.model.parameters(), lr=lr, fused=True, capturable=use_graph)
        else:
            self.opt = optim.Adam(self.model.parameters(), lr=lr)
        # schedulable loss weights live in device memory when the step is replayed from a graph
        self.loss_weights = L.DeviceLossWeights(self.device) if use_graph else None
        self.n_masks = cfg["max_n_strokes"]
        self.comm_stream = torch.cuda.Stream(device=self.device) if world_size > 1 else None
        self._heads_ready = None
        self._heads_pending = 0
        self._copy_stream = None
        self._staged = None
        self._step_stream = None
        self.use_graph = use_graph
        # Sampling pipelined across steps: the FPS / ball-query indices of sa1 and sa2 depend on the cloud alone, so the plan
        # of the NEXT batch (step(..., next_batch=...)) is computed on a side stream while this batch's heads, loss and head
        # backward run (small-grid kernels that leave most SMs idle), instead of heading the next step's critical path with a
        # 64-CTA, 0.2 ms FPS that nothing can overlap.  The same work per step, every step; MPB_PIPELINE_SAMPLING=0 disables.
        if pipeline_sampling is None:
            pipeline_sampling = os.environ.get("MPB_PIPELINE_SAMPLING", "1") == "1"
        self.pipeline_sampling = bool(pipeline_sampling)
        self._plan_cur = self._plan_next = None      # preallocated plans (fixed addresses: graph replay)
        self._plan_for = None                        # the cloud tensor `_plan_cur` was computed for
        self._next_static = None
        self._graph = None
        from .shared_mlp import PackAhead
        self._packs = PackAhead()
        self._unit_grad = None
        self._static = None
        self._static_loss = None
        self._calls = 0
        self._graph_warmup_steps = graph_warmup_steps
        if self.direct_grads:
            def heads_done():
                # called at the end of the heads' backward node: every head gradient has been ENQUEUED (on the step's
                # stream or on the side stream the stroke-mask head runs on); the collective waits for both
                producers = {torch.cuda.current_stream(), self._step_stream}
                if streams.enabled():
                    producers.add(streams.side_stream(self.device))
                for s in producers:
                    if s is not None:
                        self.comm_stream.wait_stream(s)
                with torch.cuda.stream(self.comm_stream):
                    all_reduce_sum_(self.buckets.heads, self.world_size)
                self._heads_ready = torch.cuda.Event()
                self._heads_ready.record(self.comm_stream)
            gradsink.HOOKS["heads_done"] = heads_done
        elif world_size > 1:
            # launch the head-bucket all-reduce on the side stream as soon as backward has produced the
            # LAST head gradient (counted, so no assumption about autograd's execution order)
            n_heads = len(self.buckets.head_params)

            def hook(_p):
                self._heads_pending += 1
                if self._heads_pending == n_heads:
                    # every head gradient has been ENQUEUED by now, on the step's stream or on the side stream the
                    # stroke-mask head runs on (maskplanner_b200/streams.py): the collective waits for all of them
                    producers = {torch.cuda.current_stream(), self._step_stream}
                    if streams.enabled():
                        producers.add(streams.side_stream(self.device))
                    for s in producers:
                        self.comm_stream.wait_stream(s)
                    with torch.cuda.stream(self.comm_stream):
                        all_reduce_mean_(self.buckets.heads, self.world_size)
                    self._heads_ready = torch.cuda.Event()
                    self._heads_ready.record(self.comm_stream)
            for p in self.buckets.head_params:
                p.register_post_accumulate_grad_hook(hook)

    def sync_bn_buffers(self):
        """Average the BatchNorm running statistics over the ranks (replica semantics lets them drift apart; call before
        saving a checkpoint so the result does not depend on which rank writes it)."""
        if self.world_size <= 1:
            return
        for name, buf in self.model.named_buffers():
            if name.endswith(("running_mean", "running_var")):
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                buf.div_(self.world_size)

    def to_device(self, host_batch):
        """The step's H2D boundary (train_maskplanner.py:207-208, loss_handler.py:628-629): pinned -> device, async."""
        if not host_batch["stroke_ids"].is_cuda:
            L.validate_stroke_ids(host_batch["stroke_ids"], self.n_masks)       # host-side, no synchronisation
        return {k: host_batch[k].to(self.device, dtype=torch.float32, non_blocking=True) for k in self.KEYS}

    def _step_core(self, batch, fps_seeds, plan=None, nxt=None):
        old_tf32 = torch.backends.cuda.matmul.allow_tf32
        if self.heads_tf32:
            torch.backends.cuda.matmul.allow_tf32 = True  # head GEMMs (M = batch, library calls) on TF32 tensor cores
        try:
            self._packs.begin()           # weight operands of all shared-MLP layers on a side branch (shared_mlp.PackAhead)
            return self._step_body(batch, fps_seeds, plan, nxt)
        finally:
            self._packs.end()
            torch.backends.cuda.matmul.allow_tf32 = old_tf32

    def _step_body(self, batch, fps_seeds, plan=None, nxt=None):
        """plan: sampling plan of this batch (None: computed in line).  nxt = (next cloud [B,N,3], next seeds, plan buffers):
        the next batch's plan is computed on a side stream between the encoder forward and the end of the step."""
        self._step_stream = torch.cuda.current_stream()
        if self.buckets is not None:                                              # model.zero_grad()  (:184)
            if not self.direct_grads:
                self.buckets.zero()       # direct placement overwrites every gradient each step: nothing to clear
        else:
            for p in self.model.parameters():
                p.grad = None
        cloud = batch["point_cloud"].permute(0, 2, 1)                             # :207
        fork = None

        def fork_sampling():
            # the NEXT batch's sampling plan (FPS -> centroids -> ball query of sa1 and sa2) on a side stream
            nonlocal fork
            n_cloud, n_seeds, n_out = nxt
            with streams.Fork(n_cloud, slot=2) as fork:
                with torch.no_grad():
                    self.model.sampling_plan(n_cloud.permute(0, 2, 1), n_seeds, out=n_out)

        # Where the branch starts (MPB_SAMPLING_FORK): "encode" (default) = after the encoder forward, next to heads + loss +
        # head backward; "loss_bwd" = at the start of the loss backward, next to the head backward and SA3's backward.  The
        # first placement runs the 64-CTA FPS beside the two fma-bound nearest-neighbour searches of the loss, which then take
        # 0.20-0.23 ms instead of 0.08-0.10 (profiles/r02_graph_timeline.txt); the second keeps them fast but its tail reaches the
        # persistent GEMMs of SA2's backward, which cannot share an SM with an FPS CTA: measured 3.17 vs 3.14 ms, so "encode" stays.
        where = os.environ.get("MPB_SAMPLING_FORK", "encode") if nxt is not None else None
        del L.BACKWARD_START_HOOKS[:]
        if where == "loss_bwd" and self.fused_loss and os.environ.get("MPB_FUSED_LOSS", "1") == "1":
            L.BACKWARD_START_HOOKS.append(fork_sampling)
        pred, masks, scores, _ = self.model(cloud, fps_seeds, plan=plan,
                                            after_encode=fork_sampling if where in ("encode", "loss_bwd") and not L.BACKWARD_START_HOOKS
                                            and nxt is not None else None)   # :210
        loss = L.asymm_v6_chamfer_with_stroke_masks(pred, batch["traj"], masks, scores, batch["stroke_ids"],
                                                    batch["traj_as_pc"], self.loss_cfg, fused=self.fused_loss,
                                                    weights=self.loss_weights, join_value=False)    # :212-218
        self._heads_pending = 0
        if self._unit_grad is None:
            self._unit_grad = torch.ones((), dtype=torch.float32, device=self.device)
        # :220 (the root gradient handed in: autograd would otherwise launch a fill for its ones_like(loss))
        loss.backward(self._unit_grad if loss.dtype == torch.float32 and loss.dim() == 0 else None)
        if self.world_size > 1:
            reduce_ = all_reduce_sum_ if self.direct_grads else all_reduce_mean_
            reduce_(self.buckets.encoder, self.world_size)
            if self._heads_ready is not None:
                torch.cuda.current_stream().wait_event(self._heads_ready)
                self._heads_ready = None
            else:  # hooks did not all fire (a head without gradient): reduce the bucket here
                reduce_(self.buckets.heads, self.world_size)
        handover = None
        if fork is not None:
            fork.join()
            if plan is not None:
                # the plan just computed becomes the current one: backward has finished reading `plan` and the optimizer does
                # not touch it, so the six copies run beside Adam instead of after it
                flat_dst = [t for t3 in plan for t in t3]
                flat_src = [t for t3 in nxt[2] for t in t3]
                with streams.Fork(*flat_src, *flat_dst, slot=2) as handover:
                    for dst, src in zip(flat_dst, flat_src):
                        dst.copy_(src)
        if self.direct_grads:
            self.opt.step(grad_scale=1.0 / self.world_size)                       # :221 (the buffer holds the SUM over ranks)
        else:
            self.opt.step()                                                       # :221
        L.join_loss_value()       # the loss VALUE was reduced on a side stream, off the critical path
        if handover is not None:
            handover.join()
        return loss.detach()

    def _ensure_plans(self, B):
        if self._plan_cur is None or self._plan_cur[0][0].shape[0] != B:
            from .pointnet2_utils import empty_sampling_plan
            specs = self.model.sampling_specs()
            self._plan_cur = empty_sampling_plan(B, specs, self.device)
            self._plan_next = empty_sampling_plan(B, specs, self.device)
            self._plan_for = None

    def _draw_seeds(self, B, n_points):
        """The reference's seed draws (models/pointnet2_utils.py:77: one CPU randint per SA layer, same generator consumption),
        staged through ONE pinned [2, B] tensor and one asynchronous copy."""
        host = torch.empty(2, B, dtype=torch.long, pin_memory=self.device.type == "cuda")
        host[0] = torch.randint(0, n_points, (B,), dtype=torch.long)
        host[1] = torch.randint(0, self.model.sa1.npoint, (B,), dtype=torch.long)
        dev = host.to(self.device, non_blocking=True)
        return (dev[0], dev[1])

    _PAD_BITS = {"traj": 0xC2C80000, "traj_as_pc": 0xC2C80000, "stroke_ids": 0xBF800000, "point_cloud": 0}   # -100.0f, -1.0f

    def _stage_inputs(self, batch, seeds, next_cloud, next_seeds):
        """Copy this call's inputs into the captured graph's fixed buffers with ONE launch (mpb_stage_batch): the four batch
        tensors, padded to the configured maxima with the loader's sentinels, plus (pipelined sampling) the next batch's cloud
        and seeds.  Returns False when a tensor does not qualify (then the caller falls back to pad_batch + copy_)."""
        import ctypes
        from . import _cabi
        segs = []
        for k in self.KEYS:
            src, dst = batch[k], self._static[k]
            if not (src.is_cuda and src.dtype == torch.float32 and src.is_contiguous() and src.shape[0] == dst.shape[0]
                    and src.shape[1] <= dst.shape[1] and src.shape[2:] == dst.shape[2:]):
                return False
            row = 1
            for d in dst.shape[2:]:
                row *= d
            segs.append((src, dst, dst.shape[0], src.shape[1], dst.shape[1], row, self._PAD_BITS[k]))
        if next_cloud is not None:
            dst = self._next_static[0]
            if not (next_cloud.is_cuda and next_cloud.dtype == torch.float32 and next_cloud.is_contiguous() and next_cloud.shape == dst.shape):
                return False
            segs.append((next_cloud, dst, 1, 1, 1, dst.numel(), 0))
        for src2, dst2 in ((seeds, self._static.get("seeds")), (next_seeds, self._next_static[1] if self._next_static else None)):
            if src2 is None:
                continue
            for a, b in zip(src2, dst2):
                if not (a.is_cuda and a.dtype == torch.long and a.is_contiguous() and a.shape == b.shape):
                    return False
                segs.append((a, b, 1, 1, 1, 2 * b.numel(), 0))
        n = len(segs)
        if n > 8:
            return False
        vp, i64, u32 = ctypes.c_void_p * n, ctypes.c_int64 * n, ctypes.c_uint32 * n
        _cabi.check(_cabi.load().mpb_stage_batch(
            n, vp(*[g[0].data_ptr() for g in segs]), vp(*[g[1].data_ptr() for g in segs]), i64(*[g[2] for g in segs]),
            i64(*[g[3] for g in segs]), i64(*[g[4] for g in segs]), i64(*[g[5] for g in segs]), u32(*[g[6] for g in segs]),
            _cabi.stream_ptr()), "mpb_stage_batch")
        return True

    def step(self, batch, fps_seeds=None, next_batch=None, next_fps_seeds=None):
        """One optimisation step on a device-resident batch.  Returns the loss as a 0-d device tensor.
        `fps_seeds` = (seed indices for SA1 [B], for SA2 [B]); None draws them from the CPU generator exactly
        like the reference does (models/pointnet2_utils.py:77, one randint per SA layer).
        `next_batch` (pipeline_sampling): the batch of the FOLLOWING call; its sampling plan (FPS seeds `next_fps_seeds`, or
        drawn the same way) is computed during this step.  A batch whose plan was not announced that way, or explicit
        `fps_seeds`, gets its plan computed in line first -- same results either way."""
        pipe = self.pipeline_sampling
        if not self.use_graph and not pipe:
            return self._step_core(batch, fps_seeds)
        self._calls += 1
        B, n_points = batch["point_cloud"].shape[0], batch["point_cloud"].shape[1]
        to_dev = lambda seeds: tuple(s.to(self.device, dtype=torch.long) for s in seeds)
        plan = nxt = None
        if pipe:
            self._ensure_plans(B)
            if fps_seeds is not None or self._plan_for is not batch["point_cloud"]:
                seeds = to_dev(fps_seeds) if fps_seeds is not None else self._draw_seeds(B, n_points)
                with torch.no_grad():
                    self.model.sampling_plan(batch["point_cloud"].permute(0, 2, 1), seeds, out=self._plan_cur)
            plan = self._plan_cur
            self._plan_for = None
            if next_batch is not None:
                n_seeds = to_dev(next_fps_seeds) if next_fps_seeds is not None else self._draw_seeds(B, n_points)
                n_cloud = next_batch["point_cloud"]
                assert n_cloud.shape == batch["point_cloud"].shape, "pipelined sampling needs equal batch shapes"
        if not self.use_graph:
            if pipe and next_batch is not None:
                nxt = (n_cloud, n_seeds, self._plan_next)
            out = self._step_core(batch, fps_seeds, plan, nxt)
            if nxt is not None:
                self._plan_for = n_cloud
            return out
        # host-side hyper-parameters a scheduler may have changed since the last call (learning rate:
        # train_maskplanner.py:230; loss weights: :186-199) -> device scalars the captured kernels read
        if hasattr(self.opt, "sync_hyper"):
            self.opt.sync_hyper()
        self.loss_weights.sync(self.loss_cfg)
        if not pipe:
            if fps_seeds is None:
                fps_seeds = self._draw_seeds(B, n_points)
            else:
                fps_seeds = to_dev(fps_seeds)
        if self._graph is not None:
            # steady state: ONE staging launch (padding included) + the replay
            if not self._stage_inputs(batch, None if pipe else fps_seeds, n_cloud if (pipe and next_batch is not None) else None,
                                      n_seeds if (pipe and next_batch is not None) else None):
                padded = pad_batch(batch, self.max_segments, self.max_poses)
                for k in self.KEYS:
                    self._static[k].copy_(padded[k], non_blocking=True)
                if not pipe:
                    for dst, src in zip(self._static["seeds"], fps_seeds):
                        dst.copy_(src, non_blocking=True)
                elif next_batch is not None:
                    self._next_static[0].copy_(n_cloud, non_blocking=True)
                    for dst, src in zip(self._next_static[1], n_seeds):
                        dst.copy_(src, non_blocking=True)
            self._graph.replay()
            out = self._static_loss
        else:
            batch = pad_batch(batch, self.max_segments, self.max_poses)
            if pipe and self._next_static is None:
                self._next_static = (torch.zeros_like(batch["point_cloud"]), tuple(torch.zeros(B, dtype=torch.long, device=self.device) for _ in range(2)))
            if pipe and next_batch is not None:
                self._next_static[0].copy_(n_cloud, non_blocking=True)
                for dst, src in zip(self._next_static[1], n_seeds):
                    dst.copy_(src, non_blocking=True)
            if pipe:
                # the side branch is part of the captured graph: without a next batch it recomputes a plan nobody will use
                nxt = (self._next_static[0], self._next_static[1], self._plan_next)
            if self._calls <= self._graph_warmup_steps:      # real eager steps: lazy initialisation + Adam state
                out = self._step_core(batch, fps_seeds, plan, nxt)
            else:
                self._static = {k: batch[k].clone() for k in self.KEYS}
                if not pipe:
                    self._static["seeds"] = tuple(s.clone() for s in fps_seeds)
                torch.cuda.synchronize()
                from . import _cabi
                n0 = _cabi.KERNEL_LAUNCHES
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._static_loss = self._step_core(self._static, None if pipe else self._static["seeds"], plan, nxt)
                self.kernels_per_step = _cabi.KERNEL_LAUNCHES - n0     # libmaskplanner_b200 kernels inside one replay
                self._graph.replay()
                out = self._static_loss
        if pipe:
            self._plan_for = n_cloud if next_batch is not None else None
        return out

    def prefetch(self, host_batch):
        """Start the H2D copy of a FUTURE step's pinned batch on a side stream, so it overlaps the step in flight (what
        the reference's DataLoader workers + pin_memory + non_blocking copies do, train_maskplanner.py:207-208)."""
        if self._staged is None:
            self._staged = {}
        if id(host_batch) in self._staged:
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        with torch.cuda.stream(self._copy_stream):
            dev = self.to_device(host_batch)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        while len(self._staged) >= 4:                       # bounded look-ahead
            self._staged.pop(next(iter(self._staged)))
        self._staged[id(host_batch)] = (host_batch, dev, ev, [False])

    def _take_staged(self, host_batch, pop):
        """Device copy of a host batch: the prefetched one (made visible to the current stream) or a copy issued now."""
        entry = (self._staged or {}).get(id(host_batch))
        if entry is None or entry[0] is not host_batch:
            return self.to_device(host_batch)
        _, dev, ev, seen = entry
        if not seen[0]:
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            for t in dev.values():
                t.record_stream(cur)          # allocated on the copy stream, consumed here
            seen[0] = True
        if pop:
            self._staged.pop(id(host_batch))
        return dev

    def step_from_host_async(self, host_batch, fps_seeds=None, next_host_batch=None, after_next_host_batch=None, next_fps_seeds=None):
        """step_from_host without the host-side wait: returns a PendingLoss whose .result() blocks until this step's loss
        has reached pinned host memory.  A loop that reads step i's loss after enqueuing step i+1 keeps the GPU busy while
        the host prepares the next call (staging launch, seed draws, prefetch)."""
        dev = self._take_staged(host_batch, pop=True)
        nxt = None
        if self.pipeline_sampling and next_host_batch is not None:
            if (self._staged or {}).get(id(next_host_batch)) is None:
                self.prefetch(next_host_batch)           # first call / no look-ahead given: the copy heads this step
            nxt = self._take_staged(next_host_batch, pop=False)
        loss = self.step(dev, fps_seeds, next_batch=nxt, next_fps_seeds=next_fps_seeds)
        pending = PendingLoss(loss)
        for hb in (next_host_batch, after_next_host_batch):
            if hb is not None:
                self.prefetch(hb)
        return pending

    def step_from_host(self, host_batch, fps_seeds=None, next_host_batch=None, after_next_host_batch=None, next_fps_seeds=None):
        """End-to-end step as a user calls it: pinned host batch in, Python float loss out (:223).
        `next_host_batch`: the batch of the following call.  Its H2D copy is issued right after this step has been
        enqueued and runs concurrently with it; with pipeline_sampling its sampling plan is computed during this step, which
        needs its cloud on the device already -- pass `after_next_host_batch` (the batch after that) as well and every
        copy is issued two calls ahead, off the critical path."""
        return self.step_from_host_async(host_batch, fps_seeds, next_host_batch, after_next_host_batch, next_fps_seeds).result()


class PendingLoss:
    """A step's loss on its way to the host: asynchronous device -> pinned-host copy (4 bytes) + an event."""

    def __init__(self, loss):
        self._host = torch.empty((), dtype=torch.float32, pin_memory=True)
        self._host.copy_(loss.detach(), non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def result(self):
        self._event.synchronize()
        return float(self._host)


def pin_batch(batch):
    return {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
