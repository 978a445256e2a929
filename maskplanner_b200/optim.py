"""Adam on one kernel launch (a ``torch.optim.Optimizer`` for the MaskPlanner step).

Reference: train_maskplanner.py:159 (``torch.optim.Adam(model.parameters(), lr=config.lr)``), :160
(``get_lr_scheduler(opt, ...)``), :221 (``opt.step()``) and :230 (``scheduler.step()``).  Same update rule,
defaults and ``state_dict`` layout as ``torch.optim.Adam`` (no amsgrad, no maximize); it IS a
``torch.optim.Optimizer`` subclass, so ``torch.optim.lr_scheduler`` classes and checkpoint tooling accept it.
The arithmetic runs in ``mpb_adam_step_f32`` (csrc/adam.cu): every parameter of a group in a single launch
instead of torch's three multi-tensor launches at ~1.4 TB/s.

Differences from torch's class, all deliberate:
* The step count and the learning rate live in DEVICE scalars, so a step captured in a CUDA graph keeps
  advancing.  A scheduler writes ``param_groups[i]["lr"]`` on the host; ``sync_hyper()`` pushes changed values to
  the device (``step()`` calls it when not capturing, ``Trainer`` calls it before every graph replay).
* One step counter is shared by each launch group (<= 80 tensors of one param group) and advances on every
  ``step()`` in which at least one tensor of the group had a gradient -- torch counts per parameter.  The
  reference never skips parameters, so the counters agree there; ``state_dict()["state"][i]["step"]`` reports the
  group's counter.
* ``grad_scale`` (per call) multiplies every gradient on load: data-parallel runs fold the 1/world_size of the
  gradient average into the update instead of a separate pass over the flat gradient.
"""
import ctypes

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr

_MAX_TENSORS = 80     # kAdamMaxTensors in csrc/adam.cu


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        capturable=True, fused=None, foreach=None, differentiable=False)
        super().__init__(params, defaults)
        self._launch = []        # per param group: list of (tensors, scalars[step, lr], ticket)
        for g in self.param_groups:
            ps = g["params"]
            for p in ps:
                _cabi.require_cuda(p)
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise ValueError("maskplanner_b200.optim.Adam handles contiguous fp32 parameters")
            dev = ps[0].device
            groups = []
            for i in range(0, len(ps), _MAX_TENSORS):
                chunk = ps[i:i + _MAX_TENSORS]
                scal = torch.zeros(2, dtype=torch.float32, device=dev)      # [step count, learning rate]
                ticket = torch.zeros(1, dtype=torch.int32, device=dev)      # last-CTA election of the kernel
                for p in chunk:
                    self.state[p] = dict(step=scal[0:1].view(()), exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
                groups.append((chunk, scal, ticket))
            self._launch.append(groups)
        self._lr_on_device = [None] * len(self.param_groups)

    # -- hyper-parameters that live on the device -----------------------------------------------------------------
    def sync_hyper(self):
        """Push host-side ``param_groups[i]['lr']`` changes (LR schedulers) to the device scalars the kernel reads.
        A host -> device scalar write; call it OUTSIDE any captured region (it is a no-op when nothing changed)."""
        for gi, g in enumerate(self.param_groups):
            lr = float(g["lr"])
            if lr != self._lr_on_device[gi]:
                for _, scal, _ in self._launch[gi]:
                    scal[1] = lr
                self._lr_on_device[gi] = lr

    _sync_lr = sync_hyper      # round-1 name

    def set_lr(self, lr):
        """Change the learning rate of every group (also between CUDA-graph replays)."""
        for g in self.param_groups:
            g["lr"] = lr
        self.sync_hyper()

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        lib = _cabi.load()
        for g, groups in zip(self.param_groups, self._launch):
            for params, scal, ticket in groups:
                live = [p for p in params if p.grad is not None]
                if not live:
                    continue
                grads = []
                for p in live:
                    gr = p.grad
                    if gr.dtype != torch.float32 or not gr.is_contiguous():
                        gr = gr.float().contiguous()
                    grads.append(gr)
                n = len(live)
                arr = ctypes.c_void_p * n
                check(lib.mpb_adam_step_f32(n, arr(*[p.data_ptr() for p in live]), arr(*[t.data_ptr() for t in grads]),
                                            arr(*[self.state[p]["exp_avg"].data_ptr() for p in live]),
                                            arr(*[self.state[p]["exp_avg_sq"].data_ptr() for p in live]),
                                            (ctypes.c_int64 * n)(*[p.numel() for p in live]), float(g["lr"]),
                                            ctypes.c_void_p(scal.data_ptr() + 4), g["betas"][0], g["betas"][1], g["eps"],
                                            g["weight_decay"], float(grad_scale), ptr(scal), ptr(ticket), stream_ptr()),
                      "mpb_adam_step_f32")
        return loss

    def step_flat(self, param, grad, exp_avg, exp_avg_sq, group=0, slot=0, grad_scale=1.0):
        """One launch over explicit flat fp32 buffers (data-parallel sharded update: this rank's slice of the flat
        parameter / gradient buffers and its shard of the optimizer state), sharing the group's step count and lr."""
        g = self.param_groups[group]
        _, scal, ticket = self._launch[group][slot]
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        arr = ctypes.c_void_p * 1
        check(_cabi.load().mpb_adam_step_f32(1, arr(param.data_ptr()), arr(grad.data_ptr()), arr(exp_avg.data_ptr()),
                                             arr(exp_avg_sq.data_ptr()), (ctypes.c_int64 * 1)(param.numel()), float(g["lr"]),
                                             ctypes.c_void_p(scal.data_ptr() + 4), g["betas"][0], g["betas"][1], g["eps"],
                                             g["weight_decay"], float(grad_scale), ptr(scal), ptr(ticket), stream_ptr()),
              "mpb_adam_step_f32")

    # -- checkpoints ---------------------------------------------------------------------------------------------------
    def state_dict(self):
        sd = super().state_dict()
        for st in sd["state"].values():          # detach the shared counters: a checkpoint must not alias live state
            st["step"] = st["step"].clone()
        return sd

    def load_state_dict(self, state_dict):
        # keep the device-resident buffers (their addresses may be baked into a captured graph): copy values in place
        own = {id(p): self.state[p] for g in self.param_groups for p in g["params"]}
        super().load_state_dict(state_dict)
        for gi, (g, groups) in enumerate(zip(self.param_groups, self._launch)):
            for params, scal, _ in groups:
                steps = []
                for p in params:
                    new, old = self.state[p], own[id(p)]
                    if new is not old:
                        if "exp_avg" in new:
                            old["exp_avg"].copy_(new["exp_avg"])
                            old["exp_avg_sq"].copy_(new["exp_avg_sq"])
                            steps.append(float(new["step"]))
                        self.state[p] = old
                if steps:
                    scal[0] = max(steps)
        self._lr_on_device = [None] * len(self.param_groups)
        self.sync_hyper()
