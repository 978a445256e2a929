"""Adam on one kernel launch (host mirror of ``torch.optim.Adam`` for the MaskPlanner step).

Reference: train_maskplanner.py:159 (``torch.optim.Adam(model.parameters(), lr=config.lr)``) and :221
(``opt.step()``).  Same update rule, defaults, ``param_groups`` / ``state_dict`` layout and ``zero_grad``
behaviour as torch's class, restricted to what the reference uses (one parameter group, no amsgrad, no
maximize); the arithmetic runs in ``mpb_adam_step_f32`` (csrc/adam.cu): every parameter of the model in a
single launch instead of torch's three multi-tensor launches at ~1.4 TB/s.  The step count and the learning
rate live in device scalars, so a step captured in a CUDA graph keeps advancing and an LR scheduler can
write ``param_groups[0]["lr"]`` between replays.
"""
import ctypes

import torch

from . import _cabi
from ._cabi import check, ptr, stream_ptr

_MAX_TENSORS = 80     # kAdamMaxTensors in csrc/adam.cu


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params]
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        for p in params:
            _cabi.require_cuda(p)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("maskplanner_b200.optim.Adam handles contiguous fp32 parameters")
        self.param_groups = [dict(params=params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)]
        dev = params[0].device
        self.state = {p: dict(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)) for p in params}
        self._groups = [params[i:i + _MAX_TENSORS] for i in range(0, len(params), _MAX_TENSORS)]
        # per launch group: [step count, learning rate] (float) and the ticket counter of the kernel's last-CTA election
        self._scalars = [torch.zeros(2, dtype=torch.float32, device=dev) for _ in self._groups]
        self._tickets = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in self._groups]
        self._lr_on_device = None

    # -- torch.optim.Optimizer surface used by the reference loop ------------------------------------------------
    def zero_grad(self, set_to_none=True):
        for p in self.param_groups[0]["params"]:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_().zero_()

    def _sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:          # a host -> device scalar write, outside any captured region
            for s in self._scalars:
                s[1] = lr
            self._lr_on_device = lr

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        if not torch.cuda.is_current_stream_capturing():
            self._sync_lr()
        lib = _cabi.load()
        for params, scal, ticket in zip(self._groups, self._scalars, self._tickets):
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
                                        g["weight_decay"], ptr(scal), ptr(ticket), stream_ptr()), "mpb_adam_step_f32")

    def set_lr(self, lr):
        """Change the learning rate (also between CUDA-graph replays: the kernel reads it from device memory)."""
        self.param_groups[0]["lr"] = lr
        self._sync_lr()

    def state_dict(self):
        params = self.param_groups[0]["params"]
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group.update(params=list(range(len(params))), amsgrad=False, maximize=False)
        step = {id(p): self._scalars[i // _MAX_TENSORS][0].clone() for i, p in enumerate(params)}
        return {"state": {i: dict(step=step[id(p)], exp_avg=self.state[p]["exp_avg"], exp_avg_sq=self.state[p]["exp_avg_sq"])
                          for i, p in enumerate(params)},
                "param_groups": [group]}

    def load_state_dict(self, sd):
        params = self.param_groups[0]["params"]
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in sd["param_groups"][0]:
                self.param_groups[0][k] = sd["param_groups"][0][k]
        for i, p in enumerate(params):
            st = sd["state"].get(i)
            if st is None:
                continue
            self.state[p]["exp_avg"].copy_(st["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
            self._scalars[i // _MAX_TENSORS][0] = float(st["step"])
        self._lr_on_device = None
        self._sync_lr()
