"""Fork/join of independent sub-graphs onto a side CUDA stream.

The training step has a few places where two chains of small kernels do not depend on each other (the pose head
vs. the stroke-mask head of the regressor: M = batch-size GEMMs on a handful of CTAs; chamfer term 2 vs. the
Hungarian-matched mask loss: a 64-warp solver next to a few hundred small CTAs).  Issuing one of them on a side
stream lets the GPU run both at once; inside CUDA-graph capture the fork and join become parallel branches of the
graph.  Autograd replays each backward op on the stream of its forward op, so the backward pass overlaps too.
MPB_PARALLEL_BRANCHES=0 turns it off (A/B runs, debugging).
"""
import os

import torch

_SIDE = {}


def enabled():
    return os.environ.get("MPB_PARALLEL_BRANCHES", "1") == "1"


def side_stream(device, slot=0):
    key = (torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), slot)
    s = _SIDE.get(key)
    if s is None:
        s = _SIDE[key] = torch.cuda.Stream(device=key[0])
    return s


class Fork:
    """with Fork(inputs...) as f: <ops on the side stream>;  <ops on the current stream>;  f.join(outputs...)"""

    def __init__(self, *inputs, slot=0):
        self.inputs = [t for t in inputs if torch.is_tensor(t) and t.is_cuda]
        self.active = enabled() and len(self.inputs) > 0
        self.slot = slot

    def __enter__(self):
        if not self.active:
            return self
        self.main = torch.cuda.current_stream()
        self.side = side_stream(self.inputs[0].device, self.slot)
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.side.wait_event(ev)
        for t in self.inputs:
            t.record_stream(self.side)
        self.ctx = torch.cuda.stream(self.side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.active:
            self.ctx.__exit__(*exc)
        return False

    def join(self, *outputs):
        if not self.active:
            return
        self.main.wait_stream(self.side)
        for t in outputs:
            if torch.is_tensor(t):
                t.record_stream(self.main)
