"""Per-launch CUDA-event timing of the shared-MLP GEMM kernels inside one eager training step (B = 64, windows_v2):
shape, fusion flags, microseconds and algorithmic GB/s of every launch.  Developer tool (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maskplanner_b200 import pointnet2_utils as P
from maskplanner_b200 import shared_mlp, synthetic
from maskplanner_b200.train_step import Trainer

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
P.set_mlp_precision(prec)
tr = Trainer("windows_v2", dev, use_graph=False)
batch = tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=0))
gen = torch.Generator().manual_seed(1)
for it in range(4):
    seeds = (torch.randint(0, 5120, (B,), generator=gen).to(dev), torch.randint(0, 512, (B,), generator=gen).to(dev))
    if it == 3:
        shared_mlp.GEMM_TIMELINE = []
    tr._step_core(batch, seeds)
torch.cuda.synchronize()
tl, shared_mlp.GEMM_TIMELINE = shared_mlp.GEMM_TIMELINE, None
tot = {}
for name, nbytes, flops, e0, e1, tag in tl:
    ms = e0.elapsed_time(e1)
    print("%-15s %-42s %8.1f us %8.0f GB/s %7.1f TFLOP/s" % (name, tag, ms * 1e3, nbytes / ms / 1e6, flops / ms / 1e9))
    t = tot.setdefault(name, [0.0, 0])
    t[0] += ms
    t[1] += nbytes
for k, (ms, nb) in tot.items():
    print("TOTAL %-15s %8.1f us %8.0f GB/s" % (k, ms * 1e3, nb / ms / 1e6))
