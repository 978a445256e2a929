"""Repeat the same training step (fixed batch, fixed FPS seeds, lr = 0) and report how often the loss deviates.
Last-bit differences from atomics are expected (rel ~1e-7); anything larger is a race."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer
B = int(os.environ.get("B", "4"))
N = int(os.environ.get("STEPS", "300"))
use_graph = os.environ.get("GRAPH", "0") == "1"
dev = torch.device("cuda", 0)
tr = Trainer("windows_v2", dev, seed=2, use_graph=use_graph, lr=0.0)
tr.model.dropout.p = 0.0
batches = [tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=50 + 10 * i)) for i in range(3)]
gen = torch.Generator().manual_seed(9)
seeds = [(torch.randint(0, 5120, (B,), generator=gen), torch.randint(0, 512, (B,), generator=gen)) for _ in range(3)]
vals = collections.defaultdict(list)
for i in range(N):
    j = i % 3
    vals[j].append(float(tr.step(batches[j], seeds[j]).item()))
bad = 0
for j, v in vals.items():
    ref = sorted(v)[len(v) // 2]
    dev_ = [abs(x - ref) / abs(ref) for x in v]
    nb = sum(d > 1e-5 for d in dev_)
    bad += nb
    print("batch %d: median %.4f max rel dev %.3e outliers(>1e-5) %d/%d first outlier idx %s" % (
        j, ref, max(dev_), nb, len(v), next((k for k, d in enumerate(dev_) if d > 1e-5), None)))
print("TOTAL outliers", bad)
