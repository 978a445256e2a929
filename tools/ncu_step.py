"""One eager training step between cudaProfilerStart/Stop for ncu (--profile-from-start off)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer
B = int(os.environ.get("B", "64"))
dev = torch.device("cuda", 0)
tr = Trainer("windows_v2", dev)
batches = [tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=100 * i)) for i in range(2)]
for i in range(3):
    tr.step(batches[i % 2])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(batches[1])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
