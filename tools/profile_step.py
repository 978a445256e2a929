"""Developer profile of the training step: torch.profiler kernel table + wall/GPU time split."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer
B = int(os.environ.get("B", "64"))
dev = torch.device("cuda", 0)
tr = Trainer("windows_v2", dev)
batches = [tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=100 * i)) for i in range(2)]
for i in range(4):
    tr.step(batches[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(5):
    tr.step(batches[i % 2])
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        tr.step(batches[i % 2])
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev if e.device_type == torch.autograd.DeviceType.CUDA) if False else None
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/step_trace.json")
