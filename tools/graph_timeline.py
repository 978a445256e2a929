"""Kernel timeline of ONE replay of the captured training step (torch.profiler / CUPTI): start offset, duration, stream and
name of every kernel, plus per-stream busy time and the gaps on the critical stream.  Developer tool (GPU box).
    python tools/graph_timeline.py [precision] [B] > gpurun_out/graph_timeline.txt"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from maskplanner_b200 import pointnet2_utils as P
from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
P.set_mlp_precision(prec)
tr = Trainer("windows_v2", dev, use_graph=True)
res = [tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=100 * i)) for i in range(3)]
for i in range(8):
    tr.step(res[i % 3], next_batch=res[(i + 1) % 3])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        tr.step(res[i % 3], next_batch=res[(i + 1) % 3])
    torch.cuda.synchronize()
path = "gpurun_out/graph_trace.json"
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
ev.sort(key=lambda e: e["ts"])
# split into replays: a gap of > 100 us without any kernel never happens inside a step; use the adam kernel as the end marker
ends = [i for i, e in enumerate(ev) if "adam_kernel" in e["name"]]
lo = ends[0] + 1
hi = ends[1]
# the plan copy (cur <- next) follows Adam: extend to the last event before the next replay's first kernel
while hi + 1 < len(ev) and ev[hi + 1]["ts"] - (ev[hi]["ts"] + ev[hi]["dur"]) < 30 and "Memcpy" in ev[hi + 1]["name"]:
    hi += 1
one = ev[lo:hi + 1]
t0 = one[0]["ts"]
streams = sorted({e["args"].get("stream") for e in one})
print("one replay: %d kernels/copies, %.1f us from first start to last end, streams %s" % (len(one), max(e["ts"] + e["dur"] for e in one) - t0, streams))
busy = {}
for e in one:
    busy[e["args"].get("stream")] = busy.get(e["args"].get("stream"), 0.0) + e["dur"]
print("busy us per stream:", {k: round(v, 1) for k, v in busy.items()})
print("%9s %8s %6s  %s" % ("start_us", "dur_us", "stream", "kernel"))
for e in one:
    print("%9.1f %8.1f %6s  %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:110]))
