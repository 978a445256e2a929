"""Host/device timeline of the end-to-end call sequence (Trainer.step_from_host_async): GPU idle time between consecutive
replays and the host-side operations of one call.  Developer tool (GPU box)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer, pin_batch

B = 64
dev = torch.device("cuda", 0)
tr = Trainer("windows_v2", dev, use_graph=True)
host = [pin_batch(synthetic.make_batch(B, "windows_v2", seed0=100 * i)) for i in range(3)]


def run(n, i0=0):
    pending = None
    for i in range(i0, i0 + n):
        h = tr.step_from_host_async(host[i % 3], next_host_batch=host[(i + 1) % 3], after_next_host_batch=host[(i + 2) % 3])
        if pending is not None:
            pending.result()
        pending = h
    pending.result()
    torch.cuda.synchronize()


run(12)
t0 = time.perf_counter()
run(60, 12)
print("e2e ms/step (no profiler): %.3f" % ((time.perf_counter() - t0) / 60 * 1e3))
res = [tr.to_device(h) for h in host]
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(60):
    tr.step(res[i % 3], next_batch=res[(i + 1) % 3])
torch.cuda.synchronize()
print("resident ms/step (no profiler): %.3f" % ((time.perf_counter() - t0) / 60 * 1e3))
# host cost of one call, GPU not waited for
torch.cuda.synchronize()
t0 = time.perf_counter()
hs = [tr.step_from_host_async(host[i % 3], next_host_batch=host[(i + 1) % 3], after_next_host_batch=host[(i + 2) % 3]) for i in range(72, 76)]
print("host ms per step_from_host_async call (4 calls, no wait): %.3f" % ((time.perf_counter() - t0) / 4 * 1e3))
for h in hs:
    h.result()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(4, 76)
path = "gpurun_out/e2e_trace.json"
prof.export_chrome_trace(path)
tev = json.load(open(path))["traceEvents"]
gpu = sorted([e for e in tev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e], key=lambda e: e["ts"])
adam = [e for e in gpu if "adam_kernel" in e["name"]]
stage = [e for e in gpu if "stage_batch" in e["name"]]
for a in adam[:-1]:
    nxt = [s for s in stage if s["ts"] > a["ts"]]
    if nxt:
        print("adam end -> next stage_batch start: %.1f us" % (nxt[0]["ts"] - (a["ts"] + a["dur"])))
h2d = [e for e in gpu if "HtoD" in e["name"]]
print("H2D copies in window: %d, total %.1f us" % (len(h2d), sum(e["dur"] for e in h2d)))
cpu = sorted([e for e in tev if e.get("cat") in ("cpu_op", "cuda_runtime", "user_annotation") and "ts" in e and e.get("dur", 0) > 15], key=lambda e: e["ts"])
t_first = stage[1]["ts"] if len(stage) > 1 else 0
print("host ops > 15 us around the second call:")
for e in cpu:
    if t_first - 3500 < e["ts"] < t_first + 500:
        print("  %9.1f %8.1f %s" % (e["ts"] - t_first, e["dur"], e["name"][:80]))
os.remove(path)
