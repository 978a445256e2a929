"""Per-rank CUPTI timeline of ONE graph replay under data parallelism (rank 0 prints): where the NCCL all-reduce kernels sit
relative to the encoder backward and Adam.  torchrun --nproc-per-node N tools/dp_timeline.py [per_rank_batch]"""
import json
import os
import sys

os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer

rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
tr = Trainer("windows_v2", dev, world_size=ws, use_graph=True)
res = [tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=1000 * rank + 10 * i)) for i in range(3)]
for i in range(10):
    tr.step(res[i % 3], next_batch=res[(i + 1) % 3])
dist.barrier()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(4):
        tr.step(res[i % 3], next_batch=res[(i + 1) % 3])
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    path = "/tmp/dp_trace.json"
    prof.export_chrome_trace(path)
    ev = sorted([e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e],
                key=lambda e: e["ts"])
    ends = [i for i, e in enumerate(ev) if "adam_kernel" in e["name"]]
    one = ev[ends[1] + 1:ends[2] + 1]
    t0 = one[0]["ts"]
    print("N = %d, %d samples per rank: one replay on rank 0 = %d kernels/copies, %.1f us first start -> Adam end"
          % (ws, B, len(one), one[-1]["ts"] + one[-1]["dur"] - t0))
    nccl = [e for e in one if "nccl" in e["name"].lower()]
    for e in nccl:
        print("NCCL  start %8.1f us  dur %8.1f us  stream %s  %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:80]))
    marks = [e for e in one if any(k in e["name"] for k in ("head_act_bwd", "bwd_stats_pooled", "narrow_first_layer_bwd", "adam_kernel", "loss_bwd_masks"))]
    for e in marks:
        print("mark  start %8.1f us  dur %8.1f us  stream %s  %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:60]))
    print("%9s %8s %6s  %s" % ("start_us", "dur_us", "stream", "kernel"))
    for e in one:
        print("%9.1f %8.1f %6s  %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:100]))
sys.stdout.flush()
os._exit(0)
