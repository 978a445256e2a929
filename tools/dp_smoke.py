"""2-rank DP smoke test with progress markers (run under torchrun + timeout)."""
import os, sys, time, faulthandler
faulthandler.enable()
os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", os.environ.get("ASYNC_EH", "0"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def log(*a):
    print("[r%d %.1f]" % (rank, time.time() % 1000), *a, file=sys.stderr, flush=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
log("pg up")
use_graph = os.environ.get("GRAPH", "1") == "1"
B = 16
tr = Trainer("windows_v2", dev, world_size=ws, use_graph=use_graph)
tr.model.dropout.p = 0.0
batches = [tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=100 * rank + 7 * j)) for j in range(2)]
gen = torch.Generator().manual_seed(5)
seeds = [(torch.randint(0, 5120, (B,), generator=gen), torch.randint(0, 512, (B,), generator=gen)) for _ in range(9)]
for i in range(8):
    # pipelined sampling: the next batch (and its FPS seeds) is announced with every call
    loss = tr.step(batches[i % 2], seeds[i] if i == 0 else None, next_batch=batches[(i + 1) % 2], next_fps_seeds=seeds[i + 1])
    torch.cuda.synchronize()
    log("step", i, float(loss))
# parameters must stay identical across ranks
flat = torch.cat([p.detach().flatten() for p in tr.model.parameters()])
other = flat.clone()
dist.broadcast(other, src=0)
log("param max diff vs rank0", float((flat - other).abs().max()))
log("direct_grads", tr.direct_grads, "param checksum %.9e %.9e" % (float(flat.double().sum()), float(flat.double().abs().sum())))
dist.barrier()
torch.cuda.synchronize()
log("done")
tr._graph = None                      # drop the captured NCCL work before the communicator goes away
try:
    dist.destroy_process_group()
except Exception as e:               # teardown only
    log("destroy_process_group:", e)
log("exit")
