"""The data-parallel plumbing of the Trainer on ONE GPU: a single-rank NCCL group with Trainer(world_size=2) runs the flat
gradient buckets, direct gradient placement, the head-bucket all-reduce on the communication stream and the 1/W folded into
Adam inside the captured graph.  Adam's update is invariant to the gradient scale (up to eps), so the losses must track a
plain single-GPU Trainer's.  Developer tool (GPU box):  python tools/dp_selfcheck.py [B] [steps]"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from maskplanner_b200 import synthetic
from maskplanner_b200.train_step import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_SOCKET_IFNAME", "lo")
store = tempfile.NamedTemporaryFile(delete=False)
dist.init_process_group("nccl", init_method="file://" + store.name, rank=0, world_size=1, device_id=dev)


def run(world_size):
    tr = Trainer("windows_v2", dev, use_graph=True, world_size=world_size, seed=0)
    res = [tr.to_device(synthetic.make_batch(B, "windows_v2", seed0=100 * i)) for i in range(3)]
    gen = torch.Generator().manual_seed(3)
    seeds = [(torch.randint(0, 5120, (B,), generator=gen), torch.randint(0, 512, (B,), generator=gen)) for _ in range(steps + 1)]
    out = []
    for i in range(steps):
        out.append(tr.step(res[i % 3], next_batch=res[(i + 1) % 3], fps_seeds=seeds[i] if i == 0 else None, next_fps_seeds=seeds[i + 1]).item())
    torch.cuda.synchronize()
    return out


plain = run(1)
dp = run(2)
print("plain", ["%.5f" % v for v in plain])
print("dp   ", ["%.5f" % v for v in dp])
worst = max(abs(a - b) / abs(a) for a, b in zip(plain, dp))
print("max relative difference %.2e" % worst)
dist.destroy_process_group()
assert all(v == v and abs(v) < 1e6 for v in dp), "non-finite loss under the data-parallel plumbing"
assert worst < 5e-2, "the data-parallel plumbing does not track the plain trainer"
print("dp selfcheck ok")
