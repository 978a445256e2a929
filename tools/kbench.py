"""Developer micro-benchmark of the individual kernels (CUDA events, L2 flushed between iterations).
Writes gpurun_out/kbench.json.  Not the contract bench (that is bench.py)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskplanner_b200 import pointnet2_utils as P
from maskplanner_b200 import pytorch3d_chamfer as CH
from maskplanner_b200 import synthetic

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {"ms_min": ts[0], "ms_med": ts[len(ts) // 2]}


res = {}


def rec(name, r, **kw):
    r.update(kw)
    res[name] = r
    print(name, json.dumps(r), flush=True)


# ---- FPS ----
for B, N, S, kind in [(32, 5120, 1024, "cube"), (32, 5120, 1024, "cuboid"), (64, 5120, 512, "cuboid"), (64, 512, 128, "cube"), (8, 100000, 4096, "cube")]:
    xyz = synthetic.make_clouds(B, N, seed0=1000, kind=kind).to(dev)
    seed = torch.zeros(B, dtype=torch.long, device=dev)
    r = timeit(lambda: P.farthest_point_sample(xyz, S, seed_idx=seed), iters=5 if N > 10000 else 10)
    stream_bytes = B * S * N * 20
    rec("fps_B%d_N%d_S%d_%s" % (B, N, S, kind), r, stream_GBps=stream_bytes / r["ms_med"] / 1e6, us_per_sample=r["ms_med"] * 1e3 / S)

# ---- ball query + grouping ----
for B, N, S, rad, K, kind in [(32, 5120, 1024, 0.2, 32, "cube"), (32, 5120, 1024, 0.2, 32, "cuboid"), (64, 5120, 512, 0.2, 32, "cuboid"), (64, 512, 128, 0.4, 64, "cuboid")]:
    xyz = synthetic.make_clouds(B, N, seed0=1000, kind=kind).to(dev)
    idx = P.farthest_point_sample(xyz, S, seed_idx=torch.zeros(B, dtype=torch.long))
    new_xyz = P.index_points(xyz, idx)
    r = timeit(lambda: P.query_ball_point(rad, K, xyz, new_xyz))
    rec("ball_B%d_N%d_S%d_K%d_%s" % (B, N, S, K, kind), r, gpairs_per_s=B * S * N / r["ms_med"] / 1e6)
    ball = P.query_ball_point(rad, K, xyz, new_xyz)
    D = 0 if N == 5120 else 128
    feats = torch.randn(B, N, D, device=dev) if D else None
    r = timeit(lambda: P.group_points(xyz, feats, new_xyz, ball))
    rec("group_B%d_S%d_K%d_D%d" % (B, S, K, D), r, GBps=2 * B * S * K * (3 + D) * 4 / r["ms_med"] / 1e6)

# ---- kNN grouping (stress) ----
xyz = synthetic.make_clouds(8, 100000, seed0=1000, kind="cube").to(dev)
q = xyz[:, :4096].contiguous()
r = timeit(lambda: P.knn_group(32, xyz, q), iters=3, warm=1)
rec("knn_group_B8_N100000_S4096_k32", r, gpairs_per_s=8 * 4096 * 100000 / r["ms_med"] / 1e6)

# ---- chamfer ----
def cham_case(name, B, P1, P2, D, kw, iters=10):
    x = torch.randn(B, P1, D, device=dev, requires_grad=True)
    y = torch.randn(B, P2, D, device=dev)
    r = timeit(lambda: CH.chamfer_distance(x, y, **kw), iters=iters)
    ndir = 2 if (kw.get("return_matching") or not (kw.get("asymmetric") or kw.get("reverse_asymmetric"))) else 1
    pairs = B * P1 * P2 * ndir
    rec(name + "_fwd", r, gpairs_per_s=pairs / r["ms_med"] / 1e6, tflops=pairs * 3 * D / r["ms_med"] / 1e9)

    def fb():
        x.grad = None
        CH.chamfer_distance(x, y, **kw)[0].sum().backward()
    r = timeit(fb, iters=iters)
    rec(name + "_fwdbwd", r)


cham_case("cham_mp_call1_B64_999x986x24", 64, 999, 986, 24, dict(asymmetric=True, return_matching=True, point_reduction=None, batch_reduction=None))
cham_case("cham_mp_call2_B64_3996x2959x6", 64, 3996, 2959, 6, dict(reverse_asymmetric=True))
cham_case("cham_mp_call3_B64_999x986x24", 64, 999, 986, 24, dict(reverse_asymmetric=True))
cham_case("cham_win_call1_B64_449x449x24", 64, 449, 449, 24, dict(asymmetric=True, return_matching=True, point_reduction=None, batch_reduction=None))
cham_case("cham_win_call2_B64_1796x1350x6", 64, 1796, 1350, 6, dict(reverse_asymmetric=True))
for Pn in (2048, 4096, 8192, 16384, 32768, 65536):
    cham_case("cham_sweep_B32_%dx%dx3_asym" % (Pn, Pn), 32, Pn, Pn, 3, dict(asymmetric=True), iters=3 if Pn >= 32768 else 5)
for Pn in (2048, 8192):
    cham_case("cham_sweep_B32_%dx%dx24_asym" % (Pn, Pn), 32, Pn, Pn, 24, dict(asymmetric=True), iters=3)
    cham_case("cham_sweep_B32_%dx%dx6_sym" % (Pn, Pn), 32, Pn, Pn, 6, dict(), iters=3)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/kbench.json", "w"), indent=1)
