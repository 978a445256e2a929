#!/usr/bin/env python
"""bench.py -- the driver contract bench for the MaskPlanner point-cloud hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

Workload (BASELINE.json metric "train samples/s", configs[3]): the full MaskPlanner training step --
PointNet++ SSG encoder (5120 pts -> 512 -> 128 -> global), windows_v2 heads (449 segments x 24, 22 masks,
26.2 M parameters), asymm_v6 chamfer + stroke-mask loss, backward, Adam -- on synthetic PaintNet-shaped
data, B = 64 samples per GPU, batch-sharded data parallel with one NCCL gradient all-reduce per step.
One "step" = one optimisation step over one batch.  Prints ONE JSON line (rank 0).

Other workloads (--workload sa_micro | chamfer) time BASELINE.json configs[1] / configs[2] kernels only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "sa_micro", "chamfer"])
    ap.add_argument("--category", default="windows_v2")
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU (weak scaling)")
    ap.add_argument("--ref-batch", type=int, default=None,
                    help="samples per CPU-reference step (default: --batch for --impl reference, 16 for the inline cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying the captured step")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for l in self.proc.stdout:
            self.lines.append((time.monotonic(), l.strip()))

    def stop(self, t0=None, t1=None):
        """Summarise the samples that arrived inside [t0, t1] (host monotonic clock around the timed region; the
        sampler is started during warm-up so nvidia-smi's start-up latency does not eat the region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        inside = [l for t, l in self.lines if t0 is not None and t0 <= t <= t1 + 0.03]
        sm, mx, reasons = [], [], set()
        for l in (inside or [l for _, l in self.lines[-5:]]):
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(args):
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return ws, rank, local


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port; the reference is Python and cannot travel)
# --------------------------------------------------------------------------------------------------
def reference_step_time(category, ref_batch, steps, warmup):
    from maskplanner_b200 import synthetic
    from oracle import c_oracle
    from oracle import step_oracle as SO
    torch.set_num_threads(os.cpu_count())          # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    c_oracle.set_threads(os.cpu_count())
    torch.manual_seed(0)
    cfg = synthetic.CATEGORIES[category]
    model = SO.Regressor(synthetic.out_vectors(cfg["n_pred_traj_points"]), n_stroke_masks=cfg["max_n_strokes"])
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    batches = [synthetic.make_batch(ref_batch, category, seed0=100 * i) for i in range(2)]
    for i in range(warmup):
        SO.train_step(model, opt, batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        SO.train_step(model, opt, batches[i % 2])
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt


def run_reference(args, ws, rank):
    if rank != 0:
        return
    args.ref_batch = args.ref_batch or args.batch      # the arm's own config: B = 64 per step (~3-4 s of CPU work each)
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    dt = reference_step_time(args.category, args.ref_batch, steps, warm)
    v = args.ref_batch / dt
    sample = "%d timed + %d warm-up optimisation steps of the oracle port at B=%d (same model/loss/data generator)" % (steps, warm, args.ref_batch)
    line = {"impl": "reference", "metric": "train samples/s", "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, args.gpus, cpu=True),
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def workload_config(args, n, cpu=False):
    return {"workload": "MaskPlanner training step [maskplanner,%s,longx_v2]: PointNet++ SSG encoder 5120->512->128 + heads "
                        "+ asymm_v6 chamfer/stroke-mask loss + backward + Adam" % args.category,
            "per_gpu_batch": args.ref_batch if cpu else args.batch, "global_batch": (args.ref_batch if cpu else args.batch * n),
            "pc_points": 5120, "parallelism": "cpu" if cpu else "dp%d" % n,
            "launch": "cpu threads" if cpu else ("eager (one launch per kernel)" if args.no_graph else "whole step captured once, one CUDA-graph replay per step"),
            "arithmetic": "fp32 (reference CPU path)" if cpu else
                          "shared-MLP GEMMs bf16 on tcgen05 with fp32 accumulation/statistics; head GEMMs TF32 (cuBLAS); "
                          "FPS, ball query, grouping, chamfer, mask loss, Adam fp32",
            "l2": "no flush: per-step working set (activations + 0.42 GB of parameter/optimizer state) exceeds the 126 MB L2"}


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def kernel_table(args, dev, peak):
    """FPS + grouping + chamfer kernel times at the step's own shapes, each timed alone with CUDA events
    on the launching stream (L2 flushed between iterations), with algorithmic GB/s."""
    from maskplanner_b200 import pointnet2_utils as P
    from maskplanner_b200 import pytorch3d_chamfer as CH
    from maskplanner_b200 import synthetic
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    xyz = synthetic.make_clouds(B, 5120, seed0=1000).to(dev)
    seed = torch.zeros(B, dtype=torch.long, device=dev)
    out = {}
    t = time_kernel(lambda: P.farthest_point_sample(xyz, 512, seed_idx=seed), flush=flush)
    fps_bytes = B * 512 * 5120 * 20
    out["fps_sa1"] = {"ms": t, "algorithmic_GBps": fps_bytes / t / 1e6, "frac_of_hbm_peak": fps_bytes / t / 1e6 / peak}
    idx = P.farthest_point_sample(xyz, 512, seed_idx=seed)
    new_xyz = P.index_points(xyz, idx)
    t = time_kernel(lambda: P.query_ball_point(0.2, 32, xyz, new_xyz), flush=flush)
    out["ball_query_sa1"] = {"ms": t, "gpairs_per_s": B * 512 * 5120 / t / 1e6}
    ball = P.query_ball_point(0.2, 32, xyz, new_xyz)
    # grouping as the step runs it: bf16 GEMM rows [B*S*K, pad64(3+D)] (gathered fp32 reads + bf16 row writes)
    t = time_kernel(lambda: P._GroupPointsBF16.apply(xyz, None, new_xyz, ball, 64), flush=flush)
    gb = B * 512 * 32 * (3 * 4 + 64 * 2)
    out["group_sa1"] = {"ms": t, "algorithmic_GBps": gb / t / 1e6, "frac_of_hbm_peak": gb / t / 1e6 / peak}
    f2 = torch.randn(B, 512, 128, device=dev)
    x2 = new_xyz
    idx2 = P.farthest_point_sample(x2, 128, seed_idx=seed)
    nx2 = P.index_points(x2, idx2)
    ball2 = P.query_ball_point(0.4, 64, x2, nx2)
    t = time_kernel(lambda: P._GroupPointsBF16.apply(x2, f2, nx2, ball2, 192), flush=flush)
    gb = B * 128 * 64 * (131 * 4 + 192 * 2)
    out["group_sa2"] = {"ms": t, "algorithmic_GBps": gb / t / 1e6, "frac_of_hbm_peak": gb / t / 1e6 / peak}
    cfg = synthetic.CATEGORIES[args.category]
    ov = synthetic.out_vectors(cfg["n_pred_traj_points"])
    d = synthetic.make_trajectories(B, args.category)
    y = d["traj"].to(dev)
    x = synthetic.noisy_predictions(d["traj"], ov).to(dev)
    t = time_kernel(lambda: CH.chamfer_distance(x, y, padded=True, asymmetric=True, return_matching=True, point_reduction=None,
                                                batch_reduction=None), flush=flush)
    pairs = 2 * B * ov * y.shape[1]
    out["chamfer_segments_fwd"] = {"ms": t, "gpairs_per_s": pairs / t / 1e6, "fp32_tflops": pairs * 3 * 24 / t / 1e9}
    return out, fps_bytes


def run_ours(args, ws, rank, local):
    import torch.distributed as dist
    from maskplanner_b200 import _cabi, synthetic
    from maskplanner_b200.train_step import Trainer, pin_batch
    if not torch.cuda.is_available():
        sys.exit("bench.py (impl=ours) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        # NCCL collectives are captured inside the step's CUDA graph: the watchdog thread must not poll CUDA
        # events while a capture is open (PyTorch's documented requirement for whole-network capture)
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    _cabi.load()
    peak, peak_src = peaks()
    B = args.batch
    trainer = Trainer(args.category, dev, world_size=ws, use_graph=not args.no_graph)
    args.warmup = max(args.warmup, 3) if not args.no_graph else args.warmup   # 2 eager steps + the capture step
    # distinct synthetic batches per rank (weak scaling: every rank owns B whole samples)
    host = [pin_batch(synthetic.make_batch(B, args.category, seed0=10000 * rank + 100 * i)) for i in range(3)]
    resident = [trainer.to_device(h) for h in host]
    torch.cuda.synchronize()

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        trainer.step(resident[i % 3])
    barrier()
    l0 = _cabi.KERNEL_LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.monotonic()
    a.record()
    for i in range(args.steps):
        loss = trainer.step(resident[i % 3])
    b.record()
    barrier()
    t_region1 = time.monotonic()
    ms = a.elapsed_time(b)
    launches = getattr(trainer, "kernels_per_step", None) or (_cabi.KERNEL_LAUNCHES - l0) // max(args.steps, 1)
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = B * ws / (ms_step / 1e3)

    # end to end through the public step API: pinned host batch -> H2D -> step -> loss.item() (D2H); every step's batch is
    # copied inside the timed region (the copy of batch i+1 is issued while step i runs, like a prefetching loader)
    for i in range(3):          # warm the end-to-end path (side-stream allocator pool, pinned-copy plumbing) before timing it
        trainer.step_from_host(host[i % 3], next_host_batch=host[(i + 1) % 3])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        lv = trainer.step_from_host(host[i % 3], next_host_batch=host[(i + 1) % 3])   # next batch's H2D overlaps this step
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_step = float(t.item()) / args.steps
    h2d = sum(v.numel() * 4 for k, v in host[0].items() if torch.is_tensor(v) and k in ("point_cloud", "traj", "traj_as_pc", "stroke_ids"))
    h2d += 2 * B * 8  # the two FPS seed vectors (int64), drawn on the host like the reference (:77)

    # roofline leg: the same step, eager, with CUDA events around every launch of the two GEMM kernels
    # (they are the largest share of the step, profiles/); events are recorded on the launching stream
    from maskplanner_b200 import shared_mlp
    shared_mlp.GEMM_TIMELINE = []
    gen = torch.Generator().manual_seed(1)
    for i in range(3):
        seeds = (torch.randint(0, 5120, (B,), generator=gen), torch.randint(0, 512, (B,), generator=gen))
        trainer._step_core(resident[i % 3], tuple(s.to(dev) for s in seeds))
    torch.cuda.synchronize()
    timeline, shared_mlp.GEMM_TIMELINE = shared_mlp.GEMM_TIMELINE, None
    per_kernel = {}
    for name, nbytes, flops, e0, e1 in timeline:
        d = per_kernel.setdefault(name, {"launches": 0, "bytes": 0, "flops": 0, "ms": 0.0})
        d["launches"] += 1
        d["bytes"] += nbytes
        d["flops"] += flops
        d["ms"] += e0.elapsed_time(e1)

    if rank == 0:
        ktab, fps_bytes = kernel_table(args, dev, peak)
        for name, d in per_kernel.items():
            ktab[name] = {"launches_per_step": d["launches"] // 3, "ms_per_step": d["ms"] / 3, "avg_launch_us": d["ms"] / d["launches"] * 1e3,
                          "algorithmic_GBps": d["bytes"] / d["ms"] / 1e6, "frac_of_hbm_peak": d["bytes"] / d["ms"] / 1e6 / peak,
                          "tflops": d["flops"] / d["ms"] / 1e9}
        top = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
        d = per_kernel[top]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(top, {}).get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "kernel": "%s (tcgen05/TMA shared-MLP GEMM, %d launches per step)" % (top, d["launches"] // 3),
                "achieved": d["bytes"] / d["ms"] / 1e6, "peak": peak, "unit": "GB/s", "frac": d["bytes"] / d["ms"] / 1e6 / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": d["bytes"] / d["launches"],
                "avg_launch_us": d["ms"] / d["launches"] * 1e3,
                "note": "achieved = algorithmic bytes (bf16 operands read once + output written once, DESIGN.md) summed over the "
                        "kernel's launches of one step / summed CUDA-event durations of those launches, eager replay of the timed step; "
                        "arithmetic intensity 21-85 flop/B < machine balance 258 flop/B, so HBM is the binding roof; "
                        "traffic = ncu dram bytes per launch (profiles/traffic.json); peak = " + peak_src}
        line = {"metric": "train samples/s", "value": value, "unit": "samples/s", "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(args, ws), "clocks": clocks,
                "e2e": {"value": B * ws / e2e_step, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_step * 1e3},
                "gpu_launches": launches, "roofline": roof, "kernels": ktab, "final_loss": float(lv)}
        if not args.no_cpu_baseline and ws == 1:       # reported at N = 1 only
            args.ref_batch = args.ref_batch or 16
            dt = reference_step_time(args.category, args.ref_batch, 1, 1)
            line["cpu_baseline"] = {"value": args.ref_batch / dt, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "1 timed + 1 warm-up optimisation step of the oracle port at B=%d on the host cores" % args.ref_batch}
        emit(line)
    if ws > 1:
        dist.barrier()
        torch.cuda.synchronize()
        trainer._graph = None          # drop the captured NCCL work before the communicator goes away
        sys.stdout.flush()
        try:
            dist.destroy_process_group()
        except Exception as e:        # teardown only; the measurement is already printed
            print("destroy_process_group: %s" % e, file=sys.stderr)


def _claim_stdout():
    """Reserve the process's stdout for the ONE JSON line: fd 1 is pointed at stderr for everything else
    (NCCL's version banner, library chatter), and print() of the result goes to the saved descriptor."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    args = parse()
    _claim_stdout()
    ws, rank, local = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, ws, rank)
        return
    if args.workload != "train":
        sys.exit("workloads sa_micro/chamfer: use tools/kbench.py (kernel-only developer bench)")
    run_ours(args, ws, rank, local)


if __name__ == "__main__":
    main()
