"""Turn an `ncu --set full` report into profiles/<tag>_ncu_top_kernels.csv and profiles/traffic.json.

    python tools/ncu_report.py gpurun_out/prof_r01.ncu-rep r01          # or the exported raw page (.csv)

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps, per profiled launch: duration,
DRAM bytes read / written, DRAM throughput (% of ncu's peak), tensor-pipe activity, L1/TEX and L2
throughput %, achieved warps, registers per thread and the grid.  traffic.json holds the per-kernel
average of dram read + write bytes per launch; bench.py copies the dominant kernel's entry into
`roofline.traffic`.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}

WANT = [  # (column suffix, output name, kind)
    ("gpu__time_duration.sum", "duration_us", "time"),
    ("dram__bytes_read.sum", "dram_read_MB", "bytes"),
    ("dram__bytes_write.sum", "dram_write_MB", "bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", "raw"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", "raw"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct", "raw"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct", "raw"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct", "raw"),
    ("launch__registers_per_thread", "regs", "raw"),
]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*", "", name)
    return name.replace("mpb::", "")


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
    if rep.endswith(".csv"):  # already exported on the GPU box with `ncu -i <rep> --page raw --csv`
        raw = "".join(l for l in open(rep) if not l.startswith("=="))
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {}
    for suffix, name, _ in WANT:
        col[name] = [i for i, h in enumerate(head) if h == suffix or h.endswith("." + suffix)]
    k_name, k_grid = head.index("Kernel Name"), head.index("Grid Size")
    lines, agg = [], collections.OrderedDict()
    for r in body:
        vals = {}
        for suffix, name, kind in WANT:
            c = next((i for i in col[name] if r[i] not in ("", "no data")), None)
            if c is None:
                vals[name] = float("nan")
                continue
            v = float(r[c].replace(",", ""))
            u = units[c]
            if kind == "time":
                v *= UNIT.get(u, 1.0)
            elif kind == "bytes":
                v *= UNIT.get(u, 1.0) / 1e6
            vals[name] = v
        kn = short(r[k_name])
        lines.append((kn, vals, r[k_grid].replace(",", " ")))
        base = re.sub(r"<.*", "", kn)
        keys = [base]
        m = re.match(r"((?:gemm_tn|wgrad)_kernel<\d)", kn)      # the GEMMs also per arithmetic: <0 = bf16 (shared MLP), <1 / <2 = TF32 / 3xTF32
        if m:
            keys.append(m.group(1))
        for key in keys:
            a = agg.setdefault(key, {"launches": 0, "bytes": 0.0, "us": 0.0})
            a["launches"] += 1
            a["bytes"] += (vals["dram_read_MB"] + vals["dram_write_MB"]) * 1e6
            a["us"] += vals["duration_us"]
    csv_path = os.path.join(out_dir, "%s_ncu_top_kernels.csv" % tag)
    with open(csv_path, "w") as f:
        f.write("# ncu --clock-control none --profile-from-start off (metrics of the columns below), one eager training step "
                "(tools/ncu_step.py, B=64, windows_v2); made by tools/ncu_report.py from %s\n" % os.path.basename(rep))
        f.write("# per launch (cold cache, serialised, ~40 replays): compare traffic and pipe shares, not absolutes\n")
        f.write("kernel," + ",".join(n for _, n, _ in WANT) + ",grid\n")
        for kn, vals, grid in lines:
            f.write('"%s",' % kn + ",".join("%.4g" % vals[n] for _, n, _ in WANT) + ',"%s"\n' % grid)
    traffic = {k: {"launches": a["launches"], "dram_bytes_per_launch": a["bytes"] / a["launches"],
                   "avg_duration_us": a["us"] / a["launches"],
                   "source": "profiles/%s_ncu_top_kernels.csv (ncu --set full, per launch, cold cache)" % tag}
               for k, a in agg.items()}
    with open(os.path.join(out_dir, "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("wrote", csv_path, "and traffic.json:", {k: round(v["dram_bytes_per_launch"] / 1e6, 1) for k, v in traffic.items()})


if __name__ == "__main__":
    main()
