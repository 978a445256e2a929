"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
tot = sum(ns for _, ns in rows)
agg = collections.OrderedDict()
for n, ns in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += ns
print("kernel launches in one training step: %d, summed device time %.3f ms (cold-cache, serialised: compare SHARES)" % (len(rows), tot / 1e6))
print("%-72s %8s %10s %8s" % ("kernel", "launches", "total_us", "share"))
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %8d %10.1f %7.2f%%" % (n[:72], c, ns / 1e3, 100 * ns / tot))
