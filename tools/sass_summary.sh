#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMA / TMEM path (B200_PROFILING.md), from the built library.
LIB=maskplanner_b200/_lib/libmaskplanner_b200.so
OUT=${1:-profiles/r02_sass_summary.txt}
cuobjdump -sass $LIB > /tmp/mpb_sass.txt
python - "$OUT" <<'PY'
import re, sys, collections
want = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "FMUL2", "REDUX", "RED.E", "ATOMG"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in open("/tmp/mpb_sass.txt"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        counts[cur]["_instr"] += 1
        for w in want:
            if w in line:
                counts[cur][w] += 1
                total[w] += 1
import subprocess
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
with open(sys.argv[1], "w") as f:
    f.write("# cuobjdump -sass of libmaskplanner_b200.so (sm_100a only): SASS mnemonic counts per kernel; made by tools/sass_summary.sh\n")
    f.write("# UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA load/store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier\n")
    f.write("TOTAL " + " ".join("%s=%d" % (w, total[w]) for w in want if total[w]) + "\n")
    for k, c in counts.items():
        hits = " ".join("%s=%d" % (w, c[w]) for w in want if c[w])
        f.write("%-90s instr=%-6d %s\n" % (demangle(k)[:90], c["_instr"], hits))
print(open(sys.argv[1]).read()[:1500])
PY
