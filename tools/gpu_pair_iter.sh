#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/check_gemm.py tn 0 > gpurun_out/check_gemm.txt 2>&1; echo "check_gemm rc=$?"; grep -c " ok " gpurun_out/check_gemm.txt; grep -i "fail\|error\|Traceback" gpurun_out/check_gemm.txt | head -20
python - <<'PY'
import os, sys
sys.path.insert(0, "tools")
sys.argv = ["gemm_ab.py", "none"]
import gemm_ab as g
for pair in ("1", "0"):
    print("MPB_GEMM_PAIR=" + pair, "(needs a fresh process to change: see below)")
    break
g.tn(524288, 256, 128, 1, 1)
g.tn(524288, 256, 128, 0, 1)
g.tn(8192, 256, 320, 0, 1)
PY
MPB_GEMM_PAIR=0 python - <<'PY'
import sys
sys.path.insert(0, "tools")
sys.argv = ["gemm_ab.py", "none"]
import gemm_ab as g
print("MPB_GEMM_PAIR=0")
g.tn(524288, 256, 128, 1, 1)
g.tn(524288, 256, 128, 0, 1)
PY
