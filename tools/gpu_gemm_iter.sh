#!/bin/bash
# GEMM iteration loop on the GPU box: validation, MLP parity tests, in-step per-launch timeline, quick bench.
mkdir -p gpurun_out
timeout 600 python tools/check_gemm.py ${1:-all} > gpurun_out/check_gemm.txt 2>&1; echo "check_gemm rc=$?"; grep -c " ok " gpurun_out/check_gemm.txt; grep -i "fail\|error\|Traceback" gpurun_out/check_gemm.txt | head -20
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_encoder.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | grep -v "Warning\|warn" | grep -B2 -A12 "Error\|^E \|passed\|failed" | tail -40
python tools/gemm_timeline.py bf16 64 > gpurun_out/gemm_timeline_bf16.txt 2>&1; tail -27 gpurun_out/gemm_timeline_bf16.txt
python bench.py --quick --no-cpu-baseline --steps 200 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'])"
