#!/bin/bash
# Step-level iteration on the GPU box: step/heads parity tests, then the quick bench under A/B environment switches (arguments).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_heads.py -m gpu -x -q 2>&1 | grep -E "passed|failed|Error" | tail -4
if [ $# -eq 0 ]; then set -- MPB_NOOP=1; fi
for V in "$@"; do
env $V python bench.py --quick --no-cpu-baseline --steps 200 2>gpurun_out/bench_quick.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'launches', d['gpu_launches'])"
done
