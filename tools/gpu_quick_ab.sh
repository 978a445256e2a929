#!/bin/bash
# Short GPU check: the whole GPU suite, then the quick bench under A/B environment switches (arguments).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|Error|error" | tail -6
if [ $# -eq 0 ]; then set -- MPB_NOOP=1; fi
for V in "$@"; do
env $V timeout 120 python bench.py --quick --no-cpu-baseline --steps 200 2>gpurun_out/bench_quick.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'launches', d['gpu_launches'])" || tail -5 gpurun_out/bench_quick.err
done
