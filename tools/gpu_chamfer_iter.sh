#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chamfer.py tests/test_gpu_step.py tests/test_gpu_extra.py -m gpu -x -q 2>&1 | grep -E "passed|failed|FAILED|^E " | head -8
for V in MPB_CHAMFER_SPLIT=1 MPB_CHAMFER_SPLIT=0; do
  env $V python bench.py --workload chamfer 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$V', {k: (round(v['fwd_ms'], 4), round(v.get('frac_of_fp32_peak', 0), 3)) for k, v in d['kernels'].items() if k.startswith('mp_')})"
  env $V python bench.py --quick --no-cpu-baseline --steps 200 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
