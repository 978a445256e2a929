#!/bin/bash
# Step-level iteration on the GPU box: step/heads/mlp parity tests, then the quick bench with and without pipelined sampling.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step.py tests/test_gpu_heads.py tests/test_gpu_mlp.py -m gpu -x -q 2>&1 | tail -15
for P in 1 0; do
MPB_PIPELINE_SAMPLING=$P python bench.py --quick --no-cpu-baseline --steps 200 2>gpurun_out/bench_quick_$P.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipeline=$P ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'launches', d['gpu_launches'])"
tail -3 gpurun_out/bench_quick_$P.err
done
