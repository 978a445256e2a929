#!/bin/bash
# Full GPU iteration: all parity tests, quick bench, graph timeline.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "Warning\|warn" | tail -8
python bench.py --quick --no-cpu-baseline --steps 200 2>gpurun_out/bench_quick.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], 'launches', d['gpu_launches'])"
tail -3 gpurun_out/bench_quick.err
python tools/graph_timeline.py bf16 64 > gpurun_out/graph_timeline.txt 2> gpurun_out/graph_timeline.err; head -3 gpurun_out/graph_timeline.txt
