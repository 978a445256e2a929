#!/bin/bash
mkdir -p gpurun_out
for W in sa_micro chamfer stress; do
  ( time python bench.py --workload $W ) > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "$W rc=$?"
  tail -3 gpurun_out/bench_$W.err | grep real
  head -c 1500 gpurun_out/bench_$W.json; echo
done
python bench.py --impl reference --device cuda --steps 3 --warmup 1 2>/dev/null | head -c 600; echo
