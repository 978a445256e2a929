#!/bin/bash
# One GPU-box visit: parity tests, default bench line, ncu launch list of one eager step.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; head -c 6000 gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_step.txt 2>&1
head -50 gpurun_out/launches_step.txt
