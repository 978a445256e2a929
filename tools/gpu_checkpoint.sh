#!/bin/bash
# Round checkpoint on the GPU box: parity tests, default bench line, ncu launch list + full-set capture of one eager step,
# CUPTI timeline of one graph replay, compute-sanitizer memcheck / racecheck of smoke().
TAG=${1:-r02}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','clocks')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('fp32_path',{}).get('ms_per_step'), d.get('reference_gpu',{}).get('ms_per_step'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_step.csv > gpurun_out/${TAG}_launches_train_step.txt 2>&1; head -12 gpurun_out/${TAG}_launches_train_step.txt
python tools/graph_timeline.py bf16 64 > gpurun_out/${TAG}_graph_timeline.txt 2> gpurun_out/graph_timeline.err; head -2 gpurun_out/${TAG}_graph_timeline.txt
bash tools/gpu_ncu_full.sh ${TAG}

timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_sanitizer_memcheck.txt 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_sanitizer_racecheck.txt 2>&1; tail -4 gpurun_out/${TAG}_sanitizer_racecheck.txt
rm -f gpurun_out/graph_trace.json
ls -la gpurun_out | head -40
