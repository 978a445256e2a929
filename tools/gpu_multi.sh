#!/bin/bash
# Multi-GPU visit: DP smoke (parameters identical across ranks), bench at N GPUs (weak headline + strong + fp32 legs).
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_smoke.py > gpurun_out/dp_smoke_$N.txt 2>&1; echo "dp_smoke rc=$?"; grep -v Warning gpurun_out/dp_smoke_$N.txt | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', d['e2e']['ms_per_step'], 'strong', d.get('strong'), 'fp32', d.get('fp32_path',{}).get('ms_per_step'))"
tail -3 gpurun_out/bench_n$N.err
