#!/bin/bash
N=${1:-2}
for V in NCCL_NOOP=1 NCCL_PROTO=Simple NCCL_ALGO=NVLS; do
  env $V python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 150 --warmup 5 --quick > gpurun_out/nccl_ab.json 2> gpurun_out/nccl_ab.err
  echo "$V rc=$? $(python -c "
import json
ls=[l for l in open('gpurun_out/nccl_ab.json') if l.startswith('{')]
d=json.loads(ls[-1]) if ls else {}
print(d.get('ms_per_step'), d.get('value'))")"
done
