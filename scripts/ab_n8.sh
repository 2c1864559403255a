#!/bin/bash
# N = 8: peer-store exchange vs NCCL all-gathers (value leg only, 20 steps)
for rep in 1 2; do
for ex in peer nccl; do
  OADG_EXCHANGE=$ex timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e 2>gpurun_out/ab8.err | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$ex', round(d['value']), round(d['ms_per_step'],4))" || tail -3 gpurun_out/ab8.err
done
done
