#!/bin/bash
# N = 2: group sizes of the loader loop (20 steps, value only)
run() {
  name=$1; shift
  for rep in 1 2; do
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e 2>gpurun_out/abg.err | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('$name', round(d['value'],1), round(d['ms_per_step'],4))" || tail -5 gpurun_out/abg.err
  done
}
run default X=1
run first2 OADG_FIRST_GROUP=2
run first4 OADG_FIRST_GROUP=4
run gmax2 OADG_GROUP_BATCHES=2
run gmax3 OADG_GROUP_BATCHES=3
run gmax5 OADG_GROUP_BATCHES=5
