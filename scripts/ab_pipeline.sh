#!/bin/bash
# A/B of the loader loop's staging depth / launch ramp / consumer fence on one GPU (value and e2e, 20 steps)
for cfg in "2 1" "1 1" "2 0"; do
  set -- $cfg
  for rep in 1 2; do
  OADG_STAGE_AHEAD=$1 OADG_CONSUMER_FENCE=$2 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/abp.err | tail -1 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('ahead $1 fence $2: value', round(d['value']), 'e2e', round(d['e2e']['value']))"
  grep "e2e step wall" gpurun_out/abp.err | tail -1
  done
done
OADG_STAGE_AHEAD=1 timeout 300 python -m pytest tests/test_gpu_oamix.py -x -q -k "iter_batches or starved" 2>&1 | tail -2
