#!/bin/bash
# compute-sanitizer over the OA-Mix view parity test and the OA-Loss tests (GPU box); writes gpurun_out/san_*.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 python -m pytest -x -q -m gpu \
    "tests/test_gpu_oamix.py::test_view_matches_oracle" tests/test_gpu_oaloss.py -k "not nccl and not two_gpus" \
    > gpurun_out/san_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/san_$tool.txt | tail -1)"
  grep -E "passed|failed" gpurun_out/san_$tool.txt | tail -1
done
