#!/bin/bash
# memcheck over the kernels added late in round 2: fp16-split forward, warp-per-row reduce, JSD, fused epilogue, two-view step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest -x -q -m gpu \
  tests/test_gpu_oaloss.py tests/test_gpu_two_view.py "tests/test_gpu_oamix.py::test_fused_normalize_pad_chw_output_is_exact" \
  -k "not two_gpus" > gpurun_out/san2_memcheck.txt 2>&1
echo "== memcheck: $(grep -E 'ERROR SUMMARY' gpurun_out/san2_memcheck.txt | tail -1)"
grep -E "passed|failed" gpurun_out/san2_memcheck.txt | tail -1
