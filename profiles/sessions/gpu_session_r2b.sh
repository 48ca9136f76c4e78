#!/bin/bash
# Round-2 session b (not a test): parity of the rewritten k_match / k_link + A/B of the parked-candidate batch size.
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests -m gpu -q --tb=short -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/r2b_kernels.jsonl
for b in 1 4 8 12 16 20 32; do
  GZPB_MATCH_BATCH=$b timeout 120 python tests/perf_kernels.py 3256 6 5 "batch$b" >> gpurun_out/r2b_kernels.jsonl 2>> gpurun_out/r2b_kernels.err
done
for lv in 4 8; do
  for b in 1 12; do
    GZPB_MATCH_BATCH=$b timeout 120 python tests/perf_kernels.py 3256 $lv 3 "L$lv batch$b" >> gpurun_out/r2b_kernels.jsonl 2>> gpurun_out/r2b_kernels.err
  done
done
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/r2b_kernels.jsonl | cut -c1-420
