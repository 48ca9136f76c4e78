#!/bin/bash
# Round-2 session o (not a test): lazy2 levels through the windowed parse — GPU suite, per-kernel times, gzip9 config.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/r2o_kernels.jsonl
for lv in 6 8 9 5; do timeout 150 python tests/perf_kernels.py 3256 $lv 5 "L$lv" >> gpurun_out/r2o_kernels.jsonl 2>> gpurun_out/r2o_kernels.err; done
timeout 400 python bench.py --config gzip9 --steps 4 --warmup 3 --cpu-sample-mb 8 > gpurun_out/r2o_bench_gzip9.json 2> gpurun_out/r2o_bench_gzip9.err; echo "rc=$?" >> gpurun_out/r2o_bench_gzip9.err
tail -4 gpurun_out/pytest_gpu.log
cat gpurun_out/r2o_kernels.jsonl | cut -c1-400
python -c "
import json
d=json.load(open('gpurun_out/r2o_bench_gzip9.json')); print('gzip9', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"
tail -n 2 gpurun_out/r2o_*.err
