#!/bin/bash
# Round-2 session l (not a test): k_emit walk + k_split ballot ranking measured; gzip9 batch 2368; GPU suite.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/r2l_kernels.jsonl
for lv in 6 4 9; do timeout 150 python tests/perf_kernels.py 3256 $lv 5 "L$lv" >> gpurun_out/r2l_kernels.jsonl 2>> gpurun_out/r2l_kernels.err; done
timeout 400 python bench.py --config gzip9 --steps 4 --warmup 3 --inflight 2368 --blocks 7104 --cpu-sample-mb 8 > gpurun_out/r2l_bench_gzip9_2368.json 2> gpurun_out/r2l_bench_gzip9_2368.err; echo "rc=$?" >> gpurun_out/r2l_bench_gzip9_2368.err
timeout 300 python bench.py --config mgzip --steps 5 --warmup 3 --cpu-sample-mb 8 > gpurun_out/r2l_bench_mgzip.json 2> gpurun_out/r2l_bench_mgzip.err
tail -4 gpurun_out/pytest_gpu.log
cat gpurun_out/r2l_kernels.jsonl | cut -c1-400
for c in gzip9_2368 mgzip; do python -c "
import json
d=json.load(open('gpurun_out/r2l_bench_$c.json')); print('$c', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"; done
tail -n 2 gpurun_out/r2l_*.err
