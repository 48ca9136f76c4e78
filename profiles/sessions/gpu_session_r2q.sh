#!/bin/bash
# Round-2 session q (not a test): 16+4 lists (14 link jobs per SM) vs 32+8 lists (28 per SM).
mkdir -p gpurun_out
: > gpurun_out/r2q_kernels.jsonl
run() { label=$1; shift; env "$@" timeout 150 python tests/perf_kernels.py 3256 ${LEVEL:-6} 5 "$label" >> gpurun_out/r2q_kernels.jsonl 2>> gpurun_out/r2q_kernels.err; }
run lists_16_4 GZPB_LIB=$PWD/gzp_b200/libgzpb_lists0.so
run lists_32_8 GZPB_X=0
run lists_16_4 GZPB_LIB=$PWD/gzp_b200/libgzpb_lists0.so
run lists_32_8 GZPB_X=0
timeout 300 python bench.py --config mgzip --steps 5 --warmup 3 --cpu-sample-mb 8 > gpurun_out/r2q_bench_mgzip.json 2> gpurun_out/r2q_bench_mgzip.err
cat gpurun_out/r2q_kernels.jsonl | cut -c1-400
python -c "
import json
d=json.load(open('gpurun_out/r2q_bench_mgzip.json')); print('mgzip', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"
