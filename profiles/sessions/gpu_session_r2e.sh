#!/bin/bash
# Round-2 session e (not a test): k_match inner-loop variants (compile-time, one library each) + the small-grid gather.
mkdir -p gpurun_out
: > gpurun_out/r2e_kernels.jsonl
run() { label=$1; shift; env "$@" timeout 150 python tests/perf_kernels.py 3256 ${LEVEL:-6} 5 "$label" >> gpurun_out/r2e_kernels.jsonl 2>> gpurun_out/r2e_kernels.err; }
run km0 GZPB_X=0
run km1 GZPB_LIB=$PWD/gzp_b200/libgzpb_km1.so
run km2 GZPB_LIB=$PWD/gzp_b200/libgzpb_km2.so
run km3 GZPB_LIB=$PWD/gzp_b200/libgzpb_km3.so
LEVEL=9 run L9km0 GZPB_X=0
LEVEL=9 run L9km3 GZPB_LIB=$PWD/gzp_b200/libgzpb_km3.so
for g in 148 296 592 100000; do
  GZPB_GATHER_CTAS=$g timeout 200 python bench.py --steps 6 --warmup 3 --cpu-sample-mb 8 > gpurun_out/r2e_bench_g$g.json 2> gpurun_out/r2e_bench_g$g.err
done
cat gpurun_out/r2e_kernels.jsonl | cut -c1-400
for g in 148 296 592 100000; do python -c "
import json,sys
d=json.load(open('gpurun_out/r2e_bench_g$g.json')); print('gather ctas $g', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'writer', d['e2e']['incremental_writer'] and round(d['e2e']['incremental_writer']['value_per_gpu'],3))"; done
