#!/bin/bash
# Round-2 session g (not a test), 2 GPUs: the > 2 GiB stream compare of bench.py, the in-place producer feed, k_match variants.
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 2 --warmup 1 --blocks 65120 --cpu-sample-mb 8 \
    > gpurun_out/r2g_bench_big_n2.json 2> gpurun_out/r2g_bench_big_n2.err; echo "rc=$?" >> gpurun_out/r2g_bench_big_n2.err
for feed in reserve write; do
GZPB_FULL_COPIES=4000 timeout 300 python bench.py --full-stream --gpus 2 --feed $feed > gpurun_out/r2g_fullstream_${feed}_n2.json 2> gpurun_out/r2g_fullstream_${feed}_n2.err; echo "rc=$?" >> gpurun_out/r2g_fullstream_${feed}_n2.err
done
: > gpurun_out/r2g_kernels.jsonl
run() { label=$1; shift; env "$@" timeout 150 python tests/perf_kernels.py 3256 ${LEVEL:-6} 5 "$label" >> gpurun_out/r2g_kernels.jsonl 2>> gpurun_out/r2g_kernels.err; }
run km0 GZPB_X=0
run km1_prefetch GZPB_LIB=$PWD/gzp_b200/libgzpb_km1.so
run km2_ext64 GZPB_LIB=$PWD/gzp_b200/libgzpb_km2.so
run km3_both GZPB_LIB=$PWD/gzp_b200/libgzpb_km3.so
for g in 37 74 148; do
  GZPB_GATHER_CTAS=$g timeout 200 python bench.py --steps 6 --warmup 3 --cpu-sample-mb 8 > gpurun_out/r2g_bench_g$g.json 2> gpurun_out/r2g_bench_g$g.err
done
head -c 1800 gpurun_out/r2g_bench_big_n2.json; echo; tail -n 4 gpurun_out/r2g_bench_big_n2.err
cat gpurun_out/r2g_fullstream_*_n2.json | cut -c1-500
cat gpurun_out/r2g_kernels.jsonl | cut -c1-400
for g in 37 74 148; do python -c "
import json,sys
d=json.load(open('gpurun_out/r2g_bench_g$g.json')); print('gather ctas $g', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'writer', d['e2e']['incremental_writer'] and round(d['e2e']['incremental_writer']['value_per_gpu'],3))"; done
