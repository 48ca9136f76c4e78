#!/bin/bash
mkdir -p gpurun_out
export GZPB_BENCH_WATCHDOG=25
timeout 100 python bench.py --steps 6 --warmup 3 --cpu-sample-mb 64 > gpurun_out/diag_bench.json 2> gpurun_out/diag_bench.err
echo "rc=$?" >> gpurun_out/diag_bench.err
timeout 60 python bench.py --steps 4 --warmup 3 --blocks 6512 --inflight 3256 --cpu-sample-mb 8 > gpurun_out/diag_bench3256.json 2> gpurun_out/diag_bench3256.err
echo "rc=$?" >> gpurun_out/diag_bench3256.err
tail -c 1500 gpurun_out/diag_bench.err; cat gpurun_out/diag_bench.json | head -c 600
