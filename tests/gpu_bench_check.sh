#!/bin/bash
# GPU session (not a test): bench.py both arms, exactly as the round-end driver runs them.
mkdir -p gpurun_out
export GZPB_BENCH_WATCHDOG=60
timeout 100 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?" >> gpurun_out/bench.err
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "rc=$?" >> gpurun_out/bench_ref.err
tail -n 25 gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
