#!/bin/bash
# One GPU session of round 1 (not a test): parity tests, bench, decode throughput, ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 120 python tests/perf_inflate.py 4096 6 > gpurun_out/inflate.json 2> gpurun_out/inflate.err
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_inflate -c 1 -o gpurun_out/inflate_full -f python tests/perf_inflate.py 2048 6 > gpurun_out/ncu_inflate.log 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --blocks 2048 --cpu-sample-mb 8 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/inflate.json; tail -2 gpurun_out/inflate.err
