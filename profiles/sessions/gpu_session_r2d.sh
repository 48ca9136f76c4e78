#!/bin/bash
# Round-2 session d (not a test): whole GPU suite, the new bench.py (all four configs), the full-stream writer leg.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r2d_bench_bgzf.json 2> gpurun_out/r2d_bench_bgzf.err; echo "rc=$?" >> gpurun_out/r2d_bench_bgzf.err
for c in mgzip snap gzip9; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2d_bench_$c.json 2> gpurun_out/r2d_bench_$c.err; echo "rc=$?" >> gpurun_out/r2d_bench_$c.err
done
GZPB_FULL_COPIES=2000 timeout 200 python bench.py --full-stream --gpus 1 > gpurun_out/r2d_fullstream_n1.json 2> gpurun_out/r2d_fullstream_n1.err; echo "rc=$?" >> gpurun_out/r2d_fullstream_n1.err
tail -6 gpurun_out/pytest_gpu.log
for f in gpurun_out/r2d_bench_*.json gpurun_out/r2d_fullstream_n1.json; do echo "== $f"; head -c 1500 $f; echo; done
tail -3 gpurun_out/r2d_*.err
