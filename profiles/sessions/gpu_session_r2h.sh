#!/bin/bash
# Round-2 session h (not a test): full GPU suite on the refactored chain build + folded checksum, all four bench configs.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
: > gpurun_out/r2h_kernels.jsonl
run() { label=$1; shift; env "$@" timeout 150 python tests/perf_kernels.py 3256 ${LEVEL:-6} 5 "$label" >> gpurun_out/r2h_kernels.jsonl 2>> gpurun_out/r2h_kernels.err; }
run folded GZPB_X=0
run separate_check GZPB_SEPARATE_CHECK=1
timeout 300 python bench.py > gpurun_out/r2h_bench_bgzf.json 2> gpurun_out/r2h_bench_bgzf.err; echo "rc=$?" >> gpurun_out/r2h_bench_bgzf.err
for c in mgzip snap gzip9; do
  timeout 300 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2h_bench_$c.json 2> gpurun_out/r2h_bench_$c.err; echo "rc=$?" >> gpurun_out/r2h_bench_$c.err
done
GZPB_FULL_COPIES=3000 timeout 200 python bench.py --full-stream --gpus 1 --feed reserve > gpurun_out/r2h_fullstream_n1.json 2> gpurun_out/r2h_fullstream_n1.err; echo "rc=$?" >> gpurun_out/r2h_fullstream_n1.err
tail -6 gpurun_out/pytest_gpu.log
cat gpurun_out/r2h_kernels.jsonl | cut -c1-400
for c in bgzf mgzip snap gzip9; do python -c "
import json
d=json.load(open('gpurun_out/r2h_bench_$c.json')); print('$c', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'cpu', round(d['cpu_baseline']['value'],3), d['cpu_baseline']['cores'], 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"; done
cat gpurun_out/r2h_fullstream_n1.json | cut -c1-300
for f in gpurun_out/r2h_*.err; do tail -n 2 $f; done
