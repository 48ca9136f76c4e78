#!/bin/bash
# Round-2 session w (not a test): the shipped build — GPU suite, the four bench configs, ncu launch list, ncu --set full with source.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2w_bench_bgzf.json 2> gpurun_out/r2w_bench_bgzf.err; echo "rc=$?" >> gpurun_out/r2w_bench_bgzf.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err
for c in mgzip snap gzip9; do
  timeout 400 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2w_bench_$c.json 2> gpurun_out/r2w_bench_$c.err; echo "rc=$?" >> gpurun_out/r2w_bench_$c.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2w_launches.csv \
    python bench.py --steps 2 --warmup 1 --blocks 6512 --cpu-sample-mb 8 > gpurun_out/r2w_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_match|k_split|k_link|k_gather" -c 5 -o gpurun_out/r2w_full -f \
    python tests/prof_run.py 3256 > gpurun_out/r2w_ncu_full.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"k_snap" -c 1 -o gpurun_out/r2w_full_snap -f \
    python tests/perf_formats.py low-entropy > gpurun_out/r2w_ncu_full_snap.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
for c in bgzf mgzip snap gzip9; do python -c "
import json
d=json.load(open('gpurun_out/r2w_bench_$c.json')); print('$c', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'cpu', round(d['cpu_baseline']['value'],3), d['cpu_baseline']['cores'], 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"; done
cat gpurun_out/r2w_bench_reference.json | cut -c1-300
for f in gpurun_out/r2w_*.err; do tail -n 1 $f; done
