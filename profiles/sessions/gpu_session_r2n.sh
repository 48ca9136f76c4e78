#!/bin/bash
# Round-2 session n (not a test): k_snap with batched probes (global input) — parity + throughput.
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -q --tb=short -k "snap or fuzz or device_ex or multi or compress_file" ) > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
timeout 200 python tests/perf_formats.py Snap > gpurun_out/r2n_perf_formats_snap.txt 2>&1
timeout 300 python bench.py --config snap --steps 5 --warmup 3 > gpurun_out/r2n_bench_snap.json 2> gpurun_out/r2n_bench_snap.err; echo "rc=$?" >> gpurun_out/r2n_bench_snap.err
tail -4 gpurun_out/r2n_pytest.log; cat gpurun_out/r2n_perf_formats_snap.txt
python -c "
import json
d=json.load(open('gpurun_out/r2n_bench_snap.json')); print('snap', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'kms', d['roofline']['kernel_ms_per_launch'])"
