#!/bin/bash
# Round-2 session i (not a test): the rewritten k_snap on hardware, gzip9 batch sizes, where the chain build's time goes.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -k "snap or parity or fuzz or device_ex or multi" ) > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
timeout 200 python tests/perf_formats.py Snap > gpurun_out/r2i_perf_formats_snap.txt 2>&1
timeout 300 python bench.py --config snap --steps 5 --warmup 3 > gpurun_out/r2i_bench_snap.json 2> gpurun_out/r2i_bench_snap.err; echo "rc=$?" >> gpurun_out/r2i_bench_snap.err
for f in 1184 1776; do
  timeout 300 python bench.py --config gzip9 --steps 4 --warmup 3 --inflight $f --blocks $((f*2)) --cpu-sample-mb 8 > gpurun_out/r2i_bench_gzip9_$f.json 2> gpurun_out/r2i_bench_gzip9_$f.err; echo "rc=$?" >> gpurun_out/r2i_bench_gzip9_$f.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2i_launches_prof.csv python tests/prof_run.py 3256 > gpurun_out/r2i_ncu_prof.log 2>&1
tail -4 gpurun_out/r2i_pytest.log
cat gpurun_out/r2i_perf_formats_snap.txt
for c in snap gzip9_1184 gzip9_1776; do python -c "
import json
d=json.load(open('gpurun_out/r2i_bench_$c.json')); print('$c', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"; done
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2i_launches_prof.csv')) if len(r)>10 and r[0].isdigit()]
import collections
t=collections.defaultdict(list)
for r in rows: t[r[4].split('(')[0]].append(float(r[-1]))
for k,v in t.items(): print(k, len(v), [round(x/1e6,3) for x in v[-3:]])
PY
