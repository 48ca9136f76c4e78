#!/bin/bash
# Round-2 session zz (not a test): the default build vs the same + full-tile fast paths in k_split / k_link; the winner becomes
# libgzpb.so and is run through the GPU suite, the bench configs (Mgzip at two batch sizes) and the ncu captures.
mkdir -p gpurun_out
T0=$SECONDS
el() { echo $((SECONDS - T0)); }
J=gpurun_out/r2zz_kernels.jsonl; : > $J; : > gpurun_out/r2zz_kernels.err
run() { label=$1; lib=$2; n=$3; lvl=${4:-6}; GZPB_LIB=$PWD/gzp_b200/libgzpb_$lib.so GZPB_PERF_INFLIGHT=$n timeout 150 python tests/perf_kernels.py $n $lvl 5 $label >> $J 2>> gpurun_out/r2zz_kernels.err; echo "[$(el)s] $label rc=$?"; }
run F0w0@4736 F0w0 4736
run F1w0@4736 F1w0 4736
run G0w0@4736 G0w0 4736
run G1w0@4736 G1w0 4736
run F0w1@4736 F0w1 4736
run F0w0@4736 F0w0 4736
cut -c1-330 $J
LIB=$(python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/r2zz_kernels.jsonl") if l.startswith("{")]
REF = "c7708e1ea7667a54"      # packed stream of 4736 blocks, level 6: session r2z (builds A0m1 = head-equivalent output, and B2m1)
ok = {}
for d in rows:
    if d["status_ok"] and d["sha1"] == REF:
        k = d["label"].split("@")[0]
        if k not in ok or d["ms_best"] < ok[k]["ms_best"]:
            ok[k] = d
chain = [(ok[k + "w0"]["kernel_ms"]["chain"], k) for k in ("F0", "F1", "G0", "G1") if k + "w0" in ok]
C = min(chain)[1] if chain else "F0"
W = "w1" if "F0w1" in ok and "F0w0" in ok and ok["F0w1"]["kernel_ms"]["emit"] < 0.995 * ok["F0w0"]["kernel_ms"]["emit"] else "w0"
print(C + W)
PY
)
echo "choice: $LIB" | tee gpurun_out/r2zz_choice.txt
run final@4736 $LIB 4736
run L9_final@3256 $LIB 3256 9
run L1_final@3256 $LIB 3256 1
python - <<'PY' | tee -a gpurun_out/r2zz_choice.txt
import json
rows = [json.loads(l) for l in open("gpurun_out/r2zz_kernels.jsonl") if l.startswith("{")]
want = {"final@4736": "c7708e1ea7667a54", "L9_final@3256": "21afd9f56935f0bc", "L1_final@3256": "927b2a78f1b4ed8f"}   # session r2z, reference build
for d in rows:
    if d["label"] in want:
        print(d["label"], "identical to the reference build" if d["status_ok"] and d["sha1"] == want[d["label"]] else "DIFFERS", d["ms_best"], d["kernel_ms"])
PY
cp gzp_b200/libgzpb_$LIB.so gzp_b200/libgzpb.so
( time timeout 300 python -m pytest tests -m gpu -q --tb=short -x ) > gpurun_out/r2zz_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2zz_pytest_gpu.log
tail -4 gpurun_out/r2zz_pytest_gpu.log; echo "[$(el)s] pytest done"
timeout 200 python bench.py > gpurun_out/r2zz_bench_bgzf.json 2> gpurun_out/r2zz_bench_bgzf.err; echo "rc=$?" >> gpurun_out/r2zz_bench_bgzf.err
echo "[$(el)s] bench done"; cut -c1-300 gpurun_out/r2zz_bench_bgzf.json
for n in 2368 3256; do if [ $(el) -lt 230 ]; then
  timeout 150 python bench.py --config mgzip --inflight $n --blocks $((n * 5)) --steps 5 --warmup 3 --cpu-sample-mb 8 > gpurun_out/r2zz_bench_mgzip_$n.json 2> gpurun_out/r2zz_bench_mgzip_$n.err; echo "rc=$?" >> gpurun_out/r2zz_bench_mgzip_$n.err
  echo "[$(el)s] bench mgzip $n done"; fi; done
if [ $(el) -lt 260 ]; then
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2zz_launches.csv \
    python bench.py --steps 2 --warmup 1 --blocks 9472 --cpu-sample-mb 8 > gpurun_out/r2zz_ncu_bench.log 2>&1
echo "[$(el)s] launch list done"; fi
if [ $(el) -lt 290 ]; then
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_match|k_split|k_link|k_gather" -c 6 -o gpurun_out/r2zz_full -f \
    python tests/prof_run.py 4736 > gpurun_out/r2zz_ncu_full.log 2>&1
echo "[$(el)s] ncu full done"; fi
if [ $(el) -lt 330 ]; then
timeout 100 python bench.py --config gzip9 --steps 5 --warmup 3 --cpu-sample-mb 8 > gpurun_out/r2zz_bench_gzip9.json 2> gpurun_out/r2zz_bench_gzip9.err; echo "rc=$?" >> gpurun_out/r2zz_bench_gzip9.err
echo "[$(el)s] bench gzip9 done"; fi
if [ $(el) -lt 370 ]; then
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2zz_bench_reference.json 2> gpurun_out/r2zz_bench_reference.err
echo "[$(el)s] reference arm done"; fi
for c in bgzf mgzip_2368 mgzip_3256 gzip9; do [ -s gpurun_out/r2zz_bench_$c.json ] && python -c "
import json
d=json.load(open('gpurun_out/r2zz_bench_$c.json')); print('$c', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'launches', d.get('gpu_launches'), 'checked', d['e2e'].get('decoded_input_bytes_checked'), 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"; done
tail -n 3 gpurun_out/r2zz_bench_*.err | cut -c1-200
echo "[$(el)s] end"
