#!/bin/bash
# Round-2 session z (not a test): ONE call — A/B of the kernel variants built by build_r2z_variants.sh, automatic choice,
# then the evidence for the chosen build: GPU suite, bench lines, ncu launch list, ncu --set full with source.
mkdir -p gpurun_out
T0=$SECONDS
el() { echo $((SECONDS - T0)); }
J=gpurun_out/r2z_kernels.jsonl; : > $J; : > gpurun_out/r2z_kernels.err
run() { label=$1; lib=$2; n=$3; lvl=${4:-6}; GZPB_LIB=$PWD/gzp_b200/libgzpb_$lib.so GZPB_PERF_INFLIGHT=$n timeout 150 python tests/perf_kernels.py $n $lvl 5 $label >> $J 2>> gpurun_out/r2z_kernels.err; echo "[$(el)s] $label rc=$?"; }
run head@3256 head 3256
run A0m1@3256 A0m1 3256
run B0m1@3256 B0m1 3256
run D0m1@3256 D0m1 3256
run A0m1@4736 A0m1 4736
run A2m1@4736 A2m1 4736
run A2m1@3256 A2m1 3256
run head@3256 head 3256
cut -c1-330 $J
CH=$(python profiles/sessions/r2z_pick.py $J); LIB=${CH% *}; N=${CH#* }
echo "choice: $LIB inflight $N" | tee gpurun_out/r2z_choice.txt
# the chosen combination as one build: level 6 at its batch size, level 9 (lazy2) and level 1 against the reference build
run final@$N $LIB $N
run L9_head@3256 head 3256 9
run L9_final@3256 $LIB 3256 9
run L1_head@3256 head 3256 1
run L1_final@3256 $LIB 3256 1
tail -n 5 $J | cut -c1-330
cp gzp_b200/libgzpb_$LIB.so gzp_b200/libgzpb.so
( time timeout 400 python -m pytest tests -m gpu -q --tb=short -x ) > gpurun_out/r2z_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest_gpu.log
tail -4 gpurun_out/r2z_pytest_gpu.log; echo "[$(el)s] pytest done"
B=$((N * 10))
timeout 300 python bench.py --inflight $N --blocks $B > gpurun_out/r2z_bench_bgzf.json 2> gpurun_out/r2z_bench_bgzf.err; echo "rc=$?" >> gpurun_out/r2z_bench_bgzf.err
echo "[$(el)s] bench done"; cut -c1-600 gpurun_out/r2z_bench_bgzf.json
if [ $(el) -lt 420 ]; then
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2z_launches.csv \
    python bench.py --steps 2 --warmup 1 --inflight $N --blocks $((N * 2)) --cpu-sample-mb 8 > gpurun_out/r2z_ncu_bench.log 2>&1
echo "[$(el)s] launch list done"; fi
if [ $(el) -lt 470 ]; then
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_match|k_split|k_link" -c 5 -o gpurun_out/r2z_full -f \
    python tests/prof_run.py $N > gpurun_out/r2z_ncu_full.log 2>&1
echo "[$(el)s] ncu full done"; fi
for c in mgzip gzip9 snap; do
  if [ $(el) -lt 520 ]; then
    timeout 120 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2z_bench_$c.json 2> gpurun_out/r2z_bench_$c.err; echo "rc=$?" >> gpurun_out/r2z_bench_$c.err
    echo "[$(el)s] bench $c done"
  fi
done
for c in bgzf mgzip gzip9 snap; do [ -s gpurun_out/r2z_bench_$c.json ] && python -c "
import json
d=json.load(open('gpurun_out/r2z_bench_$c.json')); print('$c', 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'kms', {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_launch'].items()})"; done
echo "[$(el)s] end"
