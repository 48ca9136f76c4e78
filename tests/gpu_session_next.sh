#!/bin/bash
# First GPU session of the next round (not a test): everything needed to decide about the sparse path in one call.
#   gpurun --timeout 900 -- 'bash tests/gpu_session_next.sh'
# Outputs under gpurun_out/: pytest log, bench line, variant A/B at levels 6 and 4, ncu launch list of the bench,
# ncu --set full captures (with source) of the default and the sparse kernel sets.  Summaries afterwards, here:
#   python profiles/summarize.py launches gpurun_out/launches.csv profiles/rN_launch_shares.txt "<cmd>"
#   python profiles/summarize.py full   gpurun_out/full_default.ncu-rep profiles/rN_ncu_full_default.txt "<cmd>"
#   python profiles/summarize.py source gpurun_out/full_default.ncu-rep profiles/rN_source_default.txt "<cmd>" 40
#   python profiles/summarize.py source gpurun_out/full_sparse.ncu-rep  profiles/rN_source_sparse.txt  "<cmd>" 40
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 400 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?" >> gpurun_out/bench.err
timeout 120 python tests/perf_variants.py 3256 6 5 > gpurun_out/variants_l6.jsonl 2> gpurun_out/variants_l6.err
timeout 120 python tests/perf_variants.py 3256 4 5 > gpurun_out/variants_l4.jsonl 2> gpurun_out/variants_l4.err
timeout 60 python tests/perf_writer.py > gpurun_out/perf_writer.json 2> gpurun_out/perf_writer.err
GZPB_BENCH_NO_VARIANTS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --blocks 3256 --cpu-sample-mb 8 > gpurun_out/ncu_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_match|k_split|k_link" -c 4 -o gpurun_out/full_default -f \
    python tests/prof_run.py 2368 > gpurun_out/ncu_full_default.log 2>&1
GZPB_SPARSE=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_smatch" -c 2 -o gpurun_out/full_sparse -f \
    python tests/prof_run.py 2368 > gpurun_out/ncu_full_sparse.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; head -c 900 gpurun_out/bench.json; echo; cat gpurun_out/variants_l6.jsonl | cut -c1-400; tail -2 gpurun_out/variants_l6.err
