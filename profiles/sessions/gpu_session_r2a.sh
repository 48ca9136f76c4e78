#!/bin/bash
# Round-2 first GPU session (not a test): evidence before code.
#   gpurun --timeout 1100 -- 'bash tests/gpu_session_r2a.sh'
mkdir -p gpurun_out
( nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv; nproc; lscpu | head -30; free -g; nvidia-smi topo -m ) > gpurun_out/r2a_host.txt 2>&1
( find / -xdev \( -name 'libdeflate*' -o -name 'bgzip*' -o -name 'pigz' -o -name 'libsnappy*' -o -name 'libz-ng*' -o -name 'cargo' \) 2>/dev/null | head -40 ) > gpurun_out/r2a_probe.txt 2>&1
( time timeout 400 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
GZPB_BENCH_NO_VARIANTS=1 timeout 200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "rc=$?" >> gpurun_out/bench.err
timeout 60 python tests/perf_writer.py > gpurun_out/perf_writer.json 2> gpurun_out/perf_writer.err
timeout 100 python tests/perf_formats.py > gpurun_out/perf_formats.txt 2>&1
GZPB_BENCH_NO_VARIANTS=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --blocks 3256 --cpu-sample-mb 8 > gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_match|k_split|k_link|k_check" -c 5 -o gpurun_out/full_default -f \
    python tests/prof_run.py 3256 > gpurun_out/ncu_full_default.log 2>&1
GZPB_SPARSE=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_smatch" -c 1 -o gpurun_out/full_sparse -f \
    python tests/prof_run.py 3256 > gpurun_out/ncu_full_sparse.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; head -c 900 gpurun_out/bench.json; echo; cat gpurun_out/r2a_probe.txt; ls -la gpurun_out
