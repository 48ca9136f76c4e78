#!/bin/bash
# Round-2 session p (not a test): k_emit templated on lazy2 — per-kernel times at levels 6 / 4 / 9, parity subset.
mkdir -p gpurun_out
: > gpurun_out/r2p_kernels.jsonl
for lv in 6 4 9; do timeout 150 python tests/perf_kernels.py 3256 $lv 5 "L$lv" >> gpurun_out/r2p_kernels.jsonl 2>> gpurun_out/r2p_kernels.err; done
( time timeout 300 python -m pytest tests -m gpu -q --tb=short -k "parity or fuzz or round2" ) > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
cat gpurun_out/r2p_kernels.jsonl | cut -c1-400; tail -4 gpurun_out/r2p_pytest.log
