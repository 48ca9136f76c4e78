#!/bin/bash
# Round-2 session c (not a test): baseline (HEAD kernels) vs rewritten k_match/k_link vs the table-fed token parse.
mkdir -p gpurun_out
: > gpurun_out/r2c_kernels.jsonl
run() { label=$1; shift; env "$@" timeout 150 python tests/perf_kernels.py 3256 ${LEVEL:-6} 5 "$label" >> gpurun_out/r2c_kernels.jsonl 2>> gpurun_out/r2c_kernels.err; }
run base GZPB_LIB=$PWD/gzp_b200/libgzpb_base.so
run new GZPB_X=0
run tparse128 GZPB_SPARSE=3
run tparse256 GZPB_SPARSE=3 GZPB_SPARSE_CHUNK=256
run tparse512 GZPB_SPARSE=3 GZPB_SPARSE_CHUNK=512
LEVEL=4 run L4new GZPB_X=0
LEVEL=4 run L4tparse GZPB_SPARSE=3
LEVEL=9 run L9new GZPB_X=0
LEVEL=9 run L9tparse GZPB_SPARSE=3
GZPB_SPARSE=3 timeout 300 python -m pytest tests -m gpu -q --tb=short -x -k "parity or fuzz or fullsize" > gpurun_out/pytest_sparse3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sparse3.log
GZPB_SPARSE=3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_smatch|k_emit" -c 2 -o gpurun_out/full_tparse -f \
    python tests/prof_run.py 3256 > gpurun_out/ncu_full_tparse.log 2>&1
tail -4 gpurun_out/pytest_sparse3.log; cat gpurun_out/r2c_kernels.jsonl | cut -c1-420; tail -3 gpurun_out/r2c_kernels.err
