#!/bin/bash
# Round-2 session s (not a test): k_match candidate path with explicit word reuse vs three ld32u.
mkdir -p gpurun_out
: > gpurun_out/r2s_kernels.jsonl
run() { label=$1; shift; env "$@" timeout 150 python tests/perf_kernels.py 3256 ${LEVEL:-6} 5 "$label" >> gpurun_out/r2s_kernels.jsonl 2>> gpurun_out/r2s_kernels.err; }
run ld32u_x3 GZPB_LIB=$PWD/gzp_b200/libgzpb_km0.so
run word_reuse GZPB_X=0
run ld32u_x3 GZPB_LIB=$PWD/gzp_b200/libgzpb_km0.so
run word_reuse GZPB_X=0
LEVEL=9 run L9_ld32u_x3 GZPB_LIB=$PWD/gzp_b200/libgzpb_km0.so
LEVEL=9 run L9_word_reuse GZPB_X=0
cat gpurun_out/r2s_kernels.jsonl | cut -c1-400
