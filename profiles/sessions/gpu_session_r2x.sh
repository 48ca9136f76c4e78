#!/bin/bash
# Round-2 session t (not a test): k_link: entry-only work (MATCH.ANY, shuffles) hoisted out of the tile-to-tile chain vs the previous build.
mkdir -p gpurun_out
: > gpurun_out/r2x_kernels.jsonl
run() { label=$1; shift; env "$@" timeout 150 python tests/perf_kernels.py 3256 ${LEVEL:-6} 5 "$label" >> gpurun_out/r2x_kernels.jsonl 2>> gpurun_out/r2x_kernels.err; }
run prev GZPB_LIB=$PWD/gzp_b200/libgzpb_prev.so
run new GZPB_X=0
run prev GZPB_LIB=$PWD/gzp_b200/libgzpb_prev.so
run new GZPB_X=0
LEVEL=9 run L9_prev GZPB_LIB=$PWD/gzp_b200/libgzpb_prev.so
LEVEL=9 run L9_new GZPB_X=0
cat gpurun_out/r2x_kernels.jsonl | cut -c1-400
