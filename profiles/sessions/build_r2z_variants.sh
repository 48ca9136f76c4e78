#!/bin/bash
# Round-2 session z, CPU side (not a test): one library per kernel-variant combination for the A/B on the GPU box.
#   link : A = group tags written into the bucket heads (default)   B = A + hash4 / hash3 jobs as two launches
#          D = tag byte per bucket, one round trip per tile (GZPB_LINK_PIPE) + two launches
#   emit : 0 = 4-tile ring, own scratch (22 units per SM)   2 = GZPB_EMIT_DIET, 3-tile ring, scratch in the ring (32 per SM)
#   match: m1 = order / prev3 loads one position ahead (default)   m0 = loads at the top of the position
set -e
cd "$(dirname "$0")/../.."
python gzp_b200/build.py >/dev/null          # the other objects
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -diag-suppress 1886"
O=gzp_b200/csrc
mkdir -p /tmp/r2z
for L in A B D; do for E in 0 2; do for M in m1 m0; do
  defs=""
  [ $L = B ] && defs="$defs -DGZPB_LINK_SPLIT=1"
  [ $L = D ] && defs="$defs -DGZPB_LINK_PIPE=1 -DGZPB_LINK_SPLIT=1"
  [ $E = 2 ] && defs="$defs -DGZPB_EMIT_DIET=1 -DGZPB_EMIT_RING=3 -DGZPB_EMIT_MINCTAS=32"
  [ $M = m0 ] && defs="$defs -DGZPB_MATCH_PIPE=0"
  ( $NV $defs -x cu -c $O/deflate_kernels.cu -o /tmp/r2z/dk_$L$E$M.o 2>&1 | grep -v deprecated || true
    nvcc -shared -o gzp_b200/libgzpb_$L$E$M.so /tmp/r2z/dk_$L$E$M.o $O/gzpb_api.o $O/gzpb_decode_api.o $O/inflate_kernels.o $O/snappy_kernels.o -lpthread 2>&1 | grep -v deprecated || true ) &
done; done; wait; done
ls -la gzp_b200/libgzpb_*.so
