#!/bin/bash
# Round-2 session zz, CPU side (not a test): one library per combination of
#   chain: F0 = default (16 + 4 lists, hash3 / hash4 link jobs as two launches)   F1 = F0 + GZPB_FULL_TILES (no activity predicates in full tiles)
#          G0 = 16 + 8 lists, every link job in 12 KiB, ONE launch (hash3 and hash4 jobs share the SMs)   G1 = G0 + GZPB_FULL_TILES
#   walk : w0 = default window walk (16 instructions per hop)   w1 = GZPB_EMIT_WALK2 (12 per hop; levels 1-7)
set -e
cd "$(dirname "$0")/../.."
python gzp_b200/build.py >/dev/null
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -diag-suppress 1886"
O=gzp_b200/csrc
mkdir -p /tmp/r2zz
rm -f gzp_b200/libgzpb_*.so
for C in F0 F1 G0 G1; do for W in w0 w1; do
  defs=""
  case $C in F1) defs="-DGZPB_FULL_TILES=1";; G0) defs="-DGZPB_LISTS=2 -DGZPB_LINK_SPLIT=0";; G1) defs="-DGZPB_LISTS=2 -DGZPB_LINK_SPLIT=0 -DGZPB_FULL_TILES=1";; esac
  [ $W = w1 ] && defs="$defs -DGZPB_EMIT_WALK2=1"
  ( $NV $defs -x cu -c $O/deflate_kernels.cu -o /tmp/r2zz/dk_$C$W.o 2>&1 | grep -v deprecated || true
    nvcc -shared -o gzp_b200/libgzpb_$C$W.so /tmp/r2zz/dk_$C$W.o $O/gzpb_api.o $O/gzpb_decode_api.o $O/inflate_kernels.o $O/snappy_kernels.o -lpthread 2>&1 | grep -v deprecated || true ) &
done; done; wait
ls -la gzp_b200/libgzpb*.so
