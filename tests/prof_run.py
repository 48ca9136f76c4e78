"""Small fixed workload for ncu captures (not a test): 592 BGZF blocks, level 6."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gzp_b200
from gzp_b200 import synth, BGZF
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 592
data = synth.text_stream(65280 * nblk)
ctx = gzp_b200.Context(BGZF, 6, max_blocks_in_flight=nblk)
ctx.set_profiling(True)
for it in range(3):
    t0 = time.time(); out = ctx.encode_stream(data); dt = time.time() - t0
    print("iter", it, len(data), "->", len(out), "%.1f ms  %.2f GB/s" % (dt * 1e3, len(data) / dt / 1e9))
for k in ("chain", "match", "emit", "gather"):
    ms, cnt = ctx.kernel_ms(k)
    print(k, "%.3f ms/launch" % (ms / max(cnt, 1)), cnt)

import ctypes as C
from gzp_b200 import _lib
L = _lib.load()
L.gzpb_debug_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
arr = (C.c_uint64 * 32)()
L.gzpb_debug_phase_cycles(ctx._h, arr, 1)
names = {0: "emit:blockstart", 1: "emit:parse", 2: "emit:sync", 3: "emit:huff_litlen", 4: "emit:huff_off", 5: "emit:precode_items",
         6: "emit:huff_pre", 7: "emit:costs", 8: "emit:header", 9: "emit:pack", 10: "emit:other", 16: "chain:warp0", 17: "chain:crc", 18: "chain:total"}
units = nblk * 3
for k, nm in names.items():
    print("%-20s %10.0f cycles/unit" % (nm, arr[k] / units))
