"""Small fixed workload for ncu captures (not a test): 592 BGZF blocks, level 6."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gzp_b200
from gzp_b200 import synth, BGZF
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 592
data = synth.text(65280 * nblk)
ctx = gzp_b200.Context(BGZF, 6, max_blocks_in_flight=nblk)
ctx.set_profiling(True)
for it in range(3):
    t0 = time.time(); out = ctx.encode_stream(data); dt = time.time() - t0
    print("iter", it, len(data), "->", len(out), "%.1f ms  %.2f GB/s" % (dt * 1e3, len(data) / dt / 1e9))
for k in ("chain", "match", "emit", "gather"):
    ms, cnt = ctx.kernel_ms(k)
    print(k, "%.3f ms/launch" % (ms / max(cnt, 1)), cnt)
