"""Ad-hoc GPU debugging helper (not a test): python tests/gpu_debug.py"""
import sys, os, zlib, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, gzp_b200
from gzp_b200 import synth, BGZF

def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    t = synth.text(65280 * 8)
    ctx = gzp_b200.Context(BGZF, level, max_blocks_in_flight=16)
    blocks = [t[i:i + 65280] for i in range(0, len(t), 65280)] + [b"", b"a" * 40, bytes(1000), t[:3000], t[:12345]]
    got = ctx.encode_blocks([(b, None, False) for b in blocks])
    nbad = 0
    for i, (b, (enc, s, a)) in enumerate(zip(blocks, got)):
        want = oracle.encode_block(oracle.BGZF, level, b)
        ok = enc == want
        dec_ok = None
        try:
            dec_ok = zlib.decompress(enc[18:-8], -15) == b
        except Exception as e:
            dec_ok = repr(e)
        if not ok:
            nbad += 1
            k = next((j for j in range(min(len(enc), len(want))) if enc[j] != want[j]), None)
            print(f"block {i} len {len(b)}: MISMATCH got {len(enc)} want {len(want)} first diff at {k}; decodes={dec_ok}")
            print("  got ", enc[:40].hex()); print("  want", want[:40].hex())
        else:
            print(f"block {i} len {len(b)}: ok ({len(enc)} bytes) decodes={dec_ok}")
    print("bad:", nbad)
    ctx.set_profiling(True)
    big = synth.text(65280 * 16)
    t0 = time.time(); out = ctx.encode_stream(big); dt = time.time() - t0
    print("stream", len(big), "->", len(out), "%.3fs" % dt, "match oracle:", out == oracle.compress_stream(oracle.BGZF, level, 65280, [big]))
    for k in ("chain", "match", "emit", "gather"):
        print(k, ctx.kernel_ms(k))

main()
