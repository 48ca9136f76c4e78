"""Small all-formats workload for compute-sanitizer (not a test; run under gpurun):
  compute-sanitizer --tool memcheck  python tests/sanitize_run.py
  compute-sanitizer --tool racecheck python tests/sanitize_run.py
Every kernel of the path runs at least once on real hardware (TMA bulk copies, mbarriers, the cross-batch offset
chain, long units with carried bucket heads, the folded checksum, the decoder) and every result is compared with
the oracle, so a sanitizer finding and a wrong byte are both visible."""
import gzip
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gzp_b200  # noqa: E402
import oracle  # noqa: E402
from gzp_b200 import BGZF, GZIP, MGZIP, SNAP, ZLIB, synth  # noqa: E402


def main():
    text = synth.corpus_stream(600_000, 4321)
    cases = [(BGZF, 6, 65280, text[:200_000]), (BGZF, 1, 65280, text[:140_000]), (BGZF, 9, 65280, text[:70_000]), (BGZF, 0, 65280, text[:70_000]),
             (MGZIP, 6, 131072, text[:300_000]), (GZIP, 6, 262144, synth.fastq(600_000)), (ZLIB, 4, 40000, text[:130_000]),
             (SNAP, 0, 131072, synth.low_entropy(200_000) + text[:100_000]), (BGZF, 6, 65280, b""), (BGZF, 6, 65280, bytes(70_000))]
    if "quick" in sys.argv:      # three cases: the lazy pipeline, long units (carried bucket heads), lazy2 — when GPU minutes are short
        cases = [(BGZF, 6, 65280, text[:140_000]), (MGZIP, 6, 131072, text[:200_000]), (BGZF, 9, 65280, text[:70_000])]
    for fmt, level, bs, data in cases:
        ctx = gzp_b200.Context(fmt, level, max_block_bytes=bs, max_blocks_in_flight=2)     # several batches: lanes wrap, offsets chain
        got = ctx.encode_stream(data, bs)
        ctx.close()
        assert got == oracle.compress_stream(fmt, level, bs, [data]), (fmt, level, len(data))
        print("ok fmt %d level %d: %d -> %d" % (fmt, level, len(data), len(got)), flush=True)
    if "quick" in sys.argv:
        return
    comp = oracle.compress_stream(BGZF, 6, 65280, [text[:300_000]])
    dec = gzp_b200.Decoder(BGZF, max_blocks_in_flight=3)
    out, used = dec.decode(comp)
    dec.close()
    assert out == text[:300_000] and used == len(comp) and gzip.decompress(comp) == out
    print("ok decode", flush=True)


if __name__ == "__main__":
    main()
