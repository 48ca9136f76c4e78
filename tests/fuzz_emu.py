"""Long differential fuzz of the kernels under the CPU SIMT emulator against the oracle (not a test of the default suite:
`python tests/fuzz_emu.py [seconds] [seed]`; combine with GZPB_EMU_SCHED=reverse|random:N).  Random inputs — uniform bytes,
text, runs, low-entropy pages, short periodic patterns, mixtures — random lengths, buffer sizes (long units included),
levels 0-9 and all formats; every stream must equal the oracle's bit for bit.  Prints one line per 50 cases and a summary."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu  # noqa: E402
import oracle  # noqa: E402
from gzp_b200 import synth  # noqa: E402


def gen(rnd, n, text):
    kind = rnd.randrange(7)
    if kind == 0:
        return bytes(rnd.randrange(255) for _ in range(n))
    if kind == 1:
        off = rnd.randrange(0, max(1, len(text) - n - 1))
        return text[off:off + n]
    if kind == 2:
        return (bytes([rnd.randrange(256)]) * rnd.randrange(1, 700) + bytes(rnd.randrange(4) for _ in range(rnd.randrange(1, 50)))) * (n // 20 + 1)
    if kind == 3:
        return synth.low_entropy(n, seed=rnd.randrange(1 << 30))
    if kind == 4:
        pat = bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 300)))
        return (pat * (n // len(pat) + 1))[:n]
    if kind == 5:
        return synth.fastq(n)
    out = bytearray()
    while len(out) < n:
        out += gen(rnd, rnd.randrange(1, 4000), text)
    return bytes(out)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rnd = random.Random(seed)
    text = synth.corpus_stream(2_000_000, 99)
    t0 = time.time()
    cases = 0
    FMTS = [oracle.BGZF, oracle.BGZF, oracle.MGZIP, oracle.GZIP, oracle.ZLIB, oracle.RAWDEFLATE, oracle.SNAP]
    while time.time() - t0 < budget:
        fmt = rnd.choice(FMTS)
        level = rnd.choice([0, 1, 2, 3, 4, 5, 6, 6, 6, 7, 8, 9])
        if fmt == oracle.BGZF:
            bs = rnd.randrange(32768, 65281)
        else:
            bs = rnd.choice([rnd.randrange(32768, 70000), rnd.randrange(70000, 300000)])
        n = rnd.choice([rnd.randrange(0, 400), rnd.randrange(0, 200000), rnd.randrange(0, 500000)])
        data = gen(rnd, n, text)[:n]
        inflight = rnd.choice([1, 2, 3, 5])
        ctx = emu.EmuContext(fmt, level, max_block_bytes=bs, max_blocks_in_flight=inflight)
        try:
            got = ctx.encode_stream(data, bs)
        finally:
            ctx.close()
        want = oracle.compress_stream(fmt, level, bs, [data])
        if got != want:
            path = "/tmp/fuzz_emu_fail_%d_%d.bin" % (seed, cases)
            open(path, "wb").write(data)
            print("MISMATCH fmt %d level %d bs %d n %d inflight %d seed %d case %d input saved to %s" % (fmt, level, bs, n, inflight, seed, cases, path), flush=True)
            sys.exit(1)
        cases += 1
        if cases % 50 == 0:
            print("[%6.0fs] %d cases ok" % (time.time() - t0, cases), flush=True)
    print("fuzz_emu: %d cases identical to the oracle in %.0f s (seed %d, sched %s)" % (cases, time.time() - t0, seed, os.environ.get("GZPB_EMU_SCHED", "default")), flush=True)


if __name__ == "__main__":
    main()
