"""GPU analogue of the reference's ignored proptests (/root/reference/src/deflate.rs:1053-1379,
src/snap.rs:163-240): random inputs (uniform bytes 0..254, runs, text, mixtures), random lengths,
buffer sizes, levels and formats — every stream must equal the oracle's bit for bit and decode
with an independent decoder."""
import gzip
import random
import zlib

import pytest

import oracle
import gzp_b200
from gzp_b200 import BGZF, GZIP, MGZIP, RAWDEFLATE, SNAP, ZLIB, synth

pytestmark = pytest.mark.gpu


def _gen(rnd, n, text):
    kind = rnd.randrange(6)
    if kind == 0:
        return bytes(rnd.randrange(255) for _ in range(n))              # proptest: 0..u8::MAX (exclusive)
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
    out = bytearray()
    while len(out) < n:
        out += _gen(rnd, rnd.randrange(1, 4000), text)
    return bytes(out)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_streams_all_formats(text_corpus, seed):
    rnd = random.Random(seed)
    for _ in range(10):
        fmt = rnd.choice([BGZF, BGZF, MGZIP, GZIP, ZLIB, RAWDEFLATE, SNAP])
        level = rnd.choice([0, 2, 3, 4, 5, 6, 7])
        bs = rnd.randrange(32768, 65280) if fmt == BGZF else rnd.randrange(32768, 140000)
        n = rnd.randrange(0, 327680)
        data = _gen(rnd, n, text_corpus)[:n]
        ctx = gzp_b200.Context(fmt, level, max_block_bytes=bs, max_blocks_in_flight=rnd.choice([1, 2, 5]))
        got = ctx.encode_stream(data, bs)
        want = oracle.compress_stream(fmt, level, bs, [data])
        assert got == want, f"fmt {fmt} level {level} bs {bs} n {n} seed {seed}"
        if fmt in (BGZF, MGZIP, GZIP):
            assert gzip.decompress(got) == data
        elif fmt == ZLIB:
            assert zlib.decompress(got) == data
        elif fmt == RAWDEFLATE:
            assert zlib.decompressobj(-15).decompress(got) == data
        ctx.close()


def test_random_streams_level1(text_corpus):
    """Level 1 (ht_matchfinder / fastest parser) over random formats, buffer sizes and inputs."""
    rnd = random.Random(101)
    for _ in range(8):
        fmt = rnd.choice([BGZF, MGZIP, GZIP, ZLIB, RAWDEFLATE])
        bs = rnd.randrange(32768, 65280) if fmt == BGZF else rnd.randrange(32768, 300000)
        n = rnd.randrange(0, 400000)
        data = _gen(rnd, n, text_corpus)[:n]
        ctx = gzp_b200.Context(fmt, 1, max_block_bytes=bs, max_blocks_in_flight=rnd.choice([1, 3, 8]))
        got = ctx.encode_stream(data, bs)
        want = oracle.compress_stream(fmt, 1, bs, [data])
        assert got == want, f"fmt {fmt} bs {bs} n {n}"
        ctx.close()


def test_parcompress_mirror_on_gpu_random_writes(text_corpus):
    import io
    rnd = random.Random(77)
    for fmt, F in ((BGZF, gzp_b200.Bgzf), (GZIP, gzp_b200.Gzip), (SNAP, gzp_b200.Snap)):
        bs = 65280 if fmt == BGZF else 50000
        data = _gen(rnd, 400000, text_corpus)[:400000]
        writes, pos = [], 0
        while pos < len(data):
            k = rnd.randrange(1, 10000) if rnd.random() < 0.7 else rnd.randrange(1, 4 * bs)
            writes.append(data[pos:pos + k]); pos += k
        flushes = {3, 9}
        sink = io.BytesIO()
        pc = gzp_b200.ParCompressBuilder(F).buffer_size(bs).compression_level(6).blocks_in_flight(3).from_writer(sink)
        for i, w in enumerate(writes):
            pc.write(w)
            if i in flushes:
                pc.flush()
        pc.finish()
        assert sink.getvalue() == oracle.compress_stream(fmt, 6, bs, writes, flushes)


def test_c_writer_matches_oracle_random_writes(text_corpus):
    """gzpb_writer_* (the C++ ParCompress) driven with random write sizes and flushes."""
    import ctypes as C
    from gzp_b200 import _lib
    L = _lib.load()
    rnd = random.Random(123)
    for fmt, bs in ((BGZF, 65280), (GZIP, 40000), (MGZIP, 131072), (SNAP, 70000), (ZLIB, 32768)):
        data = _gen(rnd, 500000, text_corpus)[:500000]
        chunks = bytearray()

        @_lib.SINK_FN
        def sink(user, ptr, n):
            chunks.extend(C.string_at(ptr, n))
            return 0

        h = C.c_void_p()
        assert L.gzpb_writer_create(C.byref(h), 0, fmt, 6, bs, 3, C.cast(sink, C.c_void_p), None) == 0
        writes, pos = [], 0
        while pos < len(data):
            k = rnd.randrange(1, 10000) if rnd.random() < 0.7 else rnd.randrange(1, 4 * bs)
            writes.append(data[pos:pos + k]); pos += k
        flushes = {2, 7}
        for i, wdata in enumerate(writes):
            assert L.gzpb_writer_write(h, wdata, len(wdata)) == 0
            if i in flushes:
                assert L.gzpb_writer_flush(h) == 0
        assert L.gzpb_writer_finish(h) == 0
        assert L.gzpb_writer_write(h, b"x", 1) == -7          # write after finish: ChannelSend
        L.gzpb_writer_destroy(h)
        assert bytes(chunks) == oracle.compress_stream(fmt, 6, bs, writes, flushes), f"fmt {fmt}"
