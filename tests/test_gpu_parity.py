"""GPU parity tests proper: the CUDA path, called through the C ABI
(libgzpb.so), must be BIT-EXACT with the CPU oracle on the same inputs, and the
streams must decode with stock decoders (the reference's own test strategy,
/root/reference/src/deflate.rs:679-1379)."""
import gzip
import random
import zlib

import pytest

import oracle
import gzp_b200
from gzp_b200 import BGZF, GZIP, MGZIP, RAWDEFLATE, SNAP, ZLIB

pytestmark = pytest.mark.gpu


def _edge_blocks():
    rnd = random.Random(7)
    blocks = [b"", b"a", b"ab" * 2, bytes(31), bytes(32), b"x" * 33, bytes(65280), b"\xff" * 65280]
    blocks += [bytes(rnd.randrange(255) for _ in range(n)) for n in (5, 100, 511, 512, 513, 4999, 5000, 10001, 40000, 65280)]
    blocks += [bytes(rnd.choice(b"abcd") for _ in range(n)) for n in (1000, 20000, 65280)]
    blocks += [bytes(rnd.choice(b"ab") for _ in range(300)) * 200]
    blocks += [(b"0123456789abcdef" * 5000)[:65280]]
    return blocks


@pytest.mark.parametrize("level", [6, 2, 3, 4, 5, 7, 0, 1])
def test_bgzf_blocks_bit_exact(text_corpus, level):
    ctx = gzp_b200.Context(BGZF, level, max_blocks_in_flight=64)
    blocks = [text_corpus[i:i + 65280] for i in range(0, 20 * 65280, 65280)] if level == 6 else \
             [text_corpus[i:i + 65280] for i in range(0, 4 * 65280, 65280)]
    blocks += _edge_blocks()
    msgs = [(b, None, i == len(blocks) - 1) for i, b in enumerate(blocks)]
    got = ctx.encode_blocks(msgs)
    for i, ((b, _, last), (enc, s, a)) in enumerate(zip(msgs, got)):
        want = oracle.encode_block(oracle.BGZF, level, b, None, last)
        assert enc == want, f"block {i} (len {len(b)}) differs from the oracle at level {level}"
    stream = b"".join(e for e, _, _ in got)
    assert gzip.decompress(stream) == b"".join(blocks)
    ctx.close()


def test_bgzf_stream_matches_oracle_and_gzip(text_corpus):
    ctx = gzp_b200.Context(BGZF, 6, max_blocks_in_flight=8)   # forces several batches / lanes
    data = text_corpus[:65280 * 21 + 1234]
    got = ctx.encode_stream(data)
    want = oracle.compress_stream(oracle.BGZF, 6, 65280, [data])
    assert got == want
    assert gzip.decompress(got) == data
    assert got.endswith(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    # exactly buffer_size bytes: one held-back block flushed by finish()
    d2 = text_corpus[:65280]
    assert ctx.encode_stream(d2) == oracle.compress_stream(oracle.BGZF, 6, 65280, [d2])
    assert ctx.encode_stream(b"") == oracle.compress_stream(oracle.BGZF, 6, 65280, [b""])
    ctx.close()


def test_mgzip_small_blocks(text_corpus):
    ctx = gzp_b200.Context(MGZIP, 6, max_block_bytes=65536, max_blocks_in_flight=16)
    blocks = [text_corpus[i:i + 65536] for i in range(0, 6 * 65536, 65536)] + [b"", b"abc"]
    got = ctx.encode_blocks([(b, None, False) for b in blocks])
    for b, (enc, _, _) in zip(blocks, got):
        assert enc == oracle.encode_block(oracle.MGZIP, 6, b)
    assert gzip.decompress(b"".join(e for e, _, _ in got)) == b"".join(blocks)
    ctx.close()


def test_mgzip_default_block_size_long_units(text_corpus):
    # 131 072-byte blocks (mgzip default, BASELINE configs[2]) -> segmented sub-units with a 32 KiB halo
    ctx = gzp_b200.Context(MGZIP, 6, max_blocks_in_flight=8)
    data = text_corpus[: 131072 * 5 + 4321]
    got = ctx.encode_stream(data)
    want = oracle.compress_stream(oracle.MGZIP, 6, 131072, [data])
    assert got == want
    assert gzip.decompress(got) == data
    rnd = random.Random(3)
    blocks = [bytes(131072), bytes(rnd.randrange(255) for _ in range(100000)), text_corpus[7:7 + 70000], b"q" * 65537]
    res = ctx.encode_blocks([(b, None, False) for b in blocks])
    for b, (enc, _, _) in zip(blocks, res):
        assert enc == oracle.encode_block(oracle.MGZIP, 6, b)
    ctx.close()


@pytest.mark.parametrize("fmt,bs,level", [(GZIP, 131072, 6), (GZIP, 40000, 3), (GZIP, 262144, 7), (RAWDEFLATE, 131072, 6),
                                          (ZLIB, 131072, 6), (ZLIB, 32768, 2)])
def test_dictionary_formats_bit_exact_and_decodable(text_corpus, fmt, bs, level):
    # Gzip / Zlib / RawDeflate: 32 KiB dictionary carry + sync-flush terminators + combined check
    ctx = gzp_b200.Context(fmt, level, max_block_bytes=bs, max_blocks_in_flight=4)
    data = text_corpus[: bs * 3 + 12345]
    got = ctx.encode_stream(data, bs)
    want = oracle.compress_stream(fmt, level, bs, [data])
    assert got == want
    if fmt == GZIP:
        assert gzip.decompress(got) == data
    elif fmt == ZLIB:
        assert zlib.decompress(got) == data
    else:
        assert zlib.decompressobj(-15).decompress(got) == data
    # per-block API with explicit dictionaries and checks
    msgs = oracle.chunk_stream(fmt, bs, [data[: bs + 999]])
    res = ctx.encode_blocks(msgs)
    for (b, d, last), (enc, s, a) in zip(msgs, res):
        assert enc == oracle.encode_block(fmt, level, b, d, last)
        if fmt == GZIP:
            assert (s, a) == (zlib.crc32(b), len(b))
        if fmt == ZLIB:
            assert (s, a) == (zlib.adler32(b), len(b))
    # empty stream and tiny stream
    for tiny in (b"", b"x", text_corpus[:300]):
        assert ctx.encode_stream(tiny, bs) == oracle.compress_stream(fmt, level, bs, [tiny])
    ctx.close()


def test_reference_regression_vector_on_gpu():
    # /root/reference/src/deflate.rs:949-992: 206 bytes, buffer_size = DICT_SIZE, Gzip level 3 (the default)
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    data = bytes(g["regression_input"])
    ctx = gzp_b200.Context(GZIP, 3, max_block_bytes=32768, max_blocks_in_flight=4)
    got = ctx.encode_stream(data, 32768)
    assert got == oracle.compress_stream(oracle.GZIP, 3, 32768, [data])
    assert gzip.decompress(got) == data
    ctx.close()


def _snappy_deframe(z):
    import pyarrow as pa
    codec = pa.Codec("snappy")
    out, pos = b"", 0
    while pos < len(z):
        if z[pos] == 0xFF:
            assert z[pos:pos + 10] == bytes.fromhex("ff060000734e61507059"); pos += 10; continue
        t = z[pos]; ln = int.from_bytes(z[pos + 1:pos + 4], "little"); crc = int.from_bytes(z[pos + 4:pos + 8], "little")
        body = z[pos + 8:pos + 4 + ln]; pos += 4 + ln
        if t == 0:
            v = sh = i = 0
            while True:
                b = body[i]; v |= (b & 0x7F) << sh; sh += 7; i += 1
                if b < 0x80:
                    break
            dec = codec.decompress(body, decompressed_size=v).to_pybytes()
        else:
            dec = body
        assert oracle.lib().oracle_crc32c_masked(dec, len(dec)) == crc
        out += dec
    return out


def test_snap_blocks_bit_exact_and_decodable(text_corpus):
    from gzp_b200 import synth
    rnd = random.Random(17)
    ctx = gzp_b200.Context(SNAP, 0, max_blocks_in_flight=8)
    blocks = [text_corpus[:131072], synth.low_entropy(131072), bytes(rnd.randrange(256) for _ in range(70000)), b"", b"a", b"ab" * 8,
              b"x" * 17, text_corpus[100:100 + 65536], text_corpus[5:5 + 65537], bytes(100000), synth.fastq(90000),
              bytes(rnd.choice(b"ab") for _ in range(3000)) * 20]
    res = ctx.encode_blocks([(b, None, False) for b in blocks])
    for i, (b, (enc, _, _)) in enumerate(zip(blocks, res)):
        assert enc == oracle.encode_block(oracle.SNAP, 0, b), f"snap block {i} (len {len(b)}) differs from the oracle"
        assert _snappy_deframe(enc) == b
    data = synth.low_entropy(131072 * 9 + 777)
    got = ctx.encode_stream(data)
    assert got == oracle.compress_stream(oracle.SNAP, 0, 131072, [data])
    assert _snappy_deframe(got) == data
    ctx.close()


@pytest.mark.parametrize("level", [8, 9])
def test_lazy2_levels(text_corpus, level):
    # levels 8-9: lazy2 parser, depth 300/600 (BASELINE configs[4] uses level 9 for Gzip)
    ctx = gzp_b200.Context(BGZF, level, max_blocks_in_flight=8)
    blocks = [text_corpus[i:i + 65280] for i in range(0, 3 * 65280, 65280)] + _edge_blocks()[:12]
    got = ctx.encode_blocks([(b, None, False) for b in blocks])
    for i, (b, (enc, _, _)) in enumerate(zip(blocks, got)):
        assert enc == oracle.encode_block(oracle.BGZF, level, b), f"block {i} level {level}"
    ctx.close()
    ctx = gzp_b200.Context(GZIP, level, max_block_bytes=262144, max_blocks_in_flight=2)
    from gzp_b200 import synth
    data = synth.fastq(262144 * 2 + 999)
    got = ctx.encode_stream(data, 262144)
    assert got == oracle.compress_stream(oracle.GZIP, level, 262144, [data])
    assert gzip.decompress(got) == data
    ctx.close()


def test_bgzf_block_size_exceeded_error():
    # bgzf.rs:218-223: a compressed payload >= 65536 bytes is an error, not a silent fallback.
    # ParCompressBuilder does not bound buffer_size for Bgzf, so 131072 random bytes hit it.
    rnd = random.Random(5)
    data = bytes(rnd.randrange(256) for _ in range(131072))
    with pytest.raises(ValueError) as eo:
        oracle.encode_block(oracle.BGZF, 6, data)
    assert eo.value.args[0] == -3
    ctx = gzp_b200.Context(BGZF, 6, max_block_bytes=131072, max_blocks_in_flight=2)
    with pytest.raises(gzp_b200.GzpError) as e:
        ctx.encode_blocks([(data, None, False)])
    assert e.value.variant == "BlockSizeExceeded" and e.value.code == -3
    # the context stays usable after a per-block error
    ok = ctx.encode_blocks([(b"hello " * 1000, None, True)])[0][0]
    assert ok == oracle.encode_block(oracle.BGZF, 6, b"hello " * 1000, None, True)
    ctx.close()


def test_units_longer_than_512k_have_the_right_checksum(text_corpus):
    """Regression (found on the emulator): k_check's x^(8*512*j) table spans 512 KiB; a 600 000-byte Mgzip / Gzip block
    needs the second table for its CRC-32."""
    import gzip
    for fmt, level, bs in ((MGZIP, 4, 600000), (GZIP, 6, 560000)):
        data = text_corpus[:bs + 7000]
        ctx = gzp_b200.Context(fmt, level, max_block_bytes=bs, max_blocks_in_flight=2)
        got = ctx.encode_stream(data, bs)
        ctx.close()
        assert got == oracle.compress_stream(fmt, level, bs, [data])
        assert gzip.decompress(got) == data
