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
from gzp_b200 import BGZF, MGZIP

pytestmark = pytest.mark.gpu


def _edge_blocks():
    rnd = random.Random(7)
    blocks = [b"", b"a", b"ab" * 2, bytes(31), bytes(32), b"x" * 33, bytes(65280), b"\xff" * 65280]
    blocks += [bytes(rnd.randrange(255) for _ in range(n)) for n in (5, 100, 511, 512, 513, 4999, 5000, 10001, 40000, 65280)]
    blocks += [bytes(rnd.choice(b"abcd") for _ in range(n)) for n in (1000, 20000, 65280)]
    blocks += [bytes(rnd.choice(b"ab") for _ in range(300)) * 200]
    blocks += [(b"0123456789abcdef" * 5000)[:65280]]
    return blocks


@pytest.mark.parametrize("level", [6, 2, 3, 4, 5, 7, 0])
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
