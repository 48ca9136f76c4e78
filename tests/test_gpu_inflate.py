"""GPU parity of the block DECODE path (k_inflate behind gzpb_decode_stream / ParDecompress).

Bit-exactness here means: decoded bytes == the original input, block CRCs verified on
the device, statuses identical to the oracle's restatement of the reference's reader /
worker loops (/root/reference/src/par/decompress.rs:163-207).  All calls go through the
C ABI (include/gzpb.h)."""
import gzip
import io
import random
import struct
import zlib

import pytest

import gzp_b200
import oracle
from gzp_b200 import synth

pytestmark = pytest.mark.gpu


def _bgzf_member(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    raw = co.compress(data) + co.flush()
    hdr = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0]) + struct.pack("<H", len(raw) + 26 - 1)
    return hdr + raw + struct.pack("<II", zlib.crc32(data), len(data))


def test_roundtrip_bgzf_all_levels():
    data = synth.text(65280 * 40 + 12345)
    dec = gzp_b200.Decoder(gzp_b200.BGZF)
    for lvl in (0, 2, 4, 6, 9):
        ctx = gzp_b200.Context(gzp_b200.BGZF, lvl, max_blocks_in_flight=64)
        comp = ctx.encode_stream(data)
        ctx.close()
        out, used = dec.decode(comp)
        assert out == data and used == len(comp)
    assert dec.launch_count() > 0
    dec.close()


def test_roundtrip_mgzip_and_oracle_agreement():
    data = synth.text(131072 * 9 + 77)
    ctx = gzp_b200.Context(gzp_b200.MGZIP, 6, max_block_bytes=131072, max_blocks_in_flight=16)
    comp = ctx.encode_stream(data, 131072)
    ctx.close()
    dec = gzp_b200.Decoder(gzp_b200.MGZIP)
    out, _ = dec.decode(comp)
    dec.close()
    assert out == data
    assert oracle.decode_stream(oracle.MGZIP, comp)[:2] == (0, data)


def test_stock_zlib_members_every_block_type():
    rnd = random.Random(3)
    blobs = [synth.text(60000), b"", b"a", bytes(60000), bytes(rnd.getrandbits(8) for _ in range(30000)), synth.text(1200)[1000:],
             b"ab" * 20000, bytes(rnd.choice(b"ACGT") for _ in range(50000)), synth.fastq(60000)]
    dec = gzp_b200.Decoder(gzp_b200.BGZF)
    for lvl, strat in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                       (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_RLE)):
        s = b"".join(_bgzf_member(b, lvl, strat) for b in blobs)
        out, _ = dec.decode(s)
        assert out == b"".join(blobs), (lvl, strat)
    dec.close()


def test_error_paths_match_the_oracle():
    text = synth.text(150000)
    ctx = gzp_b200.Context(gzp_b200.BGZF, 6, max_blocks_in_flight=8)
    good = ctx.encode_stream(text)
    ctx.close()
    dec = gzp_b200.Decoder(gzp_b200.BGZF)
    cases = {}
    b = bytearray(good); b[-28 - 8] ^= 0xFF; cases["crc"] = bytes(b)
    b = bytearray(good); b[3] = 0; cases["flag"] = bytes(b)
    b = bytearray(good); b[12] = ord("X"); cases["sid"] = bytes(b)
    cases["truncated"] = good[:len(good) - 40]
    cases["short_tail"] = good + b"\x1f\x8b\x08"
    for name, s in cases.items():
        rc_o, out_o, f_o, e_o = oracle.decode_stream(oracle.BGZF, s)
        try:
            out, _ = dec.decode(s)
            rc = 0
        except gzp_b200.GzpError as e:
            rc = e.code
            if rc == -12:
                assert (e.found, e.expected) == (f_o, e_o)
        assert rc == rc_o, (name, rc, rc_o)
        if rc == 0:
            assert out == out_o == text
    b = bytearray(good); b[40] ^= 0x55; b[41] ^= 0xAA; b[60] ^= 0x0F
    with pytest.raises(gzp_b200.GzpError) as ei:
        dec.decode(bytes(b))
    assert ei.value.code in (-12, -13)
    dec.close()


def test_fuzzed_payloads_never_hang_or_overrun():
    """Random corruption of valid members: every outcome is ok-with-identical-bytes or a clean error."""
    text = synth.text(65280 * 3)
    ctx = gzp_b200.Context(gzp_b200.BGZF, 6, max_blocks_in_flight=8)
    good = ctx.encode_stream(text)
    ctx.close()
    dec = gzp_b200.Decoder(gzp_b200.BGZF)
    rnd = random.Random(11)
    for _ in range(40):
        b = bytearray(good)
        for _k in range(rnd.randint(1, 4)):
            b[rnd.randrange(18, len(b) - 36)] ^= 1 << rnd.randrange(8)
        try:
            out, _ = dec.decode(bytes(b))
            assert out == text
        except gzp_b200.GzpError as e:
            assert e.code in (-11, -12, -13, -14, -6)
    dec.close()


def test_pardecompress_mirror_reads_like_the_reference():
    data = synth.text(65280 * 25 + 999)
    sink = io.BytesIO()
    w = gzp_b200.ParCompressBuilder(gzp_b200.Bgzf).compression_level(6).from_writer(sink)
    for off in range(0, len(data), 100003):
        w.write(data[off:off + 100003])
    w.finish()
    comp = sink.getvalue()
    assert gzip.decompress(comp) == data
    r = gzp_b200.ParDecompressBuilder(gzp_b200.Bgzf).from_reader(io.BytesIO(comp))
    r._CHUNK = 300000
    got = bytearray()
    while True:
        piece = r.read(77777)
        if not piece:
            break
        got.extend(piece)
    r.finish()
    assert bytes(got) == data


def test_native_reader_object_on_gpu():
    """gzpb_reader_* (ParDecompress as a C object): 40 MB of text through a dribbling source, chunks of 4 MiB."""
    import io
    from gzp_b200 import synth
    data = synth.text_stream(40_000_000)
    ctx = gzp_b200.Context(gzp_b200.BGZF, 6, max_blocks_in_flight=256)
    comp = ctx.encode_stream(data)
    ctx.close()

    class Dribble(io.RawIOBase):
        def __init__(self, d):
            self.d, self.p, self.k = d, 0, 0

        def read(self, n=-1):
            self.k += 1
            step = min(n, 100003 * (1 + self.k % 7))
            b = self.d[self.p:self.p + step]
            self.p += len(b)
            return b

    r = gzp_b200.NativeParDecompress(gzp_b200.Bgzf(), Dribble(comp), chunk_bytes=4 << 20)
    got = r.read()
    r.finish()
    assert got == data
