"""Kernel LOGIC under the CPU SIMT emulator (tests/emu) vs the oracle — no GPU needed.

The same .cu sources that nvcc builds for sm_100a are compiled with g++ against
tests/emu/shim (threads = fibers, warp collectives / __syncthreads / mbarrier =
rendezvous points) and driven through the same C ABI.  This checks the kernels'
arithmetic and control flow bit-for-bit in the container that has no GPU; the
`-m gpu` tests remain the parity tests proper (real hardware, real memory model).
Test infrastructure only: nothing in gzp_b200/ can load the emulated library.
"""
import gzip
import random
import zlib

import pytest

import emu
import oracle
from gzp_b200 import synth

TEXT = synth.text(300000)


def _run(fmt, level, bs, data):
    ctx = emu.EmuContext(fmt, level, max_block_bytes=bs)
    try:
        got = ctx.encode_stream(data, bs)
    finally:
        ctx.close()
    want = oracle.compress_stream(fmt, level, bs or oracle.DEFAULT_BUFSIZE[fmt], [data])
    assert got == want, "emulated kernels differ from the oracle (fmt %d level %d, %d bytes)" % (fmt, level, len(data))
    return got


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
def test_emu_bgzf_levels(level):
    got = _run(oracle.BGZF, level, 0, TEXT[:140000])
    assert gzip.decompress(got) == TEXT[:140000]


def test_emu_bgzf_edges():
    rnd = random.Random(5)
    cases = [b"", b"x", TEXT[:31], TEXT[:32], TEXT[:33], bytes(70000), b"\xff" * 40000,
             bytes(rnd.getrandbits(8) for _ in range(50000)), b"abcdefghij" * 7000, TEXT[:4999], TEXT[:10001],
             TEXT[:65280], TEXT[:65281]]
    for d in cases:
        assert gzip.decompress(_run(oracle.BGZF, 6, 0, d)) == d


def test_emu_mgzip_long_units():
    assert gzip.decompress(_run(oracle.MGZIP, 6, 131072, TEXT)) == TEXT


def test_emu_dictionary_formats():
    assert gzip.decompress(_run(oracle.GZIP, 6, 131072, TEXT)) == TEXT
    d = synth.fastq(100000) + TEXT[:200000]
    assert gzip.decompress(_run(oracle.GZIP, 9, 262144, d)) == d
    assert zlib.decompress(_run(oracle.ZLIB, 4, 40000, TEXT[:130000])) == TEXT[:130000]
    raw = _run(oracle.RAWDEFLATE, 6, 32768, TEXT[:100000])
    assert zlib.decompressobj(-15).decompress(raw) == TEXT[:100000]


def test_emu_level1_fastest_all_formats():
    """Level 1 = deflate_compress_fastest (ht_matchfinder): 15-bit buckets, depth-2 chains, 65535-byte /
    8192-match DEFLATE blocks — long units, dictionary carry and the passthrough threshold (51 bytes)."""
    for d in (TEXT[:51], TEXT[:52], synth.fastq(120000), bytes(70000)):
        assert gzip.decompress(_run(oracle.BGZF, 1, 0, d)) == d
    assert gzip.decompress(_run(oracle.MGZIP, 1, 131072, TEXT)) == TEXT
    assert gzip.decompress(_run(oracle.GZIP, 1, 131072, TEXT)) == TEXT
    assert zlib.decompress(_run(oracle.ZLIB, 1, 40000, TEXT[:130000])) == TEXT[:130000]


def test_emu_snap():
    rnd = random.Random(7)
    for d in (TEXT, bytes(70000), bytes(rnd.getrandbits(8) for _ in range(50000)), b"", b"abc"):
        _run(oracle.SNAP, 0, 131072, d)


# ---- decode path (k_inflate) -------------------------------------------------------------
import io
import struct


def _bgzf_member(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    """A BGZF member whose payload comes from stock zlib (not from this repo's encoder)."""
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    raw = co.compress(data) + co.flush()
    hdr = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0]) + struct.pack("<H", len(raw) + 26 - 1)
    return hdr + raw + struct.pack("<II", zlib.crc32(data), len(data))


@pytest.fixture()
def emu_backend(monkeypatch):
    """Route the Python mirror (gzp_b200.api) to the emulated library for the duration of a test."""
    from gzp_b200 import _lib as product
    monkeypatch.setattr(product, "_lib", emu.lib())
    yield


def test_emu_inflate_roundtrip_own_streams():
    dec = emu.EmuDecoder(oracle.BGZF)
    try:
        for lvl in (0, 2, 6, 9):
            s = oracle.compress_stream(oracle.BGZF, lvl, 65280, [TEXT])
            rc, out, _, _ = dec.decode_stream(s)
            assert rc == 0 and out == TEXT
            assert oracle.decode_stream(oracle.BGZF, s)[:2] == (0, TEXT)
    finally:
        dec.close()
    dm = emu.EmuDecoder(oracle.MGZIP)
    try:
        s = oracle.compress_stream(oracle.MGZIP, 6, 131072, [TEXT])
        rc, out, _, _ = dm.decode_stream(s)
        assert rc == 0 and out == TEXT
    finally:
        dm.close()


def test_emu_inflate_stock_zlib_members():
    rnd = random.Random(3)
    blobs = [TEXT[:60000], b"", b"a", bytes(60000), bytes(rnd.getrandbits(8) for _ in range(30000)), TEXT[1000:1200],
             b"ab" * 20000, bytes(rnd.choice(b"ACGT") for _ in range(50000))]
    dec = emu.EmuDecoder(oracle.BGZF)
    try:
        for lvl, strat in ((1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                           (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_RLE)):
            s = b"".join(_bgzf_member(b, lvl, strat) for b in blobs)
            rc, out, _, _ = dec.decode_stream(s)
            assert rc == 0 and out == b"".join(blobs), (lvl, strat)
    finally:
        dec.close()


def test_emu_inflate_error_paths_match_the_oracle():
    good = oracle.compress_stream(oracle.BGZF, 6, 65280, [TEXT[:150000]])
    dec = emu.EmuDecoder(oracle.BGZF)
    try:
        cases = {}
        b = bytearray(good); b[-28 - 8] ^= 0xFF; cases["crc"] = bytes(b)             # CRC of the last data block
        b = bytearray(good); b[3] = 0; cases["flag"] = bytes(b)                       # FEXTRA flag cleared
        b = bytearray(good); b[12] = ord("X"); cases["sid"] = bytes(b)                # bad SID
        cases["truncated"] = good[:len(good) - 40]
        b = bytearray(good); b[40] ^= 0x55; b[41] ^= 0xAA; b[60] ^= 0x0F; cases["payload"] = bytes(b)
        cases["short_tail"] = good + b"\x1f\x8b\x08"                                  # < HEADER_SIZE trailing bytes = EOF
        for name, s in cases.items():
            rc_o, out_o, f_o, e_o = oracle.decode_stream(oracle.BGZF, s)
            rc, out, f, e = dec.decode_stream(s)
            if name == "payload":
                # corrupt DEFLATE data: either a decode error or a checksum mismatch, never success
                assert rc in (-12, -13) and rc_o in (-12, -13), name
                continue
            assert rc == rc_o, (name, rc, rc_o)
            if rc == 0:
                assert out == out_o == TEXT[:150000]
            if rc == -12:
                assert (f, e) == (f_o, e_o)
    finally:
        dec.close()


def test_emu_pardecompress_mirror(emu_backend):
    import gzp_b200

    class Dribble(io.RawIOBase):
        """A reader that returns odd-sized pieces, so members straddle read() calls."""

        def __init__(self, data):
            self.d, self.p, self.k = data, 0, 0

        def read(self, n=-1):
            self.k += 1
            step = min(n if n > 0 else 1 << 30, 7919 * (1 + self.k % 5))
            b = self.d[self.p:self.p + step]
            self.p += len(b)
            return b

    sink = io.BytesIO()
    w = gzp_b200.ParCompressBuilder(gzp_b200.Bgzf).compression_level(6).from_writer(sink)
    w.write(TEXT[:100000]); w.flush(); w.write(TEXT[100000:]); w.finish()
    comp = sink.getvalue()
    assert gzip.decompress(comp) == TEXT
    r = gzp_b200.ParDecompressBuilder(gzp_b200.Bgzf).from_reader(Dribble(comp))
    r._CHUNK = 50000
    got = bytearray()
    while True:
        b = r.read(33333)
        if not b:
            break
        got.extend(b)
    r.finish()
    assert bytes(got) == TEXT
    # InvalidCheck carries found / expected (lib.rs:139-140)
    bad = bytearray(comp); bad[-28 - 8] ^= 1
    r = gzp_b200.ParDecompress.builder(gzp_b200.Bgzf).from_reader(io.BytesIO(bytes(bad)))
    with pytest.raises(gzp_b200.GzpError) as ei:
        r.read()
    assert ei.value.variant == "InvalidCheck" and ei.value.found != ei.value.expected
    with pytest.raises(gzp_b200.GzpError):
        gzp_b200.ParDecompressBuilder(gzp_b200.Bgzf).num_threads(0)
    with pytest.raises(gzp_b200.GzpError):
        gzp_b200.ParDecompressBuilder(gzp_b200.Bgzf).buffer_size(100)


def test_emu_syncz_block_writers_keep_the_reference_quirks(emu_backend):
    """BgzfSyncWriter / MgzipSyncWriter behind SyncZ / ZBuilder(num_threads <= 1): one block per write() call
    at most, BGZF_EOF after every block flushed, nothing for an empty flush."""
    import gzp_b200
    data = TEXT[:200000]

    def model(fmt, writes, blocksize, level):
        """Pure-Python restatement on top of the oracle's per-block encoder."""
        out, buf = bytearray(), bytearray()
        for w in writes:
            buf.extend(w)
            if len(buf) >= blocksize:
                out += oracle.encode_block(fmt, level, bytes(buf[:blocksize]), None, False)
                del buf[:blocksize]
        if fmt == oracle.BGZF:
            while buf:
                k = min(len(buf), 65280)
                out += oracle.encode_block(fmt, level, bytes(buf[:k]), None, False) + gzp_b200.BGZF_EOF
                del buf[:k]
        elif buf:
            out += oracle.encode_block(fmt, level, bytes(buf), None, False)
        return bytes(out)

    writes = [data[i:i + 50000] for i in range(0, len(data), 50000)]
    for fmt, F, bs in ((oracle.BGZF, gzp_b200.Bgzf, 65280), (oracle.MGZIP, gzp_b200.Mgzip, 131072)):
        sink = io.BytesIO()
        z = gzp_b200.ZBuilder(F).num_threads(1).compression_level(4).from_writer(sink)
        assert isinstance(z, gzp_b200.SyncZ)
        for w in writes:
            z.write(w)
        z.finish()
        assert sink.getvalue() == model(fmt, writes, bs, 4)
        assert gzip.decompress(sink.getvalue()) == data
    # an empty BGZF sync stream is empty: no EOF marker at all
    sink = io.BytesIO()
    gzp_b200.SyncZBuilder(gzp_b200.Bgzf).from_writer(sink).finish()
    assert sink.getvalue() == b""
    # more than one thread -> ParCompress (lib.rs:246)
    assert isinstance(gzp_b200.ZBuilder(gzp_b200.Bgzf).num_threads(4).from_writer(io.BytesIO()), gzp_b200.ParCompress)
