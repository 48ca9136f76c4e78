"""Kernel LOGIC under the CPU SIMT emulator (tests/emu) vs the oracle — no GPU needed.

The same .cu sources that nvcc builds for sm_100a are compiled with g++ against
tests/emu/shim (threads = fibers, warp collectives / __syncthreads / mbarrier =
rendezvous points) and driven through the same C ABI.  This checks the kernels'
arithmetic and control flow bit-for-bit in the container that has no GPU; the
`-m gpu` tests remain the parity tests proper (real hardware, real memory model).
Test infrastructure only: nothing in gzp_b200/ can load the emulated library.
"""
import ctypes as C
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
    """k_snap (chunk staged in shared memory, 32 probes per step, CRC-32C on the helper warps) against the oracle's
    sequential encoder: text, runs, random data (the accelerating skip crosses many probe batches), mixtures whose
    hits fall on every lane of a batch, sizes around the 17-byte / 15-byte margins and the 64 KiB chunk boundary."""
    rnd = random.Random(7)
    rand = bytes(rnd.getrandbits(8) for _ in range(50000))
    mix = b"".join(rand[i * 40:(i + 1) * 40] + TEXT[i * 7:i * 7 + rnd.randrange(1, 90)] + rand[i * 40:i * 40 + rnd.randrange(4, 40)] for i in range(900))
    for d in (TEXT, bytes(70000), rand, b"", b"abc", mix, synth.low_entropy(140000), synth.fastq(70000), bytes(rnd.choice(b"ab") for _ in range(3000)) * 25,
              TEXT[:65536], TEXT[:65537], TEXT[3:3 + 65535], rand[:4000] + bytes(5000) + rand[:4000], b"x" * 16, b"x" * 17, b"xy" * 9, TEXT[:31], TEXT[:32], TEXT[:47]):
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


def _drive_c_writer(L, fmt, level, bs, batch, data, rnd, devices=(0,), flushes=(2, 7), sink_fail_after=None):
    """gzpb_writer_* driven with random write sizes; returns (stream bytes, writes, final rc, stats)."""
    import ctypes as C
    from gzp_b200 import _lib
    chunks = bytearray()
    calls = [0]

    @_lib.SINK_FN
    def sink(user, ptr, n):
        calls[0] += 1
        if sink_fail_after is not None and calls[0] > sink_fail_after:
            return 1
        chunks.extend(C.string_at(ptr, n))
        return 0

    h = C.c_void_p()
    devs = (C.c_int * len(devices))(*devices)
    rc = L.gzpb_writer_create_multi(C.byref(h), devs, len(devices), fmt, level, bs, batch, C.cast(sink, C.c_void_p), None)
    assert rc == 0, rc
    writes, pos = [], 0
    while pos < len(data):
        k = rnd.randrange(1, 10000) if rnd.random() < 0.6 else rnd.randrange(1, 3 * bs)
        writes.append(data[pos:pos + k]); pos += k
    rc = 0
    for i, wdata in enumerate(writes):
        rc = L.gzpb_writer_write(h, wdata, len(wdata))
        if rc:
            break
        if i in flushes:
            rc = L.gzpb_writer_flush(h)
            if rc:
                break
    frc = L.gzpb_writer_finish(h)
    assert L.gzpb_writer_write(h, b"x", 1) == -7              # write after finish: ChannelSend
    st = [C.c_uint64(0) for _ in range(4)]
    L.gzpb_writer_stats(h, *[C.byref(x) for x in st])
    L.gzpb_writer_destroy(h)
    return bytes(chunks), writes, (rc or frc), [x.value for x in st]


def test_emu_pipelined_c_writer_all_formats():
    """The incremental C writer (pinned slabs, batches in flight on lazy streams, ordered sink) is byte-identical
    to the oracle's ParCompress model for random write sizes and flushes; batches of 2-3 blocks force slab
    switches, lane reuse and dictionary carry across slabs."""
    L = emu.lib()
    rnd = random.Random(2024)
    for fmt, bs, batch in ((oracle.BGZF, 65280, 2), (oracle.GZIP, 40000, 3), (oracle.MGZIP, 131072, 1),
                           (oracle.SNAP, 70000, 2), (oracle.ZLIB, 32768, 3)):
        data = (TEXT * 2)[:330000]
        got, writes, rc, st = _drive_c_writer(L, fmt, 6, bs, batch, data, rnd)
        assert rc == 0
        assert got == oracle.compress_stream(fmt, 6, bs, writes, {2, 7}), "fmt %d" % fmt
        assert st[0] == len(data) and st[1] == len(got) and st[2] >= len(data) // (bs * batch)
    # empty stream and a stream of exactly buffer_size bytes (strict '>' hold-back)
    for data in (b"", TEXT[:65280]):
        got, writes, rc, _ = _drive_c_writer(L, oracle.BGZF, 6, 65280, 2, data, rnd, flushes=())
        assert rc == 0 and got == oracle.compress_stream(oracle.BGZF, 6, 65280, writes)
        assert gzip.decompress(got) == data


def test_emu_c_writer_two_devices_round_robin(monkeypatch):
    """gzpb_writer_create_multi: batches dealt round-robin over two (emulated) devices give the same stream."""
    monkeypatch.setenv("GZPB_EMU_DEVICES", "2")
    L = emu.lib()
    rnd = random.Random(7)
    for fmt, bs in ((oracle.BGZF, 65280), (oracle.GZIP, 32768)):
        data = (TEXT * 3)[:560000]
        got, writes, rc, st = _drive_c_writer(L, fmt, 5, bs, 2, data, rnd, devices=(0, 1), flushes=(3,))
        assert rc == 0
        assert got == oracle.compress_stream(fmt, 5, bs, writes, {3})
        assert st[2] >= 6                                      # enough batches to wrap the 2 x 3 lanes
    import ctypes as C
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 5)
    from gzp_b200 import _lib
    sink = _lib.SINK_FN(lambda u, p, n: 0)
    assert L.gzpb_writer_create_multi(C.byref(h), devs, 2, oracle.BGZF, 6, 0, 2, C.cast(sink, C.c_void_p), None) == -8   # no device 5


def test_emu_c_writer_surfaces_sink_errors():
    """A failing sink (`W: Write` returning an error) fails the stream with GZPB_EIO on the next call (par/compress.rs:428-440)."""
    L = emu.lib()
    got, writes, rc, _ = _drive_c_writer(L, oracle.BGZF, 6, 65280, 1, TEXT * 2, random.Random(1), flushes=(1,), sink_fail_after=2)
    assert rc == -6


def test_emu_submit_poll_tickets():
    """gzpb_submit / gzpb_poll: tickets complete in order, GZPB_EAGAIN when every lane is in flight or the
    device is not done, outputs equal the synchronous gzpb_encode_batch; pinned blocks go by DMA in place."""
    import ctypes as C
    from gzp_b200 import _lib
    L = emu.lib()
    EAGAIN = -15
    h = C.c_void_p()
    assert L.gzpb_create(C.byref(h), 0, oracle.GZIP, 6, 40000, 4) == 0
    blocks = [TEXT[i * 40000:(i + 1) * 40000] for i in range(7)]
    pinned = L.gzpb_host_alloc(8 * 40000)
    C.memmove(pinned, TEXT[:7 * 40000], 7 * 40000)
    cap = L.gzpb_encode_capacity(oracle.GZIP, 40000) + 64
    want = [oracle.encode_block(oracle.GZIP, 6, b, (blocks[i - 1][-32768:] if i else None), i == 6) for i, b in enumerate(blocks)]

    def make(lo, hi, use_pinned):
        n = hi - lo
        ins, outs, keep = (_lib.BlockIn * n)(), (_lib.BlockOut * n)(), []
        for k, i in enumerate(range(lo, hi)):
            if use_pinned:
                ins[k].ptr = pinned + i * 40000
                if i:
                    ins[k].dict = pinned + i * 40000 - 32768
            else:
                src = C.create_string_buffer(blocks[i], 40000); keep.append(src)
                ins[k].ptr = C.cast(src, C.c_void_p)
                if i:
                    d = C.create_string_buffer(blocks[i - 1][-32768:], 32768); keep.append(d)
                    ins[k].dict = C.cast(d, C.c_void_p)
            ins[k].len = 40000
            ins[k].dict_len = 32768 if i else 0
            ins[k].is_last = int(i == 6)
            dst = C.create_string_buffer(cap); keep.append(dst)
            outs[k].dst = C.cast(dst, C.c_void_p); outs[k].cap = cap
        return ins, outs, keep

    for use_pinned in (True, False):
        batches = [make(0, 2, use_pinned), make(2, 4, use_pinned), make(4, 7, use_pinned), make(0, 1, use_pinned)]
        tickets = []
        for ins, outs, _ in batches[:3]:
            t = C.c_uint64(0)
            assert L.gzpb_submit(h, len(ins), ins, outs, C.byref(t)) == 0
            tickets.append(t.value)
        assert tickets == sorted(tickets) and len(set(tickets)) == 3
        t4 = C.c_uint64(0)
        assert L.gzpb_submit(h, 1, batches[3][0], batches[3][1], C.byref(t4)) == EAGAIN      # three lanes in flight
        assert L.gzpb_encode_batch(h, 1, batches[3][0], batches[3][1]) == EAGAIN             # tickets pending
        assert L.gzpb_poll(h, tickets[1], 0) == EAGAIN                                       # lazy device: nothing ran yet
        assert L.gzpb_poll(h, tickets[1], 1) == 0                                            # completes tickets 0 and 1
        assert L.gzpb_poll(h, tickets[0], 0) == 0
        assert L.gzpb_submit(h, 1, batches[3][0], batches[3][1], C.byref(t4)) == 0           # a lane is free again
        assert L.gzpb_poll(h, t4.value, 1) == 0                                              # also completes ticket 2
        assert L.gzpb_poll(h, t4.value + 1, 1) == -9                                         # unknown ticket
        got = []
        for ins, outs, _ in batches[:3]:
            for k in range(len(ins)):
                assert outs[k].status == 0
                got.append(C.string_at(outs[k].dst, outs[k].out_len))
                assert outs[k].check_sum == zlib.crc32(blocks[len(got) - 1]) and outs[k].check_amount == 40000
        assert got == want
        assert C.string_at(batches[3][1][0].dst, batches[3][1][0].out_len) == oracle.encode_block(oracle.GZIP, 6, blocks[0], None, False)
    # more blocks than a lane holds: invalid for submit, fine for the synchronous call
    ins, outs, _ = make(0, 7, False)
    t = C.c_uint64(0)
    assert L.gzpb_submit(h, 7, ins, outs, C.byref(t)) == -9
    assert L.gzpb_encode_batch(h, 7, ins, outs) == 0
    assert [C.string_at(outs[k].dst, outs[k].out_len) for k in range(7)] == want
    L.gzpb_host_free(pinned)
    L.gzpb_destroy(h)


def test_emu_native_parcompress_wrapper(emu_backend):
    """ParCompressBuilder(...).devices([...]) -> NativeParCompress (the C writer behind the reference's writer API)."""
    import gzp_b200
    sink = io.BytesIO()
    w = gzp_b200.ParCompressBuilder(gzp_b200.Gzip).compression_level(6).buffer_size(40000).blocks_in_flight(2).devices([0]).from_writer(sink)
    assert isinstance(w, gzp_b200.NativeParCompress)
    writes = [TEXT[:100000], TEXT[100000:100001], TEXT[100001:]]
    w.write(writes[0]); w.flush(); w.write(writes[1]); w.write(writes[2])
    st = w.stats()
    assert st["bytes_in"] == len(TEXT)
    assert w.finish() is sink
    assert sink.getvalue() == oracle.compress_stream(oracle.GZIP, 6, 40000, writes, {0})
    assert gzip.decompress(sink.getvalue()) == TEXT
    with pytest.raises(gzp_b200.GzpError):
        w.write(b"x")

    class Broken:
        def write(self, b):
            raise OSError("disk full")

    w = gzp_b200.ParCompressBuilder(gzp_b200.Bgzf).blocks_in_flight(1).devices([0]).from_writer(Broken())
    with pytest.raises(gzp_b200.GzpError) as ei:
        w.write(TEXT)
        w.finish()
    assert ei.value.variant == "Io" and isinstance(ei.value.__cause__, OSError)
    # Drop finishes the stream (par/compress.rs:391-402)
    sink = io.BytesIO()
    with gzp_b200.ParCompressBuilder(gzp_b200.Bgzf).devices([0]).from_writer(sink) as w:
        w.write(b"abc")
    assert gzip.decompress(sink.getvalue()) == b"abc"


def test_emu_compress_file_and_zero_copy_reserve(emu_backend, tmp_path):
    """gzpb_compress_file (read(2) into the pinned slabs -> ordered output file) and the reserve/commit form of
    write produce the oracle's stream."""
    import ctypes as C
    import gzp_b200
    from gzp_b200 import _lib
    src, dst = tmp_path / "in.txt", tmp_path / "out.gz"
    src.write_bytes(TEXT * 2)
    n_in, n_out = gzp_b200.compress_file(src, dst, gzp_b200.Gzip, 6, 40000, devices=(0,), blocks_in_flight=3)
    got = dst.read_bytes()
    assert (n_in, n_out) == (2 * len(TEXT), len(got))
    assert got == oracle.compress_stream(oracle.GZIP, 6, 40000, [TEXT * 2])
    assert gzip.decompress(got) == TEXT * 2
    with pytest.raises(gzp_b200.GzpError) as ei:
        gzp_b200.compress_file(tmp_path / "missing", dst)
    assert ei.value.variant == "Io"
    # reserve / commit: bytes produced in place, odd piece sizes
    L = emu.lib()
    chunks = bytearray()
    sink = _lib.SINK_FN(lambda u, p, n: (chunks.extend(C.string_at(p, n)), 0)[1])
    h = C.c_void_p()
    assert L.gzpb_writer_create(C.byref(h), 0, oracle.BGZF, 6, 65280, 2, C.cast(sink, C.c_void_p), None) == 0
    pos, rnd = 0, random.Random(5)
    while pos < len(TEXT):
        p, room = C.c_void_p(0), C.c_size_t(0)
        assert L.gzpb_writer_reserve(h, C.byref(p), C.byref(room)) == 0 and room.value > 0
        k = min(room.value, len(TEXT) - pos, rnd.randrange(1, 200000))
        C.memmove(p, TEXT[pos:pos + k], k)
        assert L.gzpb_writer_commit(h, k) == 0
        pos += k
    assert L.gzpb_writer_commit(h, 1 << 40) == -9
    assert L.gzpb_writer_finish(h) == 0
    L.gzpb_writer_destroy(h)
    assert bytes(chunks) == oracle.compress_stream(oracle.BGZF, 6, 65280, [TEXT])


def test_emu_writer_emits_the_gzi_index(emu_backend):
    """The Bgzf writer's own .gzi (collected as batches retire) equals the index scanned from the finished stream,
    including empty flush blocks and the EOF marker, which carry no entry."""
    import gzp_b200
    sink = io.BytesIO()
    w = gzp_b200.ParCompressBuilder(gzp_b200.Bgzf).compression_level(4).blocks_in_flight(2).devices([0]).from_writer(sink)
    w.write(TEXT[:70000]); w.flush(); w.flush(); w.write(TEXT[70000:])
    w.finish()
    stream = sink.getvalue()
    assert gzip.decompress(stream) == TEXT
    idx = w.bgzf_index()
    assert idx == gzp_b200.bgzf_index(stream)
    n = int.from_bytes(idx[:8], "little")
    assert n >= 4 and len(idx) == 8 + 16 * n
    # every entry points at a member whose decoded bytes start at the recorded uncompressed offset
    for k in range(n):
        coff = int.from_bytes(idx[8 + 16 * k:16 + 16 * k], "little")
        uoff = int.from_bytes(idx[16 + 16 * k:24 + 16 * k], "little")
        bsize = int.from_bytes(stream[coff + 16:coff + 18], "little") + 1
        assert gzip.decompress(stream[coff:coff + bsize]) == TEXT[uoff:uoff + len(gzip.decompress(stream[coff:coff + bsize]))]
    gz = gzp_b200.ParCompressBuilder(gzp_b200.Gzip).devices([0]).from_writer(io.BytesIO())
    with pytest.raises(gzp_b200.GzpError):
        gz.bgzf_index()
    gz.finish()


def test_emu_c_writer_block_size_exceeded_fails_the_stream():
    """Bgzf with buffer_size 65536 and incompressible bytes: the stored block (len + 5 + 26) exceeds 65536 ->
    BlockSizeExceeded (bgzf.rs:218-223); the writer surfaces it on a later call and stays failed (par/compress.rs:428-440)."""
    import ctypes as C
    from gzp_b200 import _lib
    L = emu.lib()
    rnd = random.Random(11)
    data = bytes(rnd.getrandbits(8) for _ in range(200000))
    out = bytearray()
    sink = _lib.SINK_FN(lambda u, p, n: (out.extend(C.string_at(p, n)), 0)[1])
    h = C.c_void_p()
    assert L.gzpb_writer_create(C.byref(h), 0, oracle.BGZF, 6, 65536, 1, C.cast(sink, C.c_void_p), None) == 0
    rcs = [L.gzpb_writer_write(h, data, len(data)), L.gzpb_writer_flush(h)]
    assert -3 in rcs, rcs                                     # GZPB_EBLOCKSIZE
    assert L.gzpb_writer_write(h, b"more", 4) == -3           # the stream stays failed
    assert L.gzpb_writer_finish(h) == -3
    L.gzpb_writer_destroy(h)
    assert len(out) == 0                                      # no partial output after the failing block's batch


def test_emu_native_pardecompress_reader_object(emu_backend):
    """gzpb_reader_* (ParDecompress as a C object) behind NativeParDecompress: members straddling source reads and
    chunks, both block formats, the error paths of par/decompress.rs:176-181, 193-197."""
    import gzp_b200

    class Dribble(io.RawIOBase):
        def __init__(self, data):
            self.d, self.p, self.k = data, 0, 0

        def read(self, n=-1):
            self.k += 1
            step = min(n if n > 0 else 1 << 30, 7919 * (1 + self.k % 5))
            b = self.d[self.p:self.p + step]
            self.p += len(b)
            return b

    for fmt, F, bs in ((oracle.BGZF, gzp_b200.Bgzf, 65280), (oracle.MGZIP, gzp_b200.Mgzip, 131072)):
        comp = oracle.compress_stream(fmt, 6, bs, [TEXT])
        r = gzp_b200.NativeParDecompress(F(), Dribble(comp), blocks_in_flight=3, chunk_bytes=65536)
        got = bytearray()
        while True:
            b = r.read(33333)
            if not b:
                break
            got += b
        r.finish()
        assert bytes(got) == TEXT
        r = gzp_b200.ParDecompressBuilder(F).native().from_reader(io.BytesIO(comp))
        assert isinstance(r, gzp_b200.NativeParDecompress) and r.read() == TEXT
        r.close()
    comp = oracle.compress_stream(oracle.BGZF, 6, 65280, [TEXT])
    assert gzp_b200.NativeParDecompress(gzp_b200.Bgzf(), io.BytesIO(b"")).read() == b""          # empty input = EOF
    bad = bytearray(comp); bad[-28 - 8] ^= 1                                                       # CRC of the last data block
    r = gzp_b200.NativeParDecompress(gzp_b200.Bgzf(), io.BytesIO(bytes(bad)))
    with pytest.raises(gzp_b200.GzpError) as ei:
        r.read()
    assert ei.value.variant == "InvalidCheck" and ei.value.found != ei.value.expected
    with pytest.raises(gzp_b200.GzpError):                                                         # errors are sticky
        r.read(10)
    r.close()
    r = gzp_b200.NativeParDecompress(gzp_b200.Bgzf(), io.BytesIO(comp[:-40]))                      # truncated member
    with pytest.raises(gzp_b200.GzpError) as ei:
        r.read()
    assert ei.value.variant == "Io"
    hdr = bytearray(comp); hdr[12] = ord("X")                                                      # "Bad SID"
    with pytest.raises(gzp_b200.GzpError) as ei:
        gzp_b200.NativeParDecompress(gzp_b200.Bgzf(), io.BytesIO(bytes(hdr))).read()
    assert ei.value.variant == "InvalidHeader"
    # a bad member BEHIND good ones: the good ones are delivered first (the reference's reader thread has already sent
    # them to the workers, decompress.rs:190-207), then the error
    sizes, pos = [], 0
    while pos < len(comp):
        sizes.append(struct.unpack_from("<H", comp, pos + 16)[0] + 1); pos += sizes[-1]
    third = sum(sizes[:3])
    for mutate, variant in ((lambda b: b.__setitem__(third + 12, ord("X")), "InvalidHeader"),
                            (lambda b: b.__setitem__(slice(third + sizes[3] - 4, third + sizes[3]), b"\xf0\xff\xff\xff"), "LibDelfaterDecompress")):
        bad = bytearray(comp); mutate(bad)
        r = gzp_b200.NativeParDecompress(gzp_b200.Bgzf(), io.BytesIO(bytes(bad)))
        got = bytearray()
        with pytest.raises(gzp_b200.GzpError) as ei:
            while True:
                b = r.read(65280)
                assert b
                got += b
        assert ei.value.variant == variant, ei.value.variant
        assert len(got) == 3 * 65280 and bytes(got) == TEXT[:3 * 65280]                      # an implausible ISIZE is corrupt data, not an allocation request
        r.close()

    class Failing:
        def read(self, n):
            raise OSError("network gone")

    with pytest.raises(gzp_b200.GzpError) as ei:
        gzp_b200.NativeParDecompress(gzp_b200.Bgzf(), Failing()).read()
    assert ei.value.variant == "Io" and isinstance(ei.value.__cause__, OSError)


def test_emu_memcheck_under_address_sanitizer():
    """compute-sanitizer memcheck without a GPU: the emulator built with -fsanitize=address (exact-size "device",
    pinned and dynamic-shared allocations) runs edge-size encodes, the ticket API, Snap and the decoder in a
    subprocess; any out-of-bounds access of a kernel or of the host runtime aborts it with an ASan report."""
    import os
    import subprocess
    import sys
    asan = subprocess.run(["/usr/bin/gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip() \
        if os.path.exists("/usr/bin/gcc") else ""
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no libasan on this machine")
    env = dict(os.environ, GZPB_EMU_ASAN="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    sel = "bgzf_edges or emu_snap or block_size_exceeded or c_writer_two_devices or copy_threads or reader_object or mgzip_long"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", sel, "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=1500, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert "AddressSanitizer" not in r.stdout + r.stderr, (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("sched", ["reverse", "random:7"])
def test_emu_results_do_not_depend_on_the_fiber_schedule(sched):
    """racecheck-lite: between rendezvous points the emulator may run the threads of a CTA in any order; a kernel whose
    output changes with that order is missing a __syncwarp / __syncthreads.  A subset of the parity tests is repeated in
    a subprocess with the order reversed and with a pseudo-random order per scheduling pass."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, GZPB_EMU_SCHED=sched)
    sel = "bgzf_levels and (6 or 1 or 9) or emu_snap or inflate_roundtrip or mgzip_long"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", sel, "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=1500, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert " passed" in r.stdout


def test_emu_units_longer_than_512k_checksum():
    """Regression: k_check's x^(8*512*j) table spans 512 KiB; longer units (up to 4 MiB) need the second table —
    the Mgzip / Gzip CRC-32 of a 590 000-byte block was wrong before."""
    big = (TEXT * 3)[:600000]
    assert gzip.decompress(_run(oracle.MGZIP, 2, 590000, big)) == big


def test_emu_encode_stream_multi_devices(monkeypatch):
    """gzpb_encode_stream_multi: ONE ordered stream whose device batches are dealt round-robin over three (emulated)
    GPUs — the offset chain crosses devices through mapped pinned memory — is byte-identical to the single-device
    stream and to the oracle, for pinned (zero-copy gather) and pageable buffers, all container families."""
    import ctypes as C
    monkeypatch.setenv("GZPB_EMU_DEVICES", "3")
    L = emu.lib()
    for fmt, lvl, bs, n in ((oracle.BGZF, 6, 65280, 560_000), (oracle.GZIP, 5, 32768, 300_000), (oracle.ZLIB, 2, 40000, 250_000),
                            (oracle.SNAP, 0, 65536, 500_000)):
        data = (TEXT * (n // len(TEXT) + 1))[:n]
        want = oracle.compress_stream(fmt, lvl, bs, [data])
        hs = (C.c_void_p * 3)()
        for d in range(3):
            h = C.c_void_p()
            assert L.gzpb_create(C.byref(h), d, fmt, lvl, bs, 2) == 0
            hs[d] = h
        cap = n + n // 4 + (1 << 16)
        olen = C.c_size_t(0)
        # pageable in / out
        out = C.create_string_buffer(cap)
        assert L.gzpb_encode_stream_multi(hs, 3, data, n, bs, out, cap, C.byref(olen)) == 0
        assert out.raw[:olen.value] == want
        # pinned in / out: zero-copy gather at the final stream offsets, chained across devices
        pin_in, pin_out = L.gzpb_host_alloc(n), L.gzpb_host_alloc(cap)
        C.memmove(pin_in, data, n)
        assert L.gzpb_encode_stream_multi(hs, 3, pin_in, n, bs, pin_out, cap, C.byref(olen)) == 0
        assert C.string_at(pin_out, olen.value) == want
        # one context (= gzpb_encode_stream), another device than the stream's first
        assert L.gzpb_encode_stream(hs[2], pin_in, n, bs, pin_out, cap, C.byref(olen)) == 0
        assert C.string_at(pin_out, olen.value) == want
        # a context listed twice is refused
        dup = (C.c_void_p * 2)(hs[0], hs[0])
        assert L.gzpb_encode_stream_multi(dup, 2, pin_in, n, bs, pin_out, cap, C.byref(olen)) == -9
        L.gzpb_host_free(pin_in); L.gzpb_host_free(pin_out)
        for d in range(3):
            L.gzpb_destroy(hs[d])


def test_emu_long_units_carried_heads_window_edges():
    """Long units are linked sub-unit by sub-unit with the bucket heads carried along (positions mod 65536, stale heads
    parked on a sentinel): data whose repeats sit exactly at, just inside and just outside the 32 KiB window, and
    low-entropy data whose buckets are hit in every sub-unit, over units several 64 KiB wraps long."""
    rnd = random.Random(4711)
    base = bytes(rnd.getrandbits(8) for _ in range(33000))
    for period in (32767, 32768, 32769, 20000):
        data = (base[:period] * 12)[:300000]
        assert gzip.decompress(_run(oracle.MGZIP, 6, 300000, data)) == data
    mixed = (synth.low_entropy(150000) + TEXT[:70000] + bytes(40000) + synth.fastq(140000))[:400000]
    assert gzip.decompress(_run(oracle.MGZIP, 5, 400000, mixed)) == mixed
    assert gzip.decompress(_run(oracle.GZIP, 6, 200000, mixed + mixed[:150000])) == mixed + mixed[:150000]


def test_emu_native_writer_copy_threads(emu_backend):
    """ParCompressBuilder.num_threads(n) with the native writer = n threads making the one host copy of `write`
    (gzpb_writer_set_copy_threads): writes of 8 MiB and more are cut into 2 MiB pieces copied side by side; same stream.
    (Level 0 keeps the emulated kernels cheap; the slab of 160 blocks holds 10 MB, so the first write is copied by the pool.)"""
    import gzp_b200
    data = (TEXT * 40)[:10_300_000]
    sink = io.BytesIO()
    w = gzp_b200.ParCompressBuilder(gzp_b200.Bgzf).compression_level(0).num_threads(4).blocks_in_flight(160).devices([0]).from_writer(sink)
    w.write(data[:9_500_000])
    w.write(data[9_500_000:])
    w.finish()
    assert sink.getvalue() == oracle.compress_stream(oracle.BGZF, 0, 65280, [data[:9_500_000], data[9_500_000:]])


def test_emu_launch_count_is_exact():
    """gzpb_launch_count (bench.py's `gpu_launches`) equals the kernels the emulator actually saw launched: per batch
    k_split, k_link (hash3 lists), k_link (hash4 lists), k_match, k_emit, k_scan, k_gather — and k_check instead of the
    first four at level 0, k_snap for Snap, k_check_combine per batch for the one-stream checksums."""
    L = emu.lib()
    L.gzpb_emu_kernel_launches.restype = C.c_ulonglong
    L.gzpb_launch_count.restype = C.c_uint64
    L.gzpb_launch_count.argtypes = [C.c_void_p]
    for fmt, level, bs, nbytes, per_batch in ((oracle.BGZF, 6, 0, 65280 * 5 + 7, 7), (oracle.BGZF, 0, 0, 65280 * 2, 4),
                                              (oracle.GZIP, 6, 65536, 65536 * 3, None), (oracle.SNAP, 0, 0, 131072 * 2, 3)):
        ctx = emu.EmuContext(fmt, level, max_block_bytes=bs, max_blocks_in_flight=4)
        try:
            before = L.gzpb_emu_kernel_launches()
            ctx.encode_stream(TEXT[:nbytes], bs)
            seen = L.gzpb_emu_kernel_launches() - before
            claimed = L.gzpb_launch_count(ctx.h)
        finally:
            ctx.close()
        assert claimed == seen and seen > 0, (fmt, level, claimed, seen)
        if per_batch:
            blk = bs or oracle.DEFAULT_BUFSIZE[fmt]
            nbatches = -(-max(1, -(-nbytes // blk)) // 4)
            assert seen == per_batch * nbatches, (fmt, level, seen, per_batch, nbatches)


def test_emu_arbitrary_dictionary_lengths():
    """gzpb_encode_batch with preset dictionaries of any length (the reference only ever passes 32 KiB, the ABI does not
    say so): the unit starts anywhere inside a match-table tile, so k_emit's tile ring starts at any slot and phase."""
    from gzp_b200 import _lib
    L = emu.lib()
    h = C.c_void_p()
    assert L.gzpb_create(C.byref(h), 0, oracle.RAWDEFLATE, 6, 50000, 4) == 0
    try:
        cap = L.gzpb_encode_capacity(oracle.RAWDEFLATE, 50000) + 64
        for dl in (1, 100, 127, 128, 129, 383, 384, 385, 5000, 32767, 32768):
            d, data = TEXT[70000 - dl:70000], TEXT[70000:70000 + 45000]
            ins, outs = (_lib.BlockIn * 1)(), (_lib.BlockOut * 1)()
            src, dic, dst = C.create_string_buffer(data, len(data)), C.create_string_buffer(d, dl), C.create_string_buffer(cap)
            ins[0].ptr = C.cast(src, C.c_void_p); ins[0].len = len(data)
            ins[0].dict = C.cast(dic, C.c_void_p); ins[0].dict_len = dl; ins[0].is_last = 1
            outs[0].dst = C.cast(dst, C.c_void_p); outs[0].cap = cap
            assert L.gzpb_encode_batch(h, 1, ins, outs) == 0 and outs[0].status == 0
            got = C.string_at(outs[0].dst, outs[0].out_len)
            assert got == oracle.encode_block(oracle.RAWDEFLATE, 6, data, d, True), dl
            assert zlib.decompressobj(-15, zdict=d).decompress(got) == data
    finally:
        L.gzpb_destroy(h)
