"""The incremental writer and the asynchronous ticket API on the GPU (include/gzpb.h: gzpb_writer_*,
gzpb_writer_create_multi, gzpb_submit / gzpb_poll) — ParCompress::write / flush / finish
(/root/reference/src/par/compress.rs:332-362, 377-388, 413-468) with real streams, real DMA and real
overlap; bit-exact against the oracle's ParCompress model and decoded by stock gzip."""
import ctypes as C
import gzip
import io
import random
import zlib

import pytest

import oracle
import gzp_b200
from gzp_b200 import BGZF, GZIP, MGZIP, SNAP, ZLIB, _lib, synth

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _write_all(w, data, rnd, bs, flush_at=()):
    writes, pos, i = [], 0, 0
    while pos < len(data):
        k = 65536 if rnd.random() < 0.8 else rnd.randrange(1, 3 * bs)       # benches/bench.rs:121 writes 64 KiB pieces
        writes.append(data[pos:pos + k]); pos += k
    for i, x in enumerate(writes):
        w.write(x)
        if i in flush_at:
            w.flush()
    return writes


@pytest.mark.parametrize("fmt,F,bs,level", [(BGZF, gzp_b200.Bgzf, 65280, 6), (GZIP, gzp_b200.Gzip, 131072, 4),
                                            (MGZIP, gzp_b200.Mgzip, 131072, 6), (SNAP, gzp_b200.Snap, 131072, 6),
                                            (ZLIB, gzp_b200.Zlib, 40000, 2)])
def test_native_writer_many_batches_in_flight(fmt, F, bs, level):
    """~20 MB through the pipelined writer with batches of 37 blocks (several slabs and lanes wrap around)."""
    data = synth.text_stream(20_000_000)
    rnd = random.Random(fmt)
    sink = io.BytesIO()
    w = gzp_b200.ParCompressBuilder(F).compression_level(level).buffer_size(bs).blocks_in_flight(37).devices([0]).from_writer(sink)
    writes = _write_all(w, data, rnd, bs, flush_at={5, 40})
    st = w.stats()
    w.finish()
    got = sink.getvalue()
    assert st["bytes_in"] == len(data) and st["batches"] >= 3
    assert got == oracle.compress_stream(fmt, level, bs, writes, {5, 40})
    if fmt in (BGZF, MGZIP, GZIP):
        assert gzip.GzipFile(fileobj=io.BytesIO(got)).read() == data
    elif fmt == ZLIB:
        assert zlib.decompress(got) == data


def test_native_writer_default_batch_roundtrip_on_device():
    """Default batch size (1184 blocks), 300 MB of text: the stream decodes to the input with the GPU decoder
    (ParDecompress path) and with stock gzip on a prefix; sizes add up."""
    data = synth.text_stream(300_000_000)
    sink = io.BytesIO()
    w = gzp_b200.ParCompressBuilder(gzp_b200.Bgzf).compression_level(6).blocks_in_flight(0).devices([0]).from_writer(sink)
    mv = memoryview(data)
    for off in range(0, len(data), 1 << 20):
        w.write(mv[off:off + (1 << 20)])
    st = w.stats()
    w.finish()
    comp = sink.getvalue()
    assert st["batches"] >= 3 and 0.3 < len(comp) / len(data) < 0.5
    dec = gzp_b200.Decoder(BGZF)
    out, used = dec.decode(comp)
    dec.close()
    assert used == len(comp) and out == data
    assert comp[-28:] == gzp_b200.BGZF_EOF


def test_native_writer_over_all_gpus():
    """gzpb_writer_create_multi: batches dealt round-robin over every GPU of the box give the
    same bytes as one GPU."""
    n = _ndev()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    data = synth.text_stream(40_000_000)
    outs = []
    for devs in ([0], list(range(n))):
        sink = io.BytesIO()
        w = gzp_b200.ParCompressBuilder(gzp_b200.Gzip).compression_level(6).buffer_size(131072).blocks_in_flight(16).devices(devs).from_writer(sink)
        _write_all(w, data, random.Random(3), 131072, flush_at={9})
        w.finish()
        outs.append(sink.getvalue())
    assert outs[0] == outs[1]
    assert gzip.decompress(outs[1]) == data


def test_submit_poll_tickets_on_gpu(text_corpus):
    L = _lib.load()
    h = C.c_void_p()
    bs, nblk, per = 65280, 12, 4                                 # 3 batches of 4 blocks = every lane in flight
    assert nblk * bs <= len(text_corpus)
    assert L.gzpb_create(C.byref(h), 0, BGZF, 6, bs, per) == 0
    blocks = [text_corpus[i * bs:(i + 1) * bs] for i in range(nblk)]
    want = [oracle.encode_block(BGZF, 6, b, None, False) for b in blocks]
    pinned = L.gzpb_host_alloc(nblk * bs)
    C.memmove(pinned, text_corpus[:nblk * bs], nblk * bs)
    cap = L.gzpb_encode_capacity(BGZF, bs) + 64
    for use_pinned in (True, False):
        keep, batches = [], []
        for lo in range(0, nblk, per):
            ins, outs = (_lib.BlockIn * per)(), (_lib.BlockOut * per)()
            for k in range(per):
                if use_pinned:
                    ins[k].ptr = pinned + (lo + k) * bs
                else:
                    src = C.create_string_buffer(blocks[lo + k], bs); keep.append(src)
                    ins[k].ptr = C.cast(src, C.c_void_p)
                ins[k].len = bs
                dst = C.create_string_buffer(cap); keep.append(dst)
                outs[k].dst = C.cast(dst, C.c_void_p); outs[k].cap = cap
            batches.append((ins, outs))
        tickets = []
        for ins, outs in batches:
            t = C.c_uint64(0)
            assert L.gzpb_submit(h, per, ins, outs, C.byref(t)) == 0
            tickets.append(t.value)
        t = C.c_uint64(0)
        assert L.gzpb_submit(h, per, batches[0][0], batches[0][1], C.byref(t)) == -15          # GZPB_EAGAIN: 3 lanes in flight
        while True:                                                                          # non-blocking poll until done
            rc = L.gzpb_poll(h, tickets[-1], 0)
            if rc != -15:
                break
        assert rc == 0
        got = [C.string_at(outs[k].dst, outs[k].out_len) for ins, outs in batches for k in range(per)]
        assert got == want
    L.gzpb_host_free(pinned)
    L.gzpb_destroy(h)
