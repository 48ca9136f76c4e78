"""BASELINE.json configs[1] at FULL size, as a size-independent property: every one of the 837 224 BGZF blocks of
the 54 653 940 000-byte synthetic text stream (period 5 465 394 B x 10 000, blocks drifting against the period
so all of them are distinct — SURVEY.md §8d) is encoded on the device (level 6), decoded again on the device by
the ParDecompress kernel, and compared with its input; block sizes, BSIZE / ISIZE fields, CRC-32 and the EOF
marker are checked on the way.  Nothing is copied to the host except a handful of counters."""
import ctypes as C

import pytest

BLOCK = 65280
IN_STRIDE = 65600
OUT_SLOT = 73728
EOF_LEN = 28


def build_descs(torch, packed, offs, lens, eof_on_last):
    """gzpb_block_desc[n] (as int64[n, 4]) from the compacted BGZF members of one device batch:
    header 18 B (bgzf.rs:274-303) | raw DEFLATE | CRC-32 LE | ISIZE LE (bgzf.rs:224-233).
    Also returns the BSIZE and ISIZE fields for checking."""
    n = lens.numel()
    start = offs[:-1]
    end = offs[1:].clone()
    if eof_on_last:
        end[-1] -= EOF_LEN
    ar = torch.arange(4, device=packed.device)

    def le32(pos):
        b = packed[pos[:, None] + ar].to(torch.int64)
        return b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16) | (b[:, 3] << 24)

    crc = le32(end - 8)
    isize = le32(end - 4)
    bsize = (packed[start + 16].to(torch.int64) | (packed[start + 17].to(torch.int64) << 8)) + 1
    desc = torch.empty((n, 4), dtype=torch.int64, device=packed.device)
    desc[:, 0] = start + 18
    desc[:, 1] = torch.arange(n, device=packed.device, dtype=torch.int64) * BLOCK
    desc[:, 2] = (end - start - 26) | (lens.to(torch.int64) << 32)
    desc[:, 3] = crc
    return desc, bsize, isize, end - start


def test_desc_builder_matches_scan_blocks():
    """CPU check of the helper above against gzpb_scan_blocks (the reader loop of ParDecompress) on an oracle stream."""
    import torch
    import oracle
    from gzp_b200 import _lib, synth
    L = _lib.load()
    data = synth.text(5 * BLOCK + 1234)
    comp = oracle.compress_stream(oracle.BGZF, 6, BLOCK, [data])
    nb = C.c_size_t(0)
    L.gzpb_scan_blocks(oracle.BGZF, comp, len(comp), None, 0, C.byref(nb), None, None)
    descs = (_lib.BlockDesc * nb.value)()
    L.gzpb_scan_blocks(oracle.BGZF, comp, len(comp), descs, nb.value, C.byref(nb), None, None)
    members = [d for d in descs if d.out_len]                     # drop the EOF marker member
    n = len(members)
    assert n == 6
    offs = torch.tensor([d.in_off - 18 for d in members] + [len(comp)], dtype=torch.int64)
    lens = torch.tensor([d.out_len for d in members], dtype=torch.int32)
    packed = torch.frombuffer(bytearray(comp), dtype=torch.uint8)
    desc, bsize, isize, size = build_descs(torch, packed, offs, lens, True)
    raw = bytes(desc.numpy().tobytes())
    for i, d in enumerate(members):
        got = _lib.BlockDesc.from_buffer_copy(raw[32 * i:32 * i + 32])
        assert (got.in_off, got.in_len, got.out_len, got.crc) == (d.in_off, d.in_len, d.out_len, d.crc)
        assert got.out_off == i * BLOCK
    assert bsize.tolist() == size.tolist() and isize.tolist() == lens.tolist()


@pytest.mark.gpu
def test_c2_full_size_every_block_round_trips_on_device():
    import torch
    import gzp_b200
    from gzp_b200 import _lib, synth
    L = _lib.load()
    dev = torch.device("cuda", 0)
    P = synth.TEXT_PERIOD
    total = P * 10000                                             # 54 653 940 000 B = "55 GiB" of BASELINE.json
    nblocks = (total + BLOCK - 1) // BLOCK
    assert nblocks == 837224
    B = 3256                                                      # 22 k_emit units per SM (the shipped batch size is 4736 = 32 per SM)
    S = torch.frombuffer(bytearray(synth.text_stream(P)), dtype=torch.uint8).to(dev)
    S_rep = S.repeat((B * BLOCK + P) // P + 2)
    ctx = gzp_b200.Context(gzp_b200.BGZF, 6, device=0, max_block_bytes=BLOCK, max_blocks_in_flight=B)
    dec = gzp_b200.Decoder(gzp_b200.BGZF, 0, B)
    d_in = torch.zeros((B, IN_STRIDE), dtype=torch.uint8, device=dev)
    d_len = torch.full((B,), BLOCK, dtype=torch.int32, device=dev)
    d_flags = torch.zeros((B,), dtype=torch.int32, device=dev)
    d_packed = torch.zeros((B * OUT_SLOT + 256,), dtype=torch.uint8, device=dev)
    d_off = torch.zeros((B + 1,), dtype=torch.int64, device=dev)
    d_status = torch.zeros((B,), dtype=torch.int32, device=dev)
    d_out = torch.zeros((B * BLOCK + 256,), dtype=torch.uint8, device=dev)
    d_dstatus = torch.zeros((B,), dtype=torch.int32, device=dev)
    d_crc = torch.zeros((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    bad = torch.zeros((), dtype=torch.int64, device=dev)          # batches with any failure
    comp_total = torch.zeros((), dtype=torch.int64, device=dev)
    max_member = torch.zeros((), dtype=torch.int64, device=dev)
    crc_xor = torch.zeros((), dtype=torch.int64, device=dev)      # a checksum of checksums
    eof_ok = None
    for b0 in range(0, nblocks, B):
        n = min(B, nblocks - b0)
        pos = b0 * BLOCK
        span = min(n * BLOCK, total - pos)
        start = pos % P
        flat = S_rep[start:start + span]
        full = span // BLOCK
        d_in[:full, :BLOCK] = flat[:full * BLOCK].view(full, BLOCK)
        last = b0 + n == nblocks
        if last:
            rem = span - full * BLOCK
            assert full == n - 1 and rem == total - (nblocks - 1) * BLOCK
            d_in[full, :rem] = flat[full * BLOCK:]
            d_len[full] = rem
            d_flags[full] = 1                                     # is_last: BGZF_EOF follows (bgzf.rs:24-38)
        rc = L.gzpb_encode_device(ctx._h, d_in.data_ptr(), d_len.data_ptr(), d_flags.data_ptr(), n, d_packed.data_ptr(),
                                  d_off.data_ptr(), d_status.data_ptr(), st.cuda_stream)
        assert rc == 0
        desc, bsize, isize, size = build_descs(torch, d_packed, d_off[:n + 1], d_len[:n], last)
        rc = L.gzpb_decode_device(dec._h, d_packed.data_ptr(), desc.data_ptr(), n, d_out.data_ptr(), d_dstatus.data_ptr(),
                                  d_crc.data_ptr(), st.cuda_stream)
        assert rc == 0
        ok = (d_status[:n] == 0).all() & (d_dstatus[:n] == 0).all() & torch.equal(d_out[:span], flat) \
            & (bsize == size).all() & (isize == d_len[:n]).all() & (size < 65536).all() \
            & ((d_crc[:n].to(torch.int64) & 0xFFFFFFFF) == desc[:, 3]).all()
        bad += (~ok).to(torch.int64)
        comp_total += d_off[n]
        max_member = torch.maximum(max_member, size.max())
        crc_xor ^= desc[:, 3].sum()
        if last:
            eof_ok = bytes(d_packed[int(d_off[n].item()) - EOF_LEN:int(d_off[n].item())].cpu().numpy().tobytes()) == gzp_b200.BGZF_EOF
    torch.cuda.synchronize()
    ctx.close(); dec.close()
    assert int(bad.item()) == 0, "some block did not round-trip"
    assert eof_ok
    ratio = int(comp_total.item()) / total
    assert 0.38 < ratio < 0.40, ratio                             # zlib-1.3 proxy of the survey: 0.3909
    assert int(max_member.item()) < 65536                         # BlockSizeExceeded never trips (bgzf.rs:218-223)
    assert int(crc_xor.item()) != 0
