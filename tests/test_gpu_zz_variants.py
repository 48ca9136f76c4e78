"""Kernel-set variants on the GPU (runs last: the file name sorts after the other GPU tests).

GZPB_MATCH_V2=1  k_group + k_match2 (hash groups instead of linked chains) — measured in round 1, bit-identical, not
                 faster (profiles/README.md); kept as a checked alternative.
GZPB_SPARSE=1    k_smatch + k_emit<sparse> + filtered fallback: the match table only where the parser
                 looks.  Developed and held bit-exact on the CPU SIMT emulator (tests/test_emu_kernels.py, also under
                 schedule perturbation and AddressSanitizer); this file is its first contact with real hardware.
Both are opt-in through the environment at gzpb_create; the default kernel set is untouched by them."""
import ctypes as C
import random

import pytest

import oracle
import gzp_b200
from gzp_b200 import BGZF, GZIP, _lib, synth

# method="thread": a kernel that hangs blocks inside a C call, where a signal-based timeout never fires; the thread method
# ends the process instead (these are the last tests of the run)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]


def _encode(env, monkeypatch, fmt, level, bs, data, want_variant, value="1"):
    for k in ("GZPB_MATCH_V2", "GZPB_SPARSE"):
        monkeypatch.delenv(k, raising=False)
    if env:
        monkeypatch.setenv(env, value)
    L = _lib.load()
    ctx = gzp_b200.Context(fmt, level, max_block_bytes=bs, max_blocks_in_flight=64)
    try:
        assert L.gzpb_ctx_variant(ctx._h).decode() == want_variant
        L.gzpb_debug_sparse_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
        L.gzpb_debug_sparse_stats(ctx._h, None, None, 1)          # device-wide counters: start from zero
        got = ctx.encode_stream(data, bs)
        stats = None
        if env == "GZPB_SPARSE":
            u, m = C.c_uint64(0), C.c_uint64(0)
            L.gzpb_debug_sparse_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
            L.gzpb_debug_sparse_stats(ctx._h, C.byref(u), C.byref(m), 1)
            stats = (u.value, m.value)
    finally:
        ctx.close()
    return got, stats


def test_match_v2_is_bit_identical_on_gpu(monkeypatch, text_corpus):
    for level in (1, 4, 6, 9):
        data = text_corpus[:400000]
        got, _ = _encode("GZPB_MATCH_V2", monkeypatch, BGZF, level, 65280, data, "split+group+match2")
        assert got == oracle.compress_stream(BGZF, level, 65280, [data]), level


def test_sparse_match_table_is_bit_identical_on_gpu(monkeypatch, text_corpus):
    rnd = random.Random(77)
    rand = bytes(rnd.getrandbits(8) for _ in range(30000))
    few = bytes(rnd.choice(b"ACGT") for _ in range(40000))
    mixed = (few + text_corpus[:60000])[:65280]
    for level in (2, 4, 5, 6, 7):
        data = text_corpus[:1_000_000]
        got, (units, missed) = _encode("GZPB_SPARSE", monkeypatch, BGZF, level, 65280, data, "split+link+smatch")
        assert got == oracle.compress_stream(BGZF, level, 65280, [data]), level
        assert units == 16 and missed == 0, (level, units, missed)
    for data, want_missed in ((bytes(200000), 0), (synth.low_entropy(300000), None), (synth.fastq(200000), None), (rand, 0), (b"", 0),
                              (text_corpus[:65280] + mixed + text_corpus[:30000], 1)):
        got, (units, missed) = _encode("GZPB_SPARSE", monkeypatch, BGZF, 6, 65280, data, "split+link+smatch")
        assert got == oracle.compress_stream(BGZF, 6, 65280, [data]), len(data)
        if want_missed is not None:
            assert missed == want_missed, (len(data), units, missed)
    data = text_corpus[:500000]
    got, (units, missed) = _encode("GZPB_SPARSE", monkeypatch, GZIP, 6, 32768, data, "split+link+smatch")
    assert got == oracle.compress_stream(GZIP, 6, 32768, [data]) and missed == 0


def test_sparse_tokens_and_replay_is_bit_identical_on_gpu(monkeypatch, text_corpus):
    """GZPB_SPARSE=2: k_smatch hands the stitched tokens of the true parse to k_emit<2>, which only replays the events."""
    rnd = random.Random(77)
    few = bytes(rnd.choice(b"ACGT") for _ in range(40000))
    mixed = (few + text_corpus[:60000])[:65280]
    for level in (2, 5, 6, 7, 9):
        data = text_corpus[:1_000_000]
        got, (units, missed) = _encode("GZPB_SPARSE", monkeypatch, BGZF, level, 65280, data, "split+link+smatch+replay", "2")
        assert got == oracle.compress_stream(BGZF, level, 65280, [data]), level
        assert units == 16 and missed == 0, (level, units, missed)
    for data, want_missed in ((bytes(200000), 0), (synth.low_entropy(300000), None), (synth.fastq(200000), None), (b"", 0),
                              (text_corpus[:65280] + mixed + text_corpus[:30000], 0)):        # min_len changes: a new epoch, no fallback
        got, (units, missed) = _encode("GZPB_SPARSE", monkeypatch, BGZF, 6, 65280, data, "split+link+smatch+replay", "2")
        assert got == oracle.compress_stream(BGZF, 6, 65280, [data]), len(data)
        if want_missed is not None:
            assert missed == want_missed, (len(data), units, missed)
    # long units: Mgzip 131 072-byte blocks (configs[2]) and Gzip level 9 with 262 144-byte blocks on FASTQ-shaped data (configs[4])
    from gzp_b200 import MGZIP
    for fmt, level, bs, data in ((MGZIP, 6, 131072, text_corpus[:1_200_000]), (GZIP, 9, 262144, synth.fastq(1_000_000))):
        got, (units, missed) = _encode("GZPB_SPARSE", monkeypatch, fmt, level, bs, data, "split+link+smatch+replay", "2")
        assert got == oracle.compress_stream(fmt, level, bs, [data]), (fmt, level)
        assert units >= 4 and missed == 0
