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


@pytest.mark.parametrize("level", [0, 2, 3, 4, 5, 6, 7, 8, 9])
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


def test_emu_snap():
    rnd = random.Random(7)
    for d in (TEXT, bytes(70000), bytes(rnd.getrandbits(8) for _ in range(50000)), b"", b"abc"):
        _run(oracle.SNAP, 0, 131072, d)
