"""CPU tests of the drop-in boundary and the host logic: the C-ABI library loads and
exports every symbol include/gzpb.h declares, the host helpers agree with the oracle,
the library fails loudly without a GPU, and the ParCompress mirror chunks exactly
like the reference (/root/reference/src/par/compress.rs:332-362, 413-468)."""
import io
import os
import random
import re
import zlib

import pytest

import oracle
import gzp_b200
from gzp_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gzpb.h")).read()
    declared = set(re.findall(r"\b(gzpb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 18
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gzpb.h but not exported"
    assert declared <= set(_lib.EXPORTS)


def test_host_helpers_match_oracle():
    for fmt in range(6):
        for level in (0, 1, 3, 6, 9):
            assert gzp_b200.header(fmt, level) == oracle.header(fmt, level)
        assert gzp_b200.footer(fmt, 0xDEADBEEF, 1234) == oracle.footer(fmt, 0xDEADBEEF, 1234)
        for n in (0, 1, 1279, 1280, 1281, 65280, 131072):
            assert gzp_b200.encode_capacity(fmt, n) == oracle.lib().oracle_encode_capacity(fmt, n)
    rnd = random.Random(5)
    lib = _lib.load()
    for _ in range(40):
        a = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 3000)))
        b = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 3000)))
        assert gzp_b200.crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
        assert lib.gzpb_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
    assert lib.gzpb_default_bufsize(gzp_b200.BGZF) == 65280 and lib.gzpb_default_bufsize(gzp_b200.GZIP) == 131072
    assert [lib.gzpb_needs_dict(f) for f in range(6)] == [1, 1, 1, 0, 0, 0]
    assert b"65536" in lib.gzpb_strerror(-3)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_fails_loudly_without_gpu():
    with pytest.raises(gzp_b200.GzpError) as e:
        gzp_b200.Context(gzp_b200.BGZF, 6)
    assert e.value.variant == "Cuda"


def test_builder_validation_mirrors_reference():
    with pytest.raises(gzp_b200.GzpError) as e:
        gzp_b200.ParCompressBuilder(gzp_b200.Gzip).buffer_size(32767)      # par/compress.rs:68-74
    assert e.value.variant == "BufferSize"
    with pytest.raises(gzp_b200.GzpError) as e:
        gzp_b200.ParCompressBuilder(gzp_b200.Gzip).num_threads(0)          # par/compress.rs:84-90
    assert e.value.variant == "NumThreads"
    b = gzp_b200.ParCompressBuilder(gzp_b200.Bgzf)
    assert b._buffer_size == 65280 and b._level.level() == 3              # deflate.rs:583, par/compress.rs:58
    assert gzp_b200.ParCompressBuilder(gzp_b200.Mgzip)._buffer_size == 131072


class _OracleContext:
    """Test double for the device context: answers encode_blocks from the CPU oracle so the
    host-side chunker / ordering logic can be exercised without a GPU."""
    log = []

    def __init__(self, fmt, level, device=0, max_block_bytes=0, max_blocks_in_flight=256):
        self.fmt, self.level = fmt, api._lvl(level)

    def encode_blocks(self, blocks):
        _OracleContext.log.extend(blocks)
        out = []
        for data, d, last in blocks:
            enc = oracle.encode_block(self.fmt, self.level, data, d, last)
            s = zlib.crc32(data) if self.fmt == oracle.GZIP else zlib.adler32(data) if self.fmt == oracle.ZLIB else 0
            out.append((enc, s, len(data) if self.fmt in (oracle.GZIP, oracle.ZLIB) else 0))
        return out

    def close(self):
        pass


@pytest.mark.parametrize("fmt", [gzp_b200.Bgzf, gzp_b200.Gzip, gzp_b200.Zlib, gzp_b200.Mgzip, gzp_b200.RawDeflate, gzp_b200.Snap])
def test_parcompress_mirror_chunks_like_the_reference(monkeypatch, fmt, text_corpus):
    monkeypatch.setattr(api, "Context", _OracleContext)
    rnd = random.Random(11)
    bs = 65280 if fmt is gzp_b200.Bgzf else 40000
    data = text_corpus[: 5 * bs + 777]
    writes, pos = [], 0
    while pos < len(data):
        k = rnd.randrange(1, 3 * bs)
        writes.append(data[pos: pos + k]); pos += k
    flushes = {1, 3}
    _OracleContext.log = []
    sink = io.BytesIO()
    pc = gzp_b200.ParCompressBuilder(fmt).buffer_size(bs).compression_level(6).blocks_in_flight(3).from_writer(sink)
    for i, w in enumerate(writes):
        assert pc.write(w) == len(w)
        if i in flushes:
            pc.flush()
    assert pc.finish() is sink
    want_msgs = oracle.chunk_stream(fmt.ID, bs, writes, flushes)
    assert [(bytes(b), d, l) for b, d, l in _OracleContext.log] == want_msgs
    assert sink.getvalue() == oracle.compress_stream(fmt.ID, 6, bs, writes, flushes)
    if fmt is gzp_b200.Zlib:                                               # the Adler-32 footer folds per block (check.rs:121-128)
        assert zlib.decompress(sink.getvalue()) == data
    with pytest.raises(gzp_b200.GzpError):
        pc.write(b"x")                                                     # write after finish


def test_drop_finishes_the_stream(monkeypatch):
    # test_simple_drop (deflate.rs:745): leaving scope must finish the stream
    monkeypatch.setattr(api, "Context", _OracleContext)
    import gzip
    sink = io.BytesIO()
    with gzp_b200.ParCompressBuilder(gzp_b200.Gzip).from_writer(sink) as pc:
        pc.write(b"This is a first test line\nThis is a second test line\n")
    assert gzip.decompress(sink.getvalue()) == b"This is a first test line\nThis is a second test line\n"


def test_bgzf_gzi_index_and_virtual_offsets():
    """The .gzi entries (htslib layout) address real members: decoding the member at every
    compressed offset yields the input from the matching uncompressed offset."""
    import struct
    data = bytes(random.Random(9).getrandbits(8) for _ in range(1000)) * 300
    stream = oracle.compress_stream(oracle.BGZF, 6, 65280, [data[:100000], b"", data[100000:]], flushes={0, 1})
    idx = gzp_b200.bgzf_index(stream)
    n, = struct.unpack_from("<Q", idx, 0)
    assert len(idx) == 8 + 16 * n and n >= 3
    prev_u = 0
    for i in range(n):
        c, u = struct.unpack_from("<QQ", idx, 8 + 16 * i)
        assert u > prev_u
        prev_u = u
        size = struct.unpack_from("<H", stream, c + 16)[0] + 1
        member = zlib.decompressobj(31).decompress(stream[c:c + size])
        assert member == data[u:u + len(member)] and len(member) > 0
    assert gzp_b200.bgzf_virtual_offset(0x1234, 77) == (0x1234 << 16) | 77
    assert gzp_b200.bgzf_index(oracle.compress_stream(oracle.BGZF, 6, 65280, [b""])) == struct.pack("<Q", 0)


def test_shipped_sass_properties():
    """What the shipped library's machine code must keep (cross-compiled here; `cuobjdump -sass`): sm_100a only; TMA bulk
    copies + mbarrier waits in k_match / k_emit; k_link without MATCH.ANY (its ADU pipe bounded the kernel, profiles/README.md)
    and with its ballots; k_emit at 64 registers and under 7.3 KiB of shared memory (32 resident units per SM); k_split at 32
    registers (two CTAs of 1024 threads per SM)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    so = _lib.SO_PATH
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass and not re.search(r"arch = sm_(?!100a)", sass)
    funcs = {}
    for part in sass.split("Function : ")[1:]:
        name, body = part.split("\n", 1)
        funcs[name.strip()] = body
    def body_of(key):
        hits = [b for n, b in funcs.items() if key in n]
        assert hits, key
        return "\n".join(hits)
    assert "UBLKCP" in body_of("k_match") and "UBLKCP" in body_of("k_emit") and "SYNCS" in body_of("k_emit")
    link = body_of("k_link")
    assert "MATCH" not in link and link.count("VOTE") >= 5
    res = subprocess.run([cuobjdump, "--dump-resource-usage", so], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:\d+ SHARED:(\d+)", res):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    emit = [v for k, v in usage.items() if "k_emit" in k]
    assert emit and all(r <= 64 and s <= 7296 for r, s in emit), emit
    split = [v for k, v in usage.items() if "k_split" in k]
    assert split and all(r <= 32 for r, _ in split), split
