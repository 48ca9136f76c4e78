"""CPU oracle for the gzp per-block encode path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``gzp_b200``) never does, and fails loudly when its CUDA library is missing.

The arithmetic is in the C files beside this one (each cites the reference
file:line it restates); this module is a ctypes loader plus a pure-Python model of
``ParCompress``'s chunker (/root/reference/src/par/compress.rs:332-362, 413-463).
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

GZIP, ZLIB, RAWDEFLATE, MGZIP, BGZF, SNAP = 0, 1, 2, 3, 4, 5
DICT_SIZE = 32768
DEFAULT_BUFSIZE = {GZIP: 131072, ZLIB: 131072, RAWDEFLATE: 131072, MGZIP: 131072, BGZF: 65280, SNAP: 131072}
NEEDS_DICT = {GZIP: True, ZLIB: True, RAWDEFLATE: True, MGZIP: False, BGZF: False, SNAP: False}


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        c = ctypes
        L.oracle_deflate.restype = c.c_size_t
        L.oracle_deflate.argtypes = [c.c_char_p, c.c_size_t, c.c_int, c.c_char_p, c.c_size_t]
        L.oracle_deflate_ex.restype = c.c_size_t
        L.oracle_deflate_ex.argtypes = [c.c_char_p, c.c_size_t, c.c_size_t, c.c_int, c.c_int, c.c_char_p, c.c_size_t, c.c_void_p]
        L.oracle_crc32.restype = c.c_uint32
        L.oracle_crc32.argtypes = [c.c_uint32, c.c_char_p, c.c_size_t]
        L.oracle_crc32c.restype = c.c_uint32
        L.oracle_crc32c.argtypes = [c.c_uint32, c.c_char_p, c.c_size_t]
        L.oracle_crc32c_masked.restype = c.c_uint32
        L.oracle_crc32c_masked.argtypes = [c.c_char_p, c.c_size_t]
        L.oracle_crc32_combine.restype = c.c_uint32
        L.oracle_crc32_combine.argtypes = [c.c_uint32, c.c_uint32, c.c_uint64]
        L.oracle_adler32.restype = c.c_uint32
        L.oracle_adler32.argtypes = [c.c_uint32, c.c_char_p, c.c_size_t]
        L.oracle_adler32_combine.restype = c.c_uint32
        L.oracle_adler32_combine.argtypes = [c.c_uint32, c.c_uint32, c.c_uint64]
        L.oracle_snappy_raw.restype = c.c_size_t
        L.oracle_snappy_raw.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p]
        L.oracle_snappy_frame.restype = c.c_size_t
        L.oracle_snappy_frame.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p, c.c_size_t]
        L.oracle_snappy_max_compress_len.restype = c.c_size_t
        L.oracle_snappy_max_compress_len.argtypes = [c.c_size_t]
        L.oracle_encode_capacity.restype = c.c_size_t
        L.oracle_encode_capacity.argtypes = [c.c_int, c.c_size_t]
        L.oracle_encode_block.restype = c.c_long
        L.oracle_encode_block.argtypes = [c.c_int, c.c_int, c.c_char_p, c.c_size_t, c.c_char_p, c.c_size_t, c.c_int, c.c_char_p, c.c_size_t]
        L.oracle_header.restype = c.c_size_t
        L.oracle_header.argtypes = [c.c_int, c.c_int, c.c_char_p]
        L.oracle_footer.restype = c.c_size_t
        L.oracle_footer.argtypes = [c.c_int, c.c_uint32, c.c_uint32, c.c_char_p]
        L.oracle_par_compress.restype = c.c_double
        L.oracle_par_compress.argtypes = [c.c_int, c.c_int, c.c_size_t, c.c_int, c.c_char_p, c.c_size_t, c.c_char_p, c.c_size_t, c.POINTER(c.c_size_t)]
        L.oracle_make_huffman_code.restype = None
        L.oracle_make_huffman_code.argtypes = [c.c_uint, c.c_uint, c.POINTER(c.c_uint32), c.POINTER(c.c_uint8), c.POINTER(c.c_uint32)]
        L.oracle_inflate.restype = c.c_long
        L.oracle_inflate.argtypes = [c.c_char_p, c.c_size_t, c.c_char_p, c.c_size_t]
        L.oracle_block_size.restype = c.c_long
        L.oracle_block_size.argtypes = [c.c_int, c.c_char_p, c.c_size_t]
        L.oracle_decode_stream.restype = c.c_int
        L.oracle_decode_stream.argtypes = [c.c_int, c.c_char_p, c.c_size_t, c.c_char_p, c.c_size_t, c.POINTER(c.c_size_t),
                                           c.POINTER(c.c_uint32), c.POINTER(c.c_uint32)]
        L.oracle_level_supported.restype = c.c_int
        L.oracle_level_supported.argtypes = [c.c_int]
        _lib = L
    return _lib


def deflate(data, level=6):
    """libdeflate_deflate_compress() restatement (raw DEFLATE, BFINAL set)."""
    cap = len(data) + max(128, len(data) // 10) + 16
    out = ctypes.create_string_buffer(cap)
    n = lib().oracle_deflate(bytes(data), len(data), level, out, cap)
    if n == 0:
        raise ValueError("oracle_deflate: output did not fit / unsupported level")
    return out.raw[:n]


def deflate_ex(data, level=6, dictionary=b"", flush=0):
    """oracle_deflate_ex: raw DEFLATE of `data` primed with `dictionary`; flush 0 = finish, 1 = sync flush."""
    L = lib()
    buf = bytes(dictionary) + bytes(data)
    cap = len(data) + max(128, len(data) // 10) + 64
    out = ctypes.create_string_buffer(cap)
    n = L.oracle_deflate_ex(buf, len(dictionary), len(data), level, flush, out, cap, None)
    if n == 0:
        raise ValueError("oracle_deflate_ex: output did not fit / unsupported level")
    return out.raw[:n]


def encode_block(fmt, level, data, dictionary=None, is_last=False):
    """FormatSpec::encode for one block (returns bytes, raises ValueError(code))."""
    data = bytes(data)
    cap = lib().oracle_encode_capacity(fmt, len(data)) + 64
    out = ctypes.create_string_buffer(cap)
    d = bytes(dictionary) if dictionary else None
    r = lib().oracle_encode_block(fmt, level, data, len(data), d, len(d) if d else 0, int(is_last), out, cap)
    if r < 0:
        raise ValueError(int(r))
    return out.raw[:r]


def header(fmt, level):
    b = ctypes.create_string_buffer(16)
    n = lib().oracle_header(fmt, level, b)
    return b.raw[:n]


def footer(fmt, total_sum, amount):
    b = ctypes.create_string_buffer(16)
    n = lib().oracle_footer(fmt, total_sum, amount, b)
    return b.raw[:n]


def crc32(data, crc=0):
    return lib().oracle_crc32(crc, bytes(data), len(data))


def crc32c(data, crc=0):
    return lib().oracle_crc32c(crc, bytes(data), len(data))


def chunk_stream(fmt, buffer_size, writes, flushes=()):
    """Pure-Python model of ParCompress::write / flush / finish: returns the list of
    messages (block bytes, dictionary or None, is_last) the reference would send.

    `writes` is a sequence of byte strings (one per `write()` call); `flushes` is a
    set of write indices after which `flush()` is called.
    """
    needs_dict = NEEDS_DICT[fmt]
    buf = bytearray()
    dictionary = None
    msgs = []

    def flush_last(is_last):
        nonlocal dictionary
        while True:
            k = min(len(buf), buffer_size)
            b = bytes(buf[:k]); del buf[:k]
            last = is_last and len(buf) == 0
            msgs.append((b, dictionary, last)); dictionary = None
            if len(b) >= DICT_SIZE and not last and needs_dict:
                dictionary = b[-DICT_SIZE:]
            if len(buf) == 0:
                break

    for i, w in enumerate(writes):
        buf.extend(w)
        while len(buf) > buffer_size:
            b = bytes(buf[:buffer_size]); del buf[:buffer_size]
            msgs.append((b, dictionary, False))
            dictionary = b[-DICT_SIZE:] if needs_dict else None
        if i in flushes:
            flush_last(False)
    flush_last(True)
    return msgs


def compress_stream(fmt, level, buffer_size, writes, flushes=()):
    """Whole-stream oracle: header + encoded blocks in order + footer."""
    out = bytearray(header(fmt, level))
    total = 1 if fmt == ZLIB else 0
    amount = 0
    for blk, d, last in chunk_stream(fmt, buffer_size, writes, flushes):
        out += encode_block(fmt, level, blk, d, last)
        if fmt == GZIP:
            total = lib().oracle_crc32_combine(total, crc32(blk), len(blk))
            amount = (amount + len(blk)) & 0xFFFFFFFF
        elif fmt == ZLIB:
            total = lib().oracle_adler32_combine(total, lib().oracle_adler32(1, blk, len(blk)), len(blk))
            amount = (amount + len(blk)) & 0xFFFFFFFF
    out += footer(fmt, total, amount)
    return bytes(out)


def inflate(raw, out_cap):
    """Raw DEFLATE -> bytes (decode_block restated); raises ValueError(code) on corrupt input."""
    out = ctypes.create_string_buffer(max(out_cap, 1))
    r = lib().oracle_inflate(bytes(raw), len(raw), out, out_cap)
    if r < 0:
        raise ValueError(int(r))
    return out.raw[:r]


def decode_stream(fmt, data, out_cap=None):
    """ParDecompress over concatenated Bgzf / Mgzip members: (status, bytes, found_crc, expected_crc)."""
    data = bytes(data)
    if out_cap is None:
        out_cap = 64 + 1032 * len(data)
    out = ctypes.create_string_buffer(max(out_cap, 1))
    olen = ctypes.c_size_t(0)
    found, expected = ctypes.c_uint32(0), ctypes.c_uint32(0)
    rc = lib().oracle_decode_stream(fmt, data, len(data), out, out_cap, ctypes.byref(olen), ctypes.byref(found), ctypes.byref(expected))
    return rc, out.raw[:olen.value], found.value, expected.value
