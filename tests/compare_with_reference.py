#!/usr/bin/env python
"""Ready-to-run parity check against the REAL third-party code the reference calls (SURVEY.md §8c, item iii).

gzp's Bgzf / Mgzip path calls libdeflate 1.24 (`libdeflater`, Cargo.lock:414-430) through
`Compressor::deflate_compress` (/root/reference/src/bgzf.rs:214-216, src/mgzip.rs:201-206).  That library is
not on the build machine of this repo, so the oracle's compressed bytes are pinned only through decoders
("parity unpinned").  On a machine that has libdeflate (libdeflate.so / `libdeflate-gzip`) or
htslib's `bgzip`, this script compresses the same 65 280-byte blocks with it and compares, block by block,
with the oracle's raw DEFLATE payload — the bytes the CUDA path is held bit-exact to.

  python tests/compare_with_reference.py [--level 6] [--blocks 64] [--lib /path/to/libdeflate.so]

Exit code 0 = identical (or nothing to compare with: says so), 1 = a block differs (prints the first one).
Test infrastructure: uses the oracle, never the product library."""
import argparse
import ctypes as C
import ctypes.util
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BLOCK = 65280


def find_libdeflate(explicit=None):
    for cand in [explicit, os.environ.get("LIBDEFLATE_SO"), ctypes.util.find_library("deflate"), "libdeflate.so.0", "libdeflate.so"]:
        if not cand:
            continue
        try:
            lib = C.CDLL(cand)
            lib.libdeflate_alloc_compressor.restype = C.c_void_p
            lib.libdeflate_alloc_compressor.argtypes = [C.c_int]
            lib.libdeflate_deflate_compress.restype = C.c_size_t
            lib.libdeflate_deflate_compress.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
            lib.libdeflate_free_compressor.argtypes = [C.c_void_p]
            return lib, cand
        except (OSError, AttributeError):
            continue
    return None, None


def with_libdeflate(lib, blocks, level):
    comp = lib.libdeflate_alloc_compressor(level)
    out = []
    for b in blocks:
        cap = len(b) + max(128, len(b) // 10) + 64
        dst = C.create_string_buffer(cap)
        n = lib.libdeflate_deflate_compress(comp, b, len(b), dst, cap)
        out.append(dst.raw[:n])
    lib.libdeflate_free_compressor(comp)
    return out


def with_bgzip(exe, data, level):
    """bgzip writes 0xff00-byte blocks = gzp's BGZF_BLOCK_SIZE; returns the raw DEFLATE payload of every block."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "in")
        open(src, "wb").write(data)
        comp = subprocess.run([exe, "-l", str(level), "-c", src], check=True, stdout=subprocess.PIPE).stdout
    out, pos = [], 0
    while pos + 18 <= len(comp):
        bsize = int.from_bytes(comp[pos + 16:pos + 18], "little") + 1
        out.append(comp[pos + 18:pos + bsize - 8])
        pos += bsize
    return out[:-1] if out and len(out[-1]) == 2 else out      # drop the EOF marker block


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--blocks", type=int, default=64)
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    import oracle
    from gzp_b200 import synth
    data = synth.text_stream(args.blocks * BLOCK)
    blocks = [data[i:i + BLOCK] for i in range(0, len(data), BLOCK)]
    # edge blocks the reference's tests use too
    blocks += [b"", b"a", bytes(31), bytes(32), bytes(65280), bytes(range(256)) * 255]
    mine = [oracle.deflate(b, args.level) for b in blocks]
    compared = False
    lib, where = find_libdeflate(args.lib)
    if lib:
        compared = True
        theirs = with_libdeflate(lib, blocks, args.level)
        for i, (a, b) in enumerate(zip(mine, theirs)):
            if a != b:
                print(f"DIFF vs libdeflate ({where}) at block {i}: oracle {len(a)} B, libdeflate {len(b)} B")
                return 1
        print(f"identical to libdeflate ({where}) on {len(blocks)} blocks at level {args.level}")
    exe = shutil.which("bgzip")
    if exe:
        compared = True
        theirs = with_bgzip(exe, data, args.level)
        for i, (a, b) in enumerate(zip(mine, theirs)):
            if a != b:
                print(f"DIFF vs bgzip ({exe}) at block {i}: oracle {len(a)} B, bgzip {len(b)} B (bgzip may be built on zlib rather than libdeflate)")
                return 1
        print(f"identical to bgzip ({exe}) on {len(theirs)} blocks at level {args.level}")
    if not compared:
        print("parity unpinned: neither libdeflate.so nor bgzip found on this machine; nothing compared")
    return 0


if __name__ == "__main__":
    sys.exit(main())
