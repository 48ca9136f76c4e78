"""Viability of a chunk-speculative parse (analysis for DESIGN.md §6, not a test).
The true level-L parse of one text block is compared with parses that start "cold" at chunk boundaries (everything
before the start is in the hash chains — the oracle's dictionary mode gives exactly that).  For each chunk start we
report how many positions pass until the speculative parse's token boundaries coincide with the true parse's and stay
identical (token for token) to the end of the chunk."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from gzp_b200 import synth  # noqa: E402


class BlockInfo(C.Structure):
    _fields_ = [("begin", C.c_uint32), ("length", C.c_uint32), ("ntokens", C.c_uint32), ("tok_offset", C.c_uint32), ("btype", C.c_int32),
                ("cost_dyn", C.c_uint32), ("cost_static", C.c_uint32), ("cost_stored", C.c_uint32), ("litlen_lens", C.c_uint8 * 288),
                ("offset_lens", C.c_uint8 * 32)]


class Trace(C.Structure):
    _fields_ = [("nblocks", C.c_uint32), ("blocks", BlockInfo * 64), ("tokens", C.POINTER(C.c_uint32)), ("tokens_cap", C.c_size_t),
                ("ntokens", C.c_size_t)]


def parse(L, data, start, level):
    """token list [(pos, len, off)] of the parse of data[start:] with data[:start] as history"""
    cap = len(data) + 16
    toks = (C.c_uint32 * cap)()
    tr = Trace()
    tr.tokens = C.cast(toks, C.POINTER(C.c_uint32)); tr.tokens_cap = cap
    out = C.create_string_buffer(len(data) + 4096)
    n = L.oracle_deflate_ex(data, start, len(data) - start, level, 0, out, len(out), C.byref(tr))
    assert n > 0
    res, pos = [], start
    for i in range(tr.ntokens):
        t = toks[i]
        if t & 0x80000000:
            ln, off = (t >> 16) & 0x7FFF, t & 0xFFFF
            res.append((pos, ln, off)); pos += ln
        else:
            res.append((pos, 1, 0)); pos += 1
    assert pos == len(data), (pos, len(data))
    return res


if __name__ == "__main__":
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    L = oracle.lib()
    L.oracle_deflate_ex.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_char_p, C.c_size_t, C.c_void_p]
    data = synth.text_stream(65280 * 3)[65280 * 2:]
    true = parse(L, data, 0, level)
    true_at = {t[0]: i for i, t in enumerate(true)}
    dists, never = [], 0
    for s in range(chunk, len(data) - chunk, chunk * int(os.environ.get("STRIDE", "8"))):   # every 8th chunk start by default
        spec = parse(L, data, s, level)
        end = s + chunk
        # first token of the speculative parse from which it equals the true parse up to the end of the chunk
        sync = None
        for k, t in enumerate(spec):
            if t[0] >= end:
                break
            j = true_at.get(t[0])
            if j is None:
                continue
            m = 0
            while k + m < len(spec) and spec[k + m][0] < end and j + m < len(true) and spec[k + m] == true[j + m]:
                m += 1
            if k + m >= len(spec) or spec[k + m][0] >= end:
                sync = t[0] - s
                break
        if sync is None:
            never += 1
        else:
            dists.append(sync)
    dists.sort()
    print("level %d, chunks of %d positions, %d chunk starts sampled: synchronised after median %d / mean %.1f / p90 %d / max %d positions; %d never inside the chunk"
          % (level, chunk, len(dists) + never, dists[len(dists) // 2], sum(dists) / max(1, len(dists)), dists[int(len(dists) * 0.9)], dists[-1], never))
