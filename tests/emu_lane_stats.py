"""Lock-step lane utilisation of the chain walk in k_match (linked chains, order by the k_link estimate) vs
k_match2 (hash groups, order by the exact node count), estimated on the CPU SIMT emulator (not a test).
A warp walks 32 positions in lock step; its cost is the LONGEST walk among them.  utilisation =
sum(nodes visited) / sum over warps of 32 * max(nodes visited in the warp)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import emu  # noqa: E402
import oracle  # noqa: E402
from gzp_b200 import synth  # noqa: E402


def run(v2, level, data):
    if v2:
        os.environ["GZPB_MATCH_V2"] = "1"
    else:
        os.environ.pop("GZPB_MATCH_V2", None)
    L = emu.lib()
    ctx = emu.EmuContext(oracle.BGZF, level, max_block_bytes=65280)
    ctx.encode_stream(data, 65280)
    ctx.close()
    npos = C.c_uint32.in_dll(L, "gzpb_emu_npos").value
    vis = (C.c_uint16 * 65536).in_dll(L, "gzpb_emu_visited")
    order = (C.c_uint16 * 65536).in_dll(L, "gzpb_emu_order")
    total, cost = 0, 0
    # thread t handles order[t], order[t + 1024], ...: warp w in round r covers order[r*1024 + 32 w .. +32)
    for base in range(0, npos, 32):
        vs = [vis[order[i]] for i in range(base, min(base + 32, npos))]
        total += sum(vs)
        cost += 32 * max(vs)
    return total, cost, npos


if __name__ == "__main__":
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    data = synth.text_stream(65280 * 3)[65280 * 2:]          # one full block from the middle of the stream
    for v2 in (False, True):
        total, cost, npos = run(v2, level, data)
        print("%-28s level %d: %d positions, %d nodes visited (%.1f / position), lock-step lane utilisation %.1f %%"
              % ("k_match2 (hash groups)" if v2 else "k_match (linked chains)", level, npos, total, total / npos, 100.0 * total / cost))
