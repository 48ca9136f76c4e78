"""How much of k_match's work does the parser use?  (analysis, not a test)
The GPU searches EVERY position (the parse-independence that makes the path parallel); libdeflate's
sequential parser only searches the positions it visits.  This script counts, with the oracle, the searches and chain
hops the parser asks for on one text block and compares them with the all-positions totals of the emulator run
(tests/emu_lane_stats.py prints those)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from gzp_b200 import synth  # noqa: E402

if __name__ == "__main__":
    L = oracle.lib()
    L.oracle_search_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
    data = synth.text_stream(65280 * 3)[65280 * 2:]
    for level in (4, 6, 7, 9):
        L.oracle_search_stats(None, None, 1)
        oracle.deflate(data, level)
        s, h = C.c_uint64(0), C.c_uint64(0)
        L.oracle_search_stats(C.byref(s), C.byref(h), 1)
        print("level %d: parser searched %d of %d positions (%.1f %%), %d chain hops (%.1f per searched position)"
              % (level, s.value, len(data), 100.0 * s.value / len(data), h.value, h.value / max(1, s.value)))
