"""Ad-hoc: encode the edge blocks of test_bgzf_blocks_bit_exact one call each (not a test)."""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, gzp_b200
from gzp_b200 import synth, BGZF
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_parity import _edge_blocks
level = 6
t = synth.text(65280 * 2)
blocks = [t[:65280]] + _edge_blocks()
ctx = gzp_b200.Context(BGZF, level, max_blocks_in_flight=4)
for i, b in enumerate(blocks):
    print("block", i, len(b), flush=True)
    enc = ctx.encode_blocks([(b, None, False)])[0][0]
    print("   ", "ok" if enc == oracle.encode_block(oracle.BGZF, level, b) else "MISMATCH", flush=True)
