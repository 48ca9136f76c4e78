"""Throughput of the block DECODE path (not a test): k_inflate device-resident (CUDA events on the
launching stream) and gzpb_decode_stream end to end with pinned host buffers, beside the CPU oracle and
stock zlib on one host core.  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gzp_b200
import oracle
from gzp_b200 import _lib, synth

L = _lib.load()
GIB = float(1 << 30)
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 10656
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
steps = 5
data = synth.text_stream(65280 * nblk)
ctx = gzp_b200.Context(gzp_b200.BGZF, level, max_blocks_in_flight=min(nblk, 2048))
comp = ctx.encode_stream(data)
ctx.close()
n, clen = len(data), len(comp)

dec = gzp_b200.Decoder(gzp_b200.BGZF, max_blocks_in_flight=5328)
nb = C.c_size_t(0); total = C.c_uint64(0)
L.gzpb_scan_blocks(gzp_b200.BGZF, comp, clen, None, 0, C.byref(nb), None, C.byref(total))
descs = (_lib.BlockDesc * nb.value)()
L.gzpb_scan_blocks(gzp_b200.BGZF, comp, clen, descs, nb.value, C.byref(nb), None, None)
assert total.value == n

dev = torch.device("cuda", 0)
d_comp = torch.zeros(clen + 256, dtype=torch.uint8, device=dev)
d_comp[:clen] = torch.frombuffer(bytearray(comp), dtype=torch.uint8).to(dev)
d_desc = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
d_out = torch.empty(n + 256, dtype=torch.uint8, device=dev)
d_status = torch.zeros(nb.value, dtype=torch.int32, device=dev)
d_crc = torch.zeros(nb.value, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def dev_step():
    for b0 in range(0, nb.value, 5328):          # one launch per 5328 members = 148 SMs x 36 resident warps
        k = min(5328, nb.value - b0)
        rc = L.gzpb_decode_device(dec._h, d_comp.data_ptr(), d_desc.data_ptr() + 32 * b0, k, d_out.data_ptr(), d_status.data_ptr() + 4 * b0,
                                  d_crc.data_ptr() + 4 * b0, st.cuda_stream)
        assert rc == 0


for _ in range(3):
    dev_step()
torch.cuda.synchronize()
assert int(d_status.abs().max().item()) == 0
assert bytes(d_out[:n].cpu().numpy().tobytes()) == data, "device decode differs from the input"
ms = []
for _ in range(steps):
    flush.fill_(1)                       # L2 flush between timed iterations (256 MiB > 126 MB L2)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st); dev_step(); e1.record(st)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
dev_ms = sum(ms) / len(ms)

h_in = L.gzpb_host_alloc(clen + 64); h_out = L.gzpb_host_alloc(n + 64)
C.memmove(h_in, comp, clen)
olen = C.c_size_t(0)
best = 1e9
for it in range(4):
    t0 = time.perf_counter()
    rc = L.gzpb_decode_stream(dec._h, h_in, clen, h_out, n + 64, C.byref(olen), None)
    dt = time.perf_counter() - t0
    assert rc == 0, rc
    if it:
        best = min(best, dt)
assert olen.value == n and C.string_at(h_out, n) == data

# CPU, one core: the oracle's inflate and stock zlib on a bounded sample
sample_blocks = min(nb.value, 256)
end = descs[sample_blocks - 1].in_off + descs[sample_blocks - 1].in_len + 8
t0 = time.perf_counter(); rc, out, _, _ = oracle.decode_stream(oracle.BGZF, comp[:end]); t_or = time.perf_counter() - t0
assert rc == 0
t0 = time.perf_counter()
for i in range(sample_blocks):
    zlib.decompressobj(-15).decompress(comp[descs[i].in_off:descs[i].in_off + descs[i].in_len])
t_z = time.perf_counter() - t0
line = {"metric": "bgzf_decode_output_throughput", "unit": "GiB/s", "blocks": nb.value, "level": level, "bytes_out": n, "ratio": clen / n,
        "device": {"value": n / (dev_ms / 1e3) / GIB, "ms_per_pass": dev_ms, "l2": "256 MiB flush between iterations",
                   "algorithmic_bytes": clen + 2 * n, "hbm_gbs": (clen + 2 * n) / (dev_ms / 1e3) / 1e9},
        "e2e": {"value": n / best / GIB, "api": "gzpb_decode_stream (pinned host in/out)", "h2d_bytes": clen, "d2h_bytes": n},
        "cpu_one_core": {"oracle_inflate": len(out) / t_or / GIB, "zlib_inflate": len(out) / t_z / GIB, "sample_blocks": sample_blocks}}
print(json.dumps(line), flush=True)
