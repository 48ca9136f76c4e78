"""Throughput of every format through gzpb_encode_stream with pinned buffers (not a test)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gzp_b200
from gzp_b200 import _lib, synth
L = _lib.load()
MB = 1 << 20
cases = [("Bgzf L6 65280", gzp_b200.BGZF, 6, 65280, "text", 192), ("Mgzip L6 131072", gzp_b200.MGZIP, 6, 131072, "text", 192),
         ("Gzip L6 131072+dict", gzp_b200.GZIP, 6, 131072, "text", 192), ("Gzip L7 262144+dict fastq", gzp_b200.GZIP, 7, 262144, "fastq", 128),
         ("Snap 131072 low-entropy", gzp_b200.SNAP, 0, 131072, "low", 256), ("Snap 131072 text", gzp_b200.SNAP, 0, 131072, "text", 192)]
only = sys.argv[1:] 
for name, fmt, lvl, bs, kind, mb in cases:
    if only and not any(o.lower() in name.lower() for o in only):
        continue
    n = mb * MB // bs * bs
    data = synth.text_stream(n) if kind == "text" else (synth.low_entropy(n) if kind == "low" else (synth.fastq(8 * MB) * (n // (8 * MB) + 1))[:n])
    ctx = gzp_b200.Context(fmt, lvl, max_block_bytes=bs, max_blocks_in_flight=min(1024, n // bs))
    h_in = L.gzpb_host_alloc(n); cap = n + n // 4 + (1 << 20); h_out = L.gzpb_host_alloc(cap)
    C.memmove(h_in, data, n)
    olen = C.c_size_t(0)
    best = 1e9
    for it in range(4):
        t0 = time.perf_counter()
        rc = L.gzpb_encode_stream(ctx._h, h_in, n, bs, h_out, cap, C.byref(olen))
        dt = time.perf_counter() - t0
        assert rc == 0, L.gzpb_strerror(rc)
        if it:
            best = min(best, dt)
    print("%-28s in %4d MB  ratio %.4f  e2e %.2f GiB/s" % (name, n // MB, olen.value / n, n / best / (1 << 30)), flush=True)
    L.gzpb_host_free(h_in); L.gzpb_host_free(h_out); ctx.close()
