"""Per-kernel device times of one kernel-set configuration (not a test; run under gpurun).  The configuration comes
from the environment (GZPB_LIB = another build of the library, GZPB_GATHER_CTAS, GZPB_SEPARATE_CHECK, ...), one process per configuration, so that A/B loops are
plain shell loops.  Prints one JSON line: total ms per batch, per-kernel ms, sha1 of the packed stream.
usage: python tests/perf_kernels.py [blocks] [level] [steps] [label]"""
import ctypes as C
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gzp_b200  # noqa: E402
from gzp_b200 import _lib, synth  # noqa: E402

BLOCK = 65280


def main():
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 3256
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    label = sys.argv[4] if len(sys.argv) > 4 else ""
    L = _lib.load()
    dev = torch.device("cuda", 0)
    data = synth.corpus_stream(nblk * BLOCK)
    flat = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    d_in = torch.zeros((nblk, 65600), dtype=torch.uint8, device=dev)
    d_in[:, :BLOCK] = flat.view(nblk, BLOCK)
    d_len = torch.full((nblk,), BLOCK, dtype=torch.int32, device=dev)
    d_flags = torch.zeros((nblk,), dtype=torch.int32, device=dev)
    d_off = torch.zeros((nblk + 1,), dtype=torch.int64, device=dev)
    d_status = torch.zeros((nblk,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ctx = gzp_b200.Context(gzp_b200.BGZF, level, device=0, max_block_bytes=BLOCK, max_blocks_in_flight=min(nblk, int(os.environ.get("GZPB_PERF_INFLIGHT", "3256"))))
    d_packed = torch.zeros((nblk * 73728,), dtype=torch.uint8, device=dev)

    def step():
        rc = L.gzpb_encode_device(ctx._h, d_in.data_ptr(), d_len.data_ptr(), d_flags.data_ptr(), nblk, d_packed.data_ptr(),
                                  d_off.data_ptr(), d_status.data_ptr(), st.cuda_stream)
        assert rc == 0

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    status_ok = int(d_status.abs().max().item()) == 0
    total = int(d_off[nblk].item())
    sha = hashlib.sha1(d_packed[:total].cpu().numpy().tobytes()).hexdigest()[:16]
    ctx.set_profiling(True)
    ms = []
    for _ in range(steps):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); step(); e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    kms = {k: ctx.kernel_ms(k) for k in ("crc", "chain", "match", "emit", "gather")}
    ctx.close()
    best = min(ms)
    env = {k: v for k, v in os.environ.items() if k.startswith("GZPB_")}
    print(json.dumps({"label": label, "env": env, "variant": L.gzpb_ctx_variant(ctx._h).decode() if False else None, "level": level, "blocks": nblk,
                      "ms_best": round(best, 3), "GiB/s": round(nblk * BLOCK / (best / 1e3) / (1 << 30), 3), "out_bytes": total, "status_ok": status_ok, "sha1": sha,
                      "kernel_ms": {k: round(v[0] / max(v[1], 1), 3) for k, v in kms.items()}}), flush=True)


if __name__ == "__main__":
    main()
