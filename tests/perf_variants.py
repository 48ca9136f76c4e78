"""A/B of the kernel-set variants on the GPU (not a test; run under gpurun): device-resident BGZF level-L encode of
`blocks` blocks with the default kernels (k_split + k_link + k_match) and with GZPB_MATCH_V2=1 (k_split + k_group +
k_link(hash3) + k_match2) and with GZPB_SPARSE=1 (k_split + k_link + k_smatch + k_emit<sparse> + filtered fallback),
per-kernel CUDA-event times, identical output checked.  One JSON line per variant.
usage: python tests/perf_variants.py [blocks] [level] [steps]"""
import ctypes as C
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
    L = _lib.load()
    dev = torch.device("cuda", 0)
    data = synth.text_stream(nblk * BLOCK)
    flat = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    d_in = torch.zeros((nblk, 65600), dtype=torch.uint8, device=dev)
    d_in[:, :BLOCK] = flat.view(nblk, BLOCK)
    d_len = torch.full((nblk,), BLOCK, dtype=torch.int32, device=dev)
    d_flags = torch.zeros((nblk,), dtype=torch.int32, device=dev)
    d_off = torch.zeros((nblk + 1,), dtype=torch.int64, device=dev)
    d_status = torch.zeros((nblk,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    L.gzpb_debug_sparse_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
    for variant, env, chunk in (("split+link+match", None, 0), ("split+link+smatch", "GZPB_SPARSE=1", 128), ("split+link+smatch+replay", "GZPB_SPARSE=2", 128),
                                ("split+group+match2", "GZPB_MATCH_V2=1", 0), ("split+link+match", None, 0), ("split+link+smatch", "GZPB_SPARSE=1", 128),
                                ("split+link+smatch+replay", "GZPB_SPARSE=2", 128), ("split+link+smatch+replay", "GZPB_SPARSE=2", 256),
                                ("split+link+smatch+replay", "GZPB_SPARSE=2", 512)):
        for k in ("GZPB_MATCH_V2", "GZPB_SPARSE", "GZPB_SPARSE_CHUNK"):
            os.environ.pop(k, None)
        if env:
            os.environ[env.split("=")[0]] = env.split("=")[1]
        if chunk:
            os.environ["GZPB_SPARSE_CHUNK"] = str(chunk)
        ctx = gzp_b200.Context(gzp_b200.BGZF, level, device=0, max_block_bytes=BLOCK, max_blocks_in_flight=min(nblk, 3256))
        assert L.gzpb_ctx_variant(ctx._h).decode() == variant
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
        out = d_packed[:total].clone()
        if ref is None:
            assert status_ok
            ref = out
        identical = bool(status_ok and out.numel() == ref.numel() and torch.equal(ref, out))    # recorded, not fatal: measure the rest too
        ctx.set_profiling(True)
        ms = []
        for _ in range(steps):
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); step(); e1.record(st)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        kms = {k: ctx.kernel_ms(k) for k in ("chain", "match", "emit", "gather")}
        ctx.set_profiling(False)
        su, sm = C.c_uint64(0), C.c_uint64(0)
        L.gzpb_debug_sparse_stats(ctx._h, C.byref(su), C.byref(sm), 1)
        ctx.close()
        best = min(ms)
        print(json.dumps({"variant": variant + (" chunk %d" % chunk if chunk else ""), "level": level, "blocks": nblk, "ms_best": best, "ms_all": ms,
                          "GiB/s": nblk * BLOCK / (best / 1e3) / (1 << 30), "out_bytes": total, "identical_to_default": identical, "sparse_units": su.value, "sparse_missed": sm.value,
                          "kernel_ms_per_batch": {k: v[0] / steps for k, v in kms.items()},
                          "kernel_launches_per_batch": {k: v[1] / steps for k, v in kms.items()}}), flush=True)


if __name__ == "__main__":
    main()
