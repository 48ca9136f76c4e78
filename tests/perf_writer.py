"""Throughput of the incremental writer paths on the GPU (not a test; run under gpurun):
  write      gzpb_writer_write in 64 KiB pieces (benches/bench.rs:121 of the reference writes 64 KiB chunks)
  reserve    gzpb_writer_reserve / gzpb_writer_commit, 1 MiB produced in place per call
  file       gzpb_compress_file from a tmpfs file to a tmpfs file
usage: python tests/perf_writer.py [blocks] [batch_blocks] [ndevices]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gzp_b200 import _lib, synth, BGZF  # noqa: E402

BLOCK = 65280


def main():
    nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 16280
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 3256
    ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    L = _lib.load()
    data = synth.text_stream(nblk * BLOCK)
    n = len(data)
    devs = (C.c_int * ndev)(*range(ndev))
    total = [0]

    @_lib.SINK_FN
    def sink(_u, _p, k):
        total[0] += k
        return 0

    res = {"bytes": n, "batch_blocks": batch, "devices": ndev}
    buf = C.create_string_buffer(data, n)
    base = C.addressof(buf)
    for mode in ("write", "reserve", "write", "reserve"):
        h = C.c_void_p()
        rc = L.gzpb_writer_create_multi(C.byref(h), devs, ndev, BGZF, 6, BLOCK, batch, C.cast(sink, C.c_void_p), None)
        assert rc == 0, L.gzpb_strerror(rc)
        total[0] = 0
        t0 = time.perf_counter()
        if mode == "write":
            for off in range(0, n, 65536):
                rc = L.gzpb_writer_write(h, base + off, min(65536, n - off))
                assert rc == 0
        else:
            off = 0
            p, room = C.c_void_p(0), C.c_size_t(0)
            while off < n:
                L.gzpb_writer_reserve(h, C.byref(p), C.byref(room))
                k = min(room.value, n - off, 1 << 20)
                C.memmove(p, base + off, k)
                assert L.gzpb_writer_commit(h, k) == 0
                off += k
        assert L.gzpb_writer_finish(h) == 0
        dt = time.perf_counter() - t0
        L.gzpb_writer_destroy(h)
        res[mode] = {"GiB/s": n / dt / (1 << 30), "s": dt, "out": total[0]}
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    src, dst = os.path.join(tmp, "gzpb_perf_in.bin"), os.path.join(tmp, "gzpb_perf_out.gz")
    with open(src, "wb") as f:
        f.write(data)
    for _ in range(2):
        bi, bo = C.c_uint64(0), C.c_uint64(0)
        t0 = time.perf_counter()
        rc = L.gzpb_compress_file(devs, ndev, BGZF, 6, BLOCK, batch, src.encode(), dst.encode(), C.byref(bi), C.byref(bo))
        dt = time.perf_counter() - t0
        assert rc == 0 and bi.value == n
        res["file"] = {"GiB/s": n / dt / (1 << 30), "s": dt, "out": bo.value, "includes": "context + slab allocation, read(2), write(2)"}
    os.unlink(src); os.unlink(dst)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
