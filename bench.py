#!/usr/bin/env python
"""bench.py — benchmark of the B200-native per-block encode path (gzp's ParCompress worker + writer).

Headline (default `--config bgzf`, BASELINE.json `metric`, configs[1]): compressed GiB/s of INPUT bytes,
ParCompress<Bgzf> level 6, 65 280-byte blocks, on bench-data/shakespeare.txt repeated end to end (the 54.65 GB
"shakespeare x 10000" stream; every step compresses another window of it, see `config.l2`).

A "step" = one pass of the hot path over `--blocks` consecutive blocks per GPU.  One JSON line:
  value      device-resident: inputs already in HBM in unit layout, gzpb_encode_device_ex on the launching stream,
             CUDA-event timed; N > 1: every rank its own window of the stream (independent blocks, weak scaling).
  e2e        ONE ordered stream through the reference-facing C ABI with pinned HOST buffers, H2D + kernels + D2H
             inside the timed region.  N = 1: gzpb_encode_stream.  N > 1: gzpb_encode_stream_multi — rank 0 deals the
             device batches of a stream N times as long round-robin over all N GPUs of the box and one offset chain
             puts the blocks back in stream order (src/par/compress.rs:303-313); ranks 1..N-1 wait at the barrier.
  roofline   dominant kernel, algorithmic bytes / CUDA-event time, against MEASURED_PEAKS.json.
  cpu_baseline  the oracle's ParCompress port on the host cores (bounded sample; kind "port").

Other BASELINE configs: `--config mgzip | snap | gzip9` (configs[2..4]), same JSON shape.
`--impl reference` times the reference's CPU path (oracle port of its thread topology) on the host cores.
`--full-stream` pushes the whole 54 653 940 000-byte stream once through the incremental writer on `--gpus` GPUs
of this process (not under torchrun) and prints its own JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = float(1 << 30)
UNIT = "GiB/s"
DICT = 32768
# formats (include/gzpb.h)
GZIP, ZLIB, RAWDEFLATE, MGZIP, BGZF, SNAP = 0, 1, 2, 3, 4, 5

CONFIGS = {
    # blocks: per step and GPU; inflight: blocks per device batch
    "bgzf": dict(fmt=BGZF, level=6, block=65280, inflight=4736, blocks=47360, data="corpus", metric="bgzf_l6_compress_input_throughput",
                 workload="ParCompress<Bgzf> level 6, 65280-B blocks, shakespeare.txt repeated (BASELINE configs[1]: windows of the 54.65 GB stream)"),
    "mgzip": dict(fmt=MGZIP, level=6, block=131072, inflight=3256, blocks=16280, data="corpus", metric="mgzip_l6_compress_input_throughput",
                  workload="ParCompress<Mgzip> level 6, 131072-B blocks, shakespeare.txt repeated (BASELINE configs[2])"),
    "snap": dict(fmt=SNAP, level=0, block=131072, inflight=2048, blocks=8192, data="low", metric="snap_compress_input_throughput",
                 workload="ParCompress<Snap>, 131072-B blocks, low-entropy synthetic binary (BASELINE configs[3]; SURVEY 8d generator, 256 MiB period)"),
    "gzip9": dict(fmt=GZIP, level=9, block=262144, inflight=2368, blocks=7104, data="fastq", metric="gzip_l9_dict_compress_input_throughput",
                  workload="ParCompress<Gzip> level 9 (dictionary carry), 262144-B blocks, FASTQ-shaped synthetic (BASELINE configs[4]; 64 MiB period)"),
}
# Algorithmic HBM bytes per input byte of each config's dominant kernel:
#   k_match: the input read once + 4 bytes per token written, T ~ 0.30 tokens per byte on text -> 2.2
#   k_snap : the input read once + r written
ALGO_BYTES = {"k_match": 2.2, "k_snap": None}

_T0 = time.perf_counter()


def log(msg):
    """Stage progress on stderr (stdout carries only the JSON line)."""
    print("[bench %7.2fs] %s" % (time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="bgzf", choices=sorted(CONFIGS))
    ap.add_argument("--blocks", type=int, default=0, help="gzp blocks per step and GPU (default per config: ten device batches for the deflate-family text configs, 2.1 GB)")
    ap.add_argument("--inflight", type=int, default=0, help="blocks per device batch (default per config; bgzf: 4736 = 148 SMs x 32 resident k_emit CTAs)")
    ap.add_argument("--cpu-sample-mb", type=float, default=0.0, help="override the CPU baseline sample size")
    ap.add_argument("--full-stream", action="store_true", help="the whole shakespeare x 10000 stream once through the incremental writer")
    ap.add_argument("--copy-threads", type=int, default=0, help="--full-stream: helper threads for the writer's host copy (default: all cores)")
    ap.add_argument("--feed", default="write", choices=["write", "reserve"],
                    help="--full-stream: `write` = gzpb_writer_write from a caller buffer (one host copy, made by --copy-threads threads); "
                         "`reserve` = producer threads generate the stream in place in the writer's pinned slab (gzpb_writer_reserve / _commit)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------
# workload data: a repeating pattern per config (the corpus itself, or a long synthetic period)
# ---------------------------------------------------------------------------------------------------------
def pattern_for(cfg):
    from gzp_b200 import synth
    if cfg["data"] == "corpus":
        return synth.corpus()
    cache = "/tmp/gzpb_bench_%s.bin" % cfg["data"]
    size = (256 << 20) if cfg["data"] == "low" else (64 << 20)
    if os.path.exists(cache) and os.path.getsize(cache) == size:
        return open(cache, "rb").read()
    data = synth.low_entropy(size) if cfg["data"] == "low" else synth.fastq(size)
    try:
        tmp = "%s.%d.tmp" % (cache, os.getpid())
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, cache)
    except OSError:
        pass
    return data


def window(pattern, start, nbytes):
    """`nbytes` of the pattern repeated end to end, from stream offset `start`."""
    P = len(pattern)
    start %= P
    head = pattern[start:start + nbytes]
    rest = nbytes - len(head)
    return head if rest == 0 else b"".join([head, pattern * (rest // P), pattern[:rest % P]])


def fill_pinned(ptr, pattern, start, nbytes):
    """The same bytes written into pinned memory at `ptr` piece by piece (no multi-GB Python object)."""
    P = len(pattern)
    buf = (C.c_char * P).from_buffer_copy(pattern)
    base = C.addressof(buf)
    pos, s = 0, start % P
    while pos < nbytes:
        k = min(P - s, nbytes - pos)
        C.memmove(ptr + pos, base + s, k)
        pos += k
        s = 0


# ---------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of gzp's ParCompress topology; checker infrastructure timed as the baseline)
# ---------------------------------------------------------------------------------------------------------
def best_thread_count(L, cfg, data, threads):
    """gzp's default is one worker per logical CPU (`num_cpus::get()`, par/compress.rs:57); where that oversubscribes
    the memory system a smaller pool is faster, so the baseline takes the best of T, T/2, T/4 on a short probe."""
    bs = cfg["block"]
    n = min(len(data), 48 * 65280 * threads) // bs * bs
    out = C.create_string_buffer(n + n // 8 + (4 << 20))
    olen = C.c_size_t(0)
    best, best_rate = threads, 0.0
    for th in sorted({threads, max(1, threads // 2), max(1, threads // 4)}, reverse=True):
        t = min(L.oracle_par_compress(cfg["fmt"], cfg["level"], bs, th, data, n, out, len(out), C.byref(olen)) for _ in range(2))
        if t > 0 and n / t > best_rate * 1.10:
            best, best_rate = th, n / t
        log("cpu probe: %d threads -> %.3f GiB/s" % (th, n / max(t, 1e-9) / GIB))
    return best


def cpu_baseline(cfg, data, threads, sample_mb=0.0):
    """The oracle's ParCompress port (kind 'port': the reference cannot be built here, no Rust toolchain / libdeflate
    source) on a bounded sample of the same workload: whole-stream passes repeated for >= ~10 s of CPU work."""
    import oracle
    L = oracle.lib()
    bs = cfg["block"]
    threads = best_thread_count(L, cfg, data, threads)
    n = min(len(data), int(sample_mb * 1e6) if sample_mb else 512 << 20) // bs * bs
    sample = data[:n]
    out = C.create_string_buffer(n + n // 8 + (4 << 20))
    olen = C.c_size_t(0)
    budget = 10.0 if not sample_mb else 0.0
    total_t, passes = 0.0, 0
    while True:
        t = L.oracle_par_compress(cfg["fmt"], cfg["level"], bs, threads, sample, n, out, len(out), C.byref(olen))
        if t <= 0:
            raise RuntimeError("oracle_par_compress failed")
        total_t += t; passes += 1
        if total_t >= budget or passes >= 64:
            break
    val = n * passes / total_t / GIB
    return {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{passes} pass(es) over {n} B of the same stream ({n // bs} blocks), ratio {olen.value / n:.4f}, {total_t:.2f} s"}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port, all host threads)."""
    if rank != 0:
        return
    import oracle
    L = oracle.lib()
    bs = cfg["block"]
    pattern = pattern_for(cfg)
    threads = best_thread_count(L, cfg, window(pattern, 0, 48 * 65280 * host_threads()), host_threads())
    probe = window(pattern, 0, max(8 * 65280 * threads // bs, 2) * bs)
    out = C.create_string_buffer(len(probe) + (64 << 20))
    olen = C.c_size_t(0)
    t = L.oracle_par_compress(cfg["fmt"], cfg["level"], bs, threads, probe, len(probe), out, len(out), C.byref(olen))
    rate = len(probe) / max(t, 1e-6)
    nblocks = max(threads, min(args.blocks or cfg["blocks"], int(rate * 3.0) // bs))      # ~3 s of CPU work per step
    nbytes = nblocks * bs
    shift = 16 * bs
    data = window(pattern, 0, nbytes + shift * (args.steps + args.warmup))
    out = C.create_string_buffer(nbytes + nbytes // 8 + (4 << 20))
    times = []
    for i in range(args.warmup + args.steps):
        sample = data[i * shift: i * shift + nbytes]
        t = L.oracle_par_compress(cfg["fmt"], cfg["level"], bs, threads, sample, nbytes, out, len(out), C.byref(olen))
        if i >= args.warmup:
            times.append(t)
    total = sum(times)
    val = nbytes * len(times) / total / GIB
    line = {"impl": "reference", "metric": cfg["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["workload"], "sample_blocks_per_step": nblocks, "host_threads": threads},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{nblocks} blocks ({nbytes} B) per step: oracle port of gzp's ParCompress topology + libdeflate-style encoder (reference not buildable here: no Rust toolchain)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# checks of what was timed (outside the timed regions)
# ---------------------------------------------------------------------------------------------------------
def decode_check(fmt, stream_bytes, want):
    """The stream decodes to its input with a stock decoder (gzip members / zlib stream), or — Snap, whose framed
    format no stock library here reads — with the oracle's frame reader over pyarrow-verified raw blocks."""
    if fmt in (BGZF, MGZIP, GZIP):
        import gzip
        import io
        got = gzip.GzipFile(fileobj=io.BytesIO(stream_bytes)).read()
    else:
        import pyarrow as pa
        codec = pa.Codec("snappy")                               # stock raw-Snappy decoder; the framing is parsed here
        got, pos = bytearray(), 0
        while pos < len(stream_bytes):
            typ = stream_bytes[pos]
            ln = int.from_bytes(stream_bytes[pos + 1:pos + 4], "little")
            body = stream_bytes[pos + 4:pos + 4 + ln]
            pos += 4 + ln
            if typ == 0xFF:                                      # stream identifier: every gzp block starts one (snap.rs:61-74)
                assert body == b"sNaPpY"
            elif typ == 0x01:
                got += body[4:]
            elif typ == 0x00:
                raw, n, sh, i = body[4:], 0, 0, 0
                while True:
                    n |= (raw[i] & 0x7F) << sh; sh += 7; i += 1
                    if raw[i - 1] < 0x80:
                        break
                got += codec.decompress(raw, decompressed_size=n).to_pybytes()
            else:
                return False
        got = bytes(got)
    return got == want


def main():
    args = parse_args()
    cfg = dict(CONFIGS[args.config])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if args.full_stream:
        full_stream(args, cfg)
        return

    import faulthandler
    faulthandler.enable()
    if os.environ.get("GZPB_BENCH_WATCHDOG"):
        faulthandler.dump_traceback_later(int(os.environ["GZPB_BENCH_WATCHDOG"]), repeat=True)
    log("importing torch")
    import torch
    import torch.distributed as dist
    import gzp_b200
    from gzp_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")      # host-side barrier for the ranks that idle through rank 0's one-stream leg

    log("cuda ready")
    L = _lib.load()
    fmt, level, bs = cfg["fmt"], cfg["level"], cfg["block"]
    nblk = args.blocks or cfg["blocks"]
    inflight = min(nblk, args.inflight or cfg["inflight"])
    step_bytes = nblk * bs
    dict_len = DICT if fmt in (GZIP, ZLIB, RAWDEFLATE) else 0
    nwin = 4                                             # distinct stream windows rotated over the steps
    shift = 977 * 65280 // bs * bs                       # window start moves by this much per step (whole blocks)
    pattern = pattern_for(cfg)
    rank_off = DICT + rank * 131 * bs                    # every rank compresses a different part of the stream
    stream = window(pattern, rank_off - dict_len, dict_len + step_bytes + nwin * shift)
    log("stream window ready (%d B)" % len(stream))

    ctx = gzp_b200.Context(fmt, level, device=local_rank, max_block_bytes=bs, max_blocks_in_flight=inflight)
    stride = L.gzpb_unit_stride(ctx._h)
    cpu_per_unit = (bs + 65535) // 65536 if fmt == SNAP else 1
    log("context created (unit stride %d)" % stride)

    # ---------------- device-resident arm ----------------
    host_all = torch.frombuffer(bytearray(stream), dtype=torch.uint8)
    d_ins = []
    for w in range(nwin):
        lo = dict_len + w * shift
        flat = host_all[lo - dict_len: lo + step_bytes].to(dev)
        d_in = torch.zeros((nblk, stride), dtype=torch.uint8, device=dev)
        d_in[:, dict_len:dict_len + bs] = flat[dict_len:].view(nblk, bs)
        if dict_len:                                     # every unit carries the 32 KiB in front of it (par/compress.rs:419-423)
            d_in[0, :dict_len] = flat[:dict_len]
            d_in[1:, :dict_len] = flat[dict_len:].view(nblk, bs)[:-1, bs - dict_len:]
        d_ins.append(d_in)
        del flat
    d_len = torch.full((nblk,), dict_len + bs, dtype=torch.int32, device=dev)
    d_dict = torch.full((nblk,), dict_len, dtype=torch.int32, device=dev) if dict_len else None
    d_flags = torch.full((nblk,), 2 if fmt in (GZIP, ZLIB, RAWDEFLATE) else 0, dtype=torch.int32, device=dev)
    out_cap_unit = L.gzpb_encode_capacity(fmt, bs) + 64
    d_packed = torch.empty((nblk * out_cap_unit,), dtype=torch.uint8, device=dev)
    d_off = torch.zeros((nblk * cpu_per_unit + 1,), dtype=torch.int64, device=dev)
    d_status = torch.zeros((nblk,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()

    def dev_step(i):
        rc = L.gzpb_encode_device_ex(ctx._h, d_ins[i % nwin].data_ptr(), d_len.data_ptr(), d_dict.data_ptr() if dict_len else None,
                                     d_flags.data_ptr(), nblk, d_packed.data_ptr(), d_off.data_ptr(), d_status.data_ptr(), st.cuda_stream)
        if rc != 0:
            raise RuntimeError("gzpb_encode_device_ex: " + L.gzpb_strerror(rc).decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log("device inputs staged; warm-up")
    for i in range(args.warmup):
        dev_step(i)
    torch.cuda.synchronize()
    log("device warm-up done")
    assert int(d_status.abs().max().item()) == 0
    out_bytes_dev = int(d_off[nblk * cpu_per_unit].item())

    # a sample of the device arm's output against the oracle (checker only, outside the timed region)
    if rank == 0:
        import oracle
        w = (args.warmup - 1) % nwin if args.warmup else None
        if w is not None:
            offs = d_off[: 4 * cpu_per_unit + 1].cpu().tolist()
            got = bytes(d_packed[: offs[-1]].cpu().numpy())
            lo = dict_len + w * shift
            want = b""
            for b in range(4):
                blk = stream[lo + b * bs: lo + (b + 1) * bs]
                dic = stream[lo + b * bs - dict_len: lo + b * bs] if dict_len else None
                want += oracle.encode_block(fmt, level, blk, dic, False)
            assert got == want, "device arm: the first four blocks differ from the oracle"
            log("device arm sample identical to the oracle (4 blocks)")

    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(args.steps):
        dev_step(args.warmup + i)
    e1.record(st)
    barrier()
    clocks = sampler.stop()
    dev_ms = e0.elapsed_time(e1)
    log("device arm timed: %.1f ms for %d steps" % (dev_ms, args.steps))
    launches = ctx.launch_count() - launches0
    knames = ("snap", "gather") if fmt == SNAP else ("crc", "chain", "match", "emit", "gather")
    kms = {k: ctx.kernel_ms(k) for k in knames}
    ctx.set_profiling(False)
    if world > 1:
        t = torch.tensor([dev_ms, float(launches)], device=dev)
        dist.all_reduce(t[0:1], op=dist.ReduceOp.MAX); dist.all_reduce(t[1:2], op=dist.ReduceOp.SUM)
        dev_ms, launches = float(t[0].item()), int(t[1].item())
    del d_ins, d_packed
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm: ONE ordered stream, host buffers through the C ABI ----------------
    e2e = None
    e2e_bytes = step_bytes * world                       # the stream of a step: N times as long on N GPUs
    if rank == 0:
        ctxs = [ctx]
        for d in range(1, world):
            ctxs.append(gzp_b200.Context(fmt, level, device=d, max_block_bytes=bs, max_blocks_in_flight=inflight))
        harr = (C.c_void_p * world)(*[c._h for c in ctxs])
        in_bytes = e2e_bytes + nwin * shift
        h_in = L.gzpb_host_alloc(in_bytes)
        out_cap = e2e_bytes + e2e_bytes // 8 + (8 << 20) if fmt == SNAP else e2e_bytes // 2 + e2e_bytes // 8 + (8 << 20)
        h_out = L.gzpb_host_alloc(out_cap)
        if not h_in or not h_out:
            raise RuntimeError("gzpb_host_alloc failed")
        fill_pinned(h_in, pattern, rank_off, in_bytes)
        olen = C.c_size_t(0)

        def e2e_step(i):
            if world == 1:
                rc = L.gzpb_encode_stream(ctx._h, h_in + (i % nwin) * shift, e2e_bytes, bs, h_out, out_cap, C.byref(olen))
            else:
                rc = L.gzpb_encode_stream_multi(harr, world, h_in + (i % nwin) * shift, e2e_bytes, bs, h_out, out_cap, C.byref(olen))
            if rc != 0:
                raise RuntimeError("gzpb_encode_stream: " + L.gzpb_strerror(rc).decode())

        log("e2e buffers pinned (%d + %d B); warm-up" % (in_bytes, out_cap))
        for i in range(args.warmup):
            e2e_step(i)
        log("e2e warm-up done")
        for d in range(world):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(args.warmup + i)
        for d in range(world):
            torch.cuda.synchronize(d)
        e2e_s = time.perf_counter() - t0
        log("e2e arm timed: %.1f ms for %d steps" % (e2e_s * 1e3, args.steps))
        out_bytes = olen.value
        last = (args.warmup + args.steps - 1) % nwin
        # what was timed is right: (1) on N > 1 GPUs the stream is byte-identical to the same input through ONE GPU;
        # (2) a stock decoder turns (a bounded prefix of) the stream back into the input
        same_as_one_gpu = None
        if world > 1:
            h_one = L.gzpb_host_alloc(out_cap)            # the same input through ONE GPU, compared in place (streams exceed 2 GiB)
            olen1 = C.c_size_t(0)
            rc = L.gzpb_encode_stream(ctx._h, h_in + last * shift, e2e_bytes, bs, h_one, out_cap, C.byref(olen1))
            assert rc == 0, L.gzpb_strerror(rc)
            libc = C.CDLL(None)
            libc.memcmp.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
            libc.memcmp.restype = C.c_int
            same_as_one_gpu = (olen1.value == out_bytes and libc.memcmp(h_out, h_one, out_bytes) == 0)
            L.gzpb_host_free(h_one)
            assert same_as_one_gpu, "the N-GPU stream differs from the one-GPU stream"
            log("N-GPU stream identical to the one-GPU stream (%d B)" % out_bytes)
        want_all_len = e2e_bytes if e2e_bytes <= (1200 << 20) else None
        if want_all_len:
            ok = decode_check(fmt, C.string_at(h_out, out_bytes), window(pattern, rank_off + last * shift, e2e_bytes))
            checked = e2e_bytes
        else:
            # a prefix of whole blocks (BGZF / Mgzip members are self-contained)
            nb = (256 << 20) // bs
            pre_in = window(pattern, rank_off + last * shift, nb * bs)
            ok, checked = True, 0
            if fmt in (BGZF, MGZIP):
                import zlib
                raw = C.string_at(h_out, min(out_bytes, 200 << 20))
                pos, got = 0, []
                hdr = 18 if fmt == BGZF else 20
                while len(got) < nb and pos + hdr < len(raw):
                    size = (int.from_bytes(raw[pos + 16:pos + 18], "little") + 1) if fmt == BGZF else int.from_bytes(raw[pos + 16:pos + 20], "little")
                    if pos + size > len(raw):
                        break
                    got.append(zlib.decompress(raw[pos + hdr: pos + size - 8], -15))
                    pos += size
                joined = b"".join(got)
                ok = joined == pre_in[:len(joined)] and len(joined) > 0
                checked = len(joined)
            elif fmt in (GZIP, ZLIB, RAWDEFLATE):
                # one continuous stream (dictionary carry): a streaming stock decoder over a prefix of it
                import zlib
                d = zlib.decompressobj(31 if fmt == GZIP else 15 if fmt == ZLIB else -15)
                joined = d.decompress(C.string_at(h_out, min(out_bytes, 200 << 20)), len(pre_in))
                ok = joined == pre_in[:len(joined)] and len(joined) > 0
                checked = len(joined)
        assert ok, "e2e output does not decode to the input"
        log("e2e output verified with the stock decoder (%d B of input)" % checked)
        e2e = {"value": e2e_bytes * args.steps / e2e_s / GIB, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": out_bytes,
               "ms_per_step": 1e3 * e2e_s / args.steps,
               "api": "gzpb_encode_stream (pinned host in/out)" if world == 1 else
                      "gzpb_encode_stream_multi: ONE ordered stream of %d B per step, device batches of %d blocks dealt round-robin over %d GPUs by rank 0, pinned host in/out" % (e2e_bytes, inflight, world),
               "one_ordered_stream": True, "identical_to_one_gpu_stream": same_as_one_gpu, "decoded_input_bytes_checked": checked}
        for c in ctxs[1:]:
            c.close()
        L.gzpb_host_free(h_in); L.gzpb_host_free(h_out)
    if world > 1:
        dist.barrier(group=cpu_group)                    # ranks 1..N-1 wait here on the host while rank 0 drives all GPUs

    # ---------------- the incremental writer (ParCompress::write in 64 KiB pieces, benches/bench.rs:121) ----------------
    writer = None
    if rank == 0 and fmt == BGZF:
        try:
            total_out = [0]

            @_lib.SINK_FN
            def _sink(_u, _p, k):
                total_out[0] += k
                return 0

            wh = C.c_void_p()
            rc = L.gzpb_writer_create(C.byref(wh), local_rank, fmt, level, bs, inflight, C.cast(_sink, C.c_void_p), None)
            if rc != 0:
                raise RuntimeError("gzpb_writer_create: " + L.gzpb_strerror(rc).decode())
            src = C.create_string_buffer(stream[:step_bytes], step_bytes)
            sbase = C.addressof(src)
            t0 = time.perf_counter()
            for off in range(0, step_bytes, 65536):
                rc = L.gzpb_writer_write(wh, sbase + off, min(65536, step_bytes - off))
                if rc != 0:
                    raise RuntimeError("gzpb_writer_write: " + L.gzpb_strerror(rc).decode())
            rc = L.gzpb_writer_finish(wh)
            w_s = time.perf_counter() - t0
            L.gzpb_writer_destroy(wh)
            if rc != 0:
                raise RuntimeError("gzpb_writer_finish: " + L.gzpb_strerror(rc).decode())
            del src
            writer = {"value_per_gpu": step_bytes / w_s / GIB, "unit": UNIT, "out_bytes": total_out[0],
                      "api": "gzpb_writer_write in 64 KiB pieces from pageable caller memory + gzpb_writer_finish (one pass of one step, pipeline fill included; one GPU)"}
            log("incremental writer timed: %.1f ms" % (w_s * 1e3))
        except Exception as e:                                   # noqa: BLE001 - informational leg, never fatal
            log("incremental writer leg skipped: %r" % (e,))

    total_in = step_bytes * args.steps * world
    value = total_in / (dev_ms / 1e3) / GIB

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        dom = "snap" if fmt == SNAP else "match"
        dom_kernel = "k_snap" if fmt == SNAP else "k_match"
        dom_ms, dom_n = kms[dom]
        ratio = out_bytes_dev / step_bytes
        algo = ALGO_BYTES[dom_kernel] if ALGO_BYTES[dom_kernel] is not None else 1.0 + ratio
        units_per_launch = inflight
        per_launch_bytes = algo * bs * units_per_launch
        achieved = per_launch_bytes / (dom_ms / max(dom_n, 1) / 1e3) / 1e9 if dom_n else 0.0
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            traffic = tr[args.config][dom_kernel]["dram_bytes_per_input_byte"] * bs * units_per_launch / 1e9
        except Exception:
            pass
        log("timing the CPU baseline")
        cpu = cpu_baseline(cfg, stream[dict_len:], host_threads(), args.cpu_sample_mb)
        log("cpu baseline done")
        share = {k: round(v[0] / max(sum(x[0] for x in kms.values()), 1e-9), 4) for k, v in kms.items()}
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["workload"], "name": args.config,
                       "blocks_per_step_per_gpu": nblk, "bytes_per_step_per_gpu": step_bytes, "blocks_per_device_batch": inflight, "ratio": ratio,
                       "l2": "each step's input (%.0f MB per GPU) exceeds L2 and rotates over %d stream windows" % (step_bytes / 1e6, nwin),
                       "parallelism": f"value: independent blocks sharded over {world} GPU(s), no data collective; e2e: one ordered stream dealt over {world} GPU(s)"},
            "e2e": dict(e2e, incremental_writer=writer),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write, profiles/r2_traffic.json)",
                         "algorithmic_bytes_per_input_byte": algo, "units_per_launch": units_per_launch,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                         "note": "this path is not HBM-bound: k_match is bound by shared-memory wavefronts and instruction issue (profiles/), the HBM fraction is reported as asked",
                         "kernel_ms_per_launch": {k: v[0] / max(v[1], 1) for k, v in kms.items()}, "kernel_time_share": share},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# --full-stream: the whole 54 653 940 000-byte stream once through the incremental writer (one process, N GPUs)
# ---------------------------------------------------------------------------------------------------------
def full_stream(args, cfg):
    import zlib
    import gzp_b200  # noqa: F401
    from gzp_b200 import _lib, synth
    L = _lib.load()
    fmt, level, bs = cfg["fmt"], cfg["level"], cfg["block"]
    assert fmt == BGZF, "--full-stream is BASELINE configs[1]"
    ndev = args.gpus
    inflight = args.inflight or cfg["inflight"]
    S = synth.CORPUS_BYTES
    copies = 10000 if not os.environ.get("GZPB_FULL_COPIES") else int(os.environ["GZPB_FULL_COPIES"])
    per_write = 40                                       # the caller's buffer: 40 copies of the corpus (218.6 MB) per write()
    corpus = synth.corpus()
    buf = C.create_string_buffer(corpus * per_write, S * per_write)
    total_in = S * copies
    devs = (C.c_int * ndev)(*range(ndev))
    batch_bytes = inflight * bs
    samples = {}
    state = {"out": 0, "calls": 0}

    @_lib.SINK_FN
    def sink(_u, p, k):
        j = state["calls"]
        state["calls"] += 1
        if j % 24 == 5 and len(samples) < 12:            # keep a few whole batches for the decode check afterwards
            samples[j] = C.string_at(p, k)
        state["out"] += k
        return 0

    h = C.c_void_p()
    rc = L.gzpb_writer_create_multi(C.byref(h), devs, ndev, fmt, level, bs, inflight, C.cast(sink, C.c_void_p), None)
    assert rc == 0, L.gzpb_strerror(rc)
    threads = args.copy_threads or host_threads()
    assert L.gzpb_writer_set_copy_threads(h, threads) == 0
    log("writer over %d GPU(s) ready, %d copy threads" % (ndev, threads))
    cbase = C.addressof(buf)

    def produce(dst, stream_off, k):
        """k bytes of the repeated corpus from stream offset `stream_off`, written at dst (ctypes releases the GIL)."""
        pos, s0 = 0, stream_off % S
        while pos < k:
            m = min(S * per_write - s0, k - pos)
            C.memmove(dst + pos, cbase + s0, m)
            pos += m
            s0 = 0

    t0 = time.perf_counter()
    if args.feed == "write":
        left = copies
        while left:
            k = min(per_write, left)
            rc = L.gzpb_writer_write(h, buf, S * k)
            assert rc == 0, L.gzpb_strerror(rc)
            left -= k
    else:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=threads)
        done = 0
        p, room = C.c_void_p(0), C.c_size_t(0)
        while done < total_in:
            assert L.gzpb_writer_reserve(h, C.byref(p), C.byref(room)) == 0
            n = min(room.value, total_in - done)
            piece = max(1 << 20, (n + threads - 1) // threads)
            futs = [pool.submit(produce, p.value + o, done + o, min(piece, n - o)) for o in range(0, n, piece)]
            for f in futs:
                f.result()
            rc = L.gzpb_writer_commit(h, n)
            assert rc == 0, L.gzpb_strerror(rc)
            done += n
        pool.shutdown()
    rc = L.gzpb_writer_finish(h)
    dt = time.perf_counter() - t0
    assert rc == 0, L.gzpb_strerror(rc)
    bi, bo, nb, sc = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    L.gzpb_writer_stats(h, C.byref(bi), C.byref(bo), C.byref(nb), C.byref(sc))
    L.gzpb_writer_destroy(h)
    assert bi.value == total_in and bo.value == state["out"]
    # every kept batch j holds blocks [j * inflight, (j + 1) * inflight) of the stream: stock zlib must give the input back
    checked = 0
    for j, data in samples.items():
        pos, got = 0, []
        while pos + 18 < len(data):
            size = int.from_bytes(data[pos + 16:pos + 18], "little") + 1
            got.append(zlib.decompress(data[pos + 18: pos + size - 8], -15))
            pos += size
        joined = b"".join(got)
        assert joined == synth.corpus_stream(len(joined), j * batch_bytes), "batch %d does not decode to its input" % j
        checked += len(joined)
    print(json.dumps({"metric": "bgzf_l6_full_stream_throughput", "value": total_in / dt / GIB, "unit": UNIT, "n_gpus": ndev, "seconds": dt,
                      "bytes_in": total_in, "bytes_out": bo.value, "ratio": bo.value / total_in, "device_batches": nb.value, "sink_calls": sc.value,
                      "copy_threads": threads, "host_threads": host_threads(), "feed": args.feed,
                      "api": ("gzpb_writer_create_multi + gzpb_writer_write(218.6 MB per call, pageable caller buffer, one host copy by %d threads) + gzpb_writer_finish; counting sink" % threads) if args.feed == "write" else
                             ("gzpb_writer_create_multi + gzpb_writer_reserve / _commit: %d producer threads generate the stream in place in the pinned slabs (no host copy) + gzpb_writer_finish; counting sink" % threads),
                      "decoded_input_bytes_checked": checked, "batches_checked": len(samples),
                      "workload": "shakespeare.txt x %d = %d B (BASELINE configs[1]), ONE ordered BGZF level-6 stream, 65280-B blocks" % (copies, total_in)}), flush=True)


if __name__ == "__main__":
    main()
