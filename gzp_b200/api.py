"""Host-side mirror of gzp's writer API on top of the C ABI (include/gzpb.h).

Names, argument meaning and error behaviour follow the reference:
  ParCompressBuilder / ParCompress  /root/reference/src/par/compress.rs:33-233, 248-468
  ZBuilder                          /root/reference/src/lib.rs:181-265
  format types                      /root/reference/src/deflate.rs:61,169,279,357,506; src/snap.rs:35
  GzpError                          /root/reference/src/lib.rs:114-163
  ParDecompressBuilder / ParDecompress  /root/reference/src/par/decompress.rs:16-111, 113-352
The chunker keeps the reference's exact semantics (strict '>' hold-back, flush()
emitting whatever is buffered — even an empty block —, finish() always sending a
final is_last block, the 32 KiB dictionary rule); the worker pool is replaced by
device batches handed to gzpb_encode_batch.
"""
import ctypes as C
import os

from . import _lib
from ._lib import BGZF, GZIP, MGZIP, RAWDEFLATE, SNAP, ZLIB, BlockIn, BlockOut

BUFSIZE = 131072      # lib.rs:105
DICT_SIZE = 32768     # lib.rs:108

_ERRNAMES = {-1: "BufferSize", -2: "NumThreads", -3: "BlockSizeExceeded", -4: "LibDeflaterCompress",
             -5: "LibDeflaterCompressionLvl", -6: "Io", -7: "ChannelSend", -8: "Cuda", -9: "Unknown", -10: "NoMem",
             -11: "InvalidHeader", -12: "InvalidCheck", -13: "LibDelfaterDecompress", -14: "InvalidBlockSize"}


class GzpError(Exception):
    """GzpError (lib.rs:114-163); `.variant` names the enum variant, `.code` the C status."""

    def __init__(self, code, detail=None):
        self.code = int(code)
        self.variant = _ERRNAMES.get(self.code, "Unknown")
        msg = detail or _lib.load().gzpb_strerror(self.code).decode()
        super().__init__(f"{self.variant}: {msg}")


class Compression:
    """flate2::Compression re-export (lib.rs:81)."""

    def __init__(self, level):
        self._level = int(level)

    def level(self):
        return self._level

    @classmethod
    def new(cls, level):
        return cls(level)

    @classmethod
    def none(cls):
        return cls(0)

    @classmethod
    def fast(cls):
        return cls(1)

    @classmethod
    def best(cls):
        return cls(9)

    @classmethod
    def default(cls):
        return cls(6)


def _lvl(x):
    return x.level() if isinstance(x, Compression) else int(x)


class _Format:
    ID = None
    NAME = ""

    @classmethod
    def new(cls):
        return cls()

    @property
    def DEFAULT_BUFSIZE(self):
        return _lib.load().gzpb_default_bufsize(self.ID)

    def needs_dict(self):
        return bool(_lib.load().gzpb_needs_dict(self.ID))

    def header(self, level):
        return header(self.ID, _lvl(level))

    def footer(self, check_sum, check_amount):
        return footer(self.ID, check_sum, check_amount)


class Gzip(_Format):
    ID, NAME = GZIP, "Gzip"


class Zlib(_Format):
    ID, NAME = ZLIB, "Zlib"


class RawDeflate(_Format):
    ID, NAME = RAWDEFLATE, "RawDeflate"


class Mgzip(_Format):
    ID, NAME = MGZIP, "Mgzip"


class Bgzf(_Format):
    ID, NAME = BGZF, "Bgzf"


class Snap(_Format):
    ID, NAME = SNAP, "Snap"


def header(fmt, level):
    b = C.create_string_buffer(16)
    n = _lib.load().gzpb_header(fmt, level, b)
    return b.raw[:n]


def footer(fmt, check_sum, check_amount):
    b = C.create_string_buffer(16)
    n = _lib.load().gzpb_footer(fmt, check_sum & 0xFFFFFFFF, check_amount & 0xFFFFFFFF, b)
    return b.raw[:n]


def crc32_combine(a, b, len_b):
    return _lib.load().gzpb_crc32_combine(a, b, len_b)


def adler32_combine(a, b, len_b):
    """Check::combine for Adler32 (check.rs:121-128)."""
    return _lib.load().gzpb_adler32_combine(a, b, len_b)


def encode_capacity(fmt, n):
    return _lib.load().gzpb_encode_capacity(fmt, n)


class Context:
    """One device context = the per-worker `Compressor` of the reference
    (FormatSpec::create_compressor, lib.rs:343-346), but for a whole GPU."""

    def __init__(self, fmt, level, device=0, max_block_bytes=0, max_blocks_in_flight=256):
        self._lib = _lib.load()
        self.fmt = fmt.ID if isinstance(fmt, _Format) or (isinstance(fmt, type) and issubclass(fmt, _Format)) else int(fmt)
        self.level = _lvl(level)
        h = C.c_void_p()
        rc = self._lib.gzpb_create(C.byref(h), device, self.fmt, self.level, max_block_bytes, max_blocks_in_flight)
        if rc != 0:
            raise GzpError(rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gzpb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode_blocks(self, blocks):
        """FormatSpec::encode + Check::update for a list of (bytes, dict|None, is_last).
        Returns a list of (encoded bytes, check_sum, check_amount); raises GzpError."""
        n = len(blocks)
        if n == 0:
            return []
        ins = (BlockIn * n)()
        outs = (BlockOut * n)()
        keep = []
        for i, (data, d, last) in enumerate(blocks):
            data = bytes(data)
            src = C.create_string_buffer(data, len(data)) if data else C.create_string_buffer(1)
            cap = self._lib.gzpb_encode_capacity(self.fmt, len(data)) + 64
            dst = C.create_string_buffer(cap)
            keep.append((src, dst))
            ins[i].ptr = C.cast(src, C.c_void_p)
            ins[i].len = len(data)
            if d:
                db = C.create_string_buffer(bytes(d), len(d))
                keep.append(db)
                ins[i].dict = C.cast(db, C.c_void_p)
                ins[i].dict_len = len(d)
            ins[i].is_last = int(bool(last))
            outs[i].dst = C.cast(dst, C.c_void_p)
            outs[i].cap = cap
        rc = self._lib.gzpb_encode_batch(self._h, n, ins, outs)
        if rc != 0:
            raise GzpError(rc)
        res = []
        k = 0
        for i in range(n):
            if outs[i].status != 0:
                raise GzpError(outs[i].status)
            dst = keep[k][1]
            k += 2 if blocks[i][1] else 1
            res.append((dst.raw[:outs[i].out_len], outs[i].check_sum, outs[i].check_amount))
        return res

    def encode_stream(self, data, buffer_size=0):
        """ParCompress over an in-memory input: header + write(data) + finish()."""
        data = bytes(data)
        n = len(data)
        bs = buffer_size or self._lib.gzpb_default_bufsize(self.fmt)
        nblocks = max(1, (n + bs - 1) // bs)
        cap = 64 + sum(self._lib.gzpb_encode_capacity(self.fmt, min(bs, n - i * bs)) for i in range(nblocks)) if nblocks < 4096 \
            else 64 + nblocks * self._lib.gzpb_encode_capacity(self.fmt, bs)
        out = C.create_string_buffer(cap)
        olen = C.c_size_t(0)
        src = C.create_string_buffer(data, n) if n else C.create_string_buffer(1)
        rc = self._lib.gzpb_encode_stream(self._h, src, n, buffer_size, out, cap, C.byref(olen))
        if rc != 0:
            raise GzpError(rc)
        return out.raw[:olen.value]

    def set_profiling(self, on):
        self._lib.gzpb_set_profiling(self._h, int(on))

    def kernel_ms(self, name):
        ms = C.c_double(0)
        cnt = C.c_uint64(0)
        self._lib.gzpb_kernel_ms(self._h, name.encode(), C.byref(ms), C.byref(cnt))
        return ms.value, cnt.value

    def launch_count(self):
        return self._lib.gzpb_launch_count(self._h)


def encode_stream_multi(contexts, data, buffer_size=0):
    """ParCompress over an in-memory input as ONE ordered stream dealt over several GPUs (gzpb_encode_stream_multi):
    `contexts` = one Context per GPU, created with the same format, level and sizes.  Byte-identical to
    `contexts[0].encode_stream(data)`."""
    lib = _lib.load()
    data = bytes(data)
    n, c0 = len(data), contexts[0]
    bs = buffer_size or lib.gzpb_default_bufsize(c0.fmt)
    nblocks = max(1, (n + bs - 1) // bs)
    cap = 64 + nblocks * lib.gzpb_encode_capacity(c0.fmt, bs)
    out = C.create_string_buffer(cap)
    olen = C.c_size_t(0)
    src = C.create_string_buffer(data, n) if n else C.create_string_buffer(1)
    hs = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    rc = lib.gzpb_encode_stream_multi(hs, len(contexts), src, n, buffer_size, out, cap, C.byref(olen))
    if rc != 0:
        raise GzpError(rc)
    return out.raw[:olen.value]


class ParCompressBuilder:
    """ParCompressBuilder<F> (par/compress.rs:33-138).  `num_threads` sizes nothing on the GPU; with the native
    writer (`devices(...)`) it is the number of threads that make the one host copy of `write`
    (gzpb_writer_set_copy_threads).  `blocks_in_flight` is the device analogue of the 2*num_threads channel bound."""

    def __init__(self, fmt=Gzip):
        self.format = fmt() if isinstance(fmt, type) else fmt
        self._buffer_size = self.format.DEFAULT_BUFSIZE       # par/compress.rs:56
        self._num_threads = 0
        self._level = Compression.new(3)                      # par/compress.rs:58
        self._pin = None
        self._device = 0
        self._devices = None
        self._blocks_in_flight = 256

    @classmethod
    def new(cls, fmt=Gzip):
        return cls(fmt)

    def buffer_size(self, n):
        if n < DICT_SIZE:                                      # par/compress.rs:68-74
            raise GzpError(-1, f"Invalid buffer size ({n}), must be >= {DICT_SIZE}")
        self._buffer_size = int(n)
        return self

    def compression_level(self, level):
        self._level = level if isinstance(level, Compression) else Compression.new(level)
        return self

    def num_threads(self, n):
        if n == 0:                                             # par/compress.rs:84-90
            raise GzpError(-2, "Invalid number of threads (0) selected.")
        self._num_threads = int(n)
        return self

    def pin_threads(self, first_core):
        self._pin = first_core
        return self

    def device(self, index):
        self._device = int(index)
        return self

    def blocks_in_flight(self, n):
        self._blocks_in_flight = int(n)
        return self

    def devices(self, indices):
        """Use the native pipelined writer (gzpb_writer_create_multi) over these GPUs: device batches of
        `blocks_in_flight` blocks are dealt round-robin, one ordered output."""
        self._devices = [int(i) for i in indices]
        if not self._devices:
            raise GzpError(-9, "devices() needs at least one GPU index")
        return self

    def from_writer(self, writer):
        if self._devices is not None:
            return NativeParCompress(self.format, writer, self._level, self._buffer_size, self._devices, self._blocks_in_flight,
                                     copy_threads=max(1, self._num_threads))
        return ParCompress(self.format, writer, self._level, self._buffer_size, self._device, self._blocks_in_flight)

    from_borrowed_writer = from_writer


class ParCompress:
    """ParCompress<F, W> as a Python `write`/`flush`/`finish` object
    (par/compress.rs:221-233, 332-362, 377-388, 413-468)."""

    def __init__(self, fmt, writer, level, buffer_size, device=0, blocks_in_flight=256):
        self.format = fmt
        self.writer = writer
        self.level = level
        self.buffer_size = buffer_size
        self._ctx = None
        self._ctx_args = (fmt.ID, level, device, buffer_size, blocks_in_flight)
        self._buf = bytearray()
        self._dict = None
        self._pending = []          # messages not yet handed to the device (FIFO = ticket order)
        self._max_pending = blocks_in_flight
        self._sum = 1 if fmt.ID == ZLIB else 0
        self._amount = 0
        self._finished = False
        self._error = None
        self._wrote_header = False

    @classmethod
    def builder(cls, fmt=Gzip):
        return ParCompressBuilder(fmt)

    # -- writer side (par/compress.rs:303-313) --
    def _drain(self):
        if not self._wrote_header:
            self.writer.write(self.format.header(self.level))
            self._wrote_header = True
        if not self._pending:
            return
        msgs, self._pending = self._pending, []
        if self._ctx is None:
            self._ctx = Context(*self._ctx_args)     # FormatSpec::create_compressor (par/compress.rs:278)
        try:
            res = self._ctx.encode_blocks(msgs)
        except GzpError as e:
            self._error = e
            raise
        for (data, _d, _l), (enc, s, a) in zip(msgs, res):
            # running_check.combine(&check) (par/compress.rs:308): CRC-32 for Gzip (check.rs:162), Adler-32 for Zlib (:121-128)
            if self.format.ID == GZIP:
                self._sum = crc32_combine(self._sum, s, len(data))
                self._amount = (self._amount + len(data)) & 0xFFFFFFFF
            elif self.format.ID == ZLIB:
                if len(data):
                    self._sum = adler32_combine(self._sum, s, len(data))
                self._amount = (self._amount + len(data)) & 0xFFFFFFFF
            self.writer.write(enc)

    def _send(self, block, dictionary, is_last):
        self._pending.append((block, dictionary, is_last))
        if len(self._pending) >= self._max_pending:
            self._drain()

    def write(self, data):
        if self._finished:
            raise GzpError(-7)
        if self._error:
            raise self._error
        self._buf.extend(data)
        while len(self._buf) > self.buffer_size:               # strict '>' (par/compress.rs:415)
            b = bytes(self._buf[:self.buffer_size])
            del self._buf[:self.buffer_size]
            d, self._dict = self._dict, (b[-DICT_SIZE:] if self.format.needs_dict() else None)
            self._send(b, d, False)
        return len(data)

    def _flush_last(self, is_last):
        while True:
            k = min(len(self._buf), self.buffer_size)
            b = bytes(self._buf[:k])
            del self._buf[:k]
            last = is_last and len(self._buf) == 0
            d, self._dict = self._dict, None
            if len(b) >= DICT_SIZE and not last and self.format.needs_dict():
                self._dict = b[-DICT_SIZE:]
            self._send(b, d, last)
            if len(self._buf) == 0:
                break

    def flush(self):
        if self._finished:
            raise GzpError(-7)
        self._flush_last(False)
        self._drain()

    def finish(self):
        if self._finished:
            raise GzpError(-7)
        self._flush_last(True)
        self._drain()
        self.writer.write(self.format.footer(self._sum, self._amount))
        if hasattr(self.writer, "flush"):
            self.writer.flush()
        self._finished = True
        if self._ctx is not None:
            self._ctx.close()
        return self.writer

    def __enter__(self):
        return self

    def __exit__(self, *exc):                                   # Drop -> finish (par/compress.rs:391-402)
        if not self._finished and exc[0] is None:
            self.finish()
        return False


class NativeParCompress:
    """ParCompress<F, W> on the C writer object (gzpb_writer_*, include/gzpb.h): the chunker, the pinned slabs,
    the batches in flight on one or several GPUs and the ordered hand-over to `writer.write` all live in
    libgzpb.so; Python only forwards `write` / `flush` / `finish` (par/compress.rs:377-388, 413-468)."""

    def __init__(self, fmt, writer, level, buffer_size, devices=(0,), blocks_in_flight=0, copy_threads=1):
        self._lib = _lib.load()
        self.format = fmt
        self.writer = writer
        self.level = level
        self.buffer_size = buffer_size
        self._sink_error = None
        self._finished = False
        self._index = None

        def _sink(_user, ptr, n):
            try:
                self.writer.write(C.string_at(ptr, n))
                return 0
            except Exception as e:                             # surfaces as GzpError::Io on the next call
                self._sink_error = e
                return 1

        self._cb = _lib.SINK_FN(_sink)
        h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        rc = self._lib.gzpb_writer_create_multi(C.byref(h), devs, len(devices), fmt.ID, _lvl(level), buffer_size,
                                                blocks_in_flight, C.cast(self._cb, C.c_void_p), None)
        if rc != 0:
            raise GzpError(rc)
        self._h = h
        if copy_threads > 1:
            self._lib.gzpb_writer_set_copy_threads(h, int(copy_threads))

    def _check(self, rc):
        if rc != 0:
            err = GzpError(rc)
            err.__cause__ = self._sink_error
            raise err

    def write(self, data):
        if self._h is None:
            raise GzpError(-7)
        data = bytes(data)
        self._check(self._lib.gzpb_writer_write(self._h, data, len(data)))
        return len(data)

    def flush(self):
        if self._h is None:
            raise GzpError(-7)
        self._check(self._lib.gzpb_writer_flush(self._h))

    def stats(self):
        v = [C.c_uint64(0) for _ in range(4)]
        self._lib.gzpb_writer_stats(self._h, *[C.byref(x) for x in v])
        return dict(zip(("bytes_in", "bytes_out", "batches", "sink_calls"), (x.value for x in v)))

    def bgzf_index(self):
        """.gzi index of the blocks written so far (Bgzf only); after finish(): of the whole stream."""
        if self._h is None:
            return self._index
        n = C.c_size_t(0)
        rc = self._lib.gzpb_writer_bgzf_index(self._h, None, 0, C.byref(n))
        if rc != 0:
            raise GzpError(rc)
        out = C.create_string_buffer(n.value)
        self._check(self._lib.gzpb_writer_bgzf_index(self._h, out, n.value, C.byref(n)))
        return out.raw[:n.value]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gzpb_writer_destroy(self._h)
            self._h = None
            self._cb = None

    def finish(self):
        if self._h is None:
            raise GzpError(-7)
        rc = self._lib.gzpb_writer_finish(self._h)
        self._finished = True
        self._index = self.bgzf_index() if self.format.ID == BGZF and rc == 0 else None
        self.close()
        self._check(rc)
        if hasattr(self.writer, "flush"):
            self.writer.flush()
        return self.writer

    def __enter__(self):
        return self

    def __exit__(self, *exc):                                   # Drop -> finish (par/compress.rs:391-402)
        if self._h is not None:
            if exc[0] is None:
                self.finish()
            else:
                self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ZBuilder:
    """ZBuilder<F, W> facade (lib.rs:181-265)."""

    def __init__(self, fmt=Gzip):
        self._b = ParCompressBuilder(fmt)
        self._threads = os.cpu_count() or 1     # num_cpus::get() (lib.rs:199)

    @classmethod
    def new(cls, fmt=Gzip):
        return cls(fmt)

    def buffer_size(self, n):
        self._b._buffer_size = int(n)       # ZBuilder defers validation to from_writer (lib.rs:210-213)
        return self

    def compression_level(self, level):
        self._b.compression_level(level)
        return self

    def num_threads(self, n):
        self._threads = int(n)
        self._b._num_threads = int(n)
        return self

    def pin_threads(self, first_core):
        self._b.pin_threads(first_core)
        return self

    def from_writer(self, writer):
        # lib.rs:242-264: more than one thread -> ParCompress, otherwise the synchronous writer
        # (quirk 8 of SURVEY App. D: num_threads == 1 goes to SyncZ, not to a 1-worker ParCompress)
        if self._threads > 1:
            if self._b._buffer_size < DICT_SIZE:
                raise GzpError(-1, f"Invalid buffer size ({self._b._buffer_size}), must be >= {DICT_SIZE}")
            return self._b.from_writer(writer)
        return SyncZBuilder(self._b.format).compression_level(self._b._level).device(self._b._device).from_writer(writer)


BGZF_EOF = bytes([0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0])  # bgzf.rs:24-38


class _BlockSyncWriter:
    """BgzfSyncWriter / MgzipSyncWriter (bgzf.rs:92-145, 315-355; mgzip.rs:74-125, 286-325): the single-threaded
    block writers behind SyncZ.  Every block goes through the same per-block encode path (`compress`,
    bgzf.rs:204-237 / mgzip.rs:187-218) — here one unit on the GPU.  The reference's quirks are kept:
    `write` emits at most ONE block per call, and only once `blocksize` bytes are buffered."""

    FORMAT = None

    def __init__(self, writer, level, blocksize, device=0):
        self.writer = writer
        self.level = level
        self.blocksize = int(blocksize)
        self._buf = bytearray()
        self._ctx = None
        self._device = device

    def _compress(self, block):
        if self._ctx is None:
            self._ctx = Context(self.FORMAT, self.level, self._device, max(self.blocksize, DICT_SIZE), 4)
        return self._ctx.encode_blocks([(bytes(block), None, False)])[0][0]

    def write(self, data):
        self._buf.extend(data)
        if len(self._buf) >= self.blocksize:                    # `if`, not `while` (bgzf.rs:322, mgzip.rs:293)
            b = bytes(self._buf[:self.blocksize])
            del self._buf[:self.blocksize]
            self.writer.write(self._compress(b))
        return len(data)

    def finish(self):
        self.flush()
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
        return self.writer


class BgzfSyncWriter(_BlockSyncWriter):
    FORMAT = BGZF

    def __init__(self, writer, level, blocksize=65280, device=0):
        assert blocksize <= 65280                               # bgzf.rs:124
        super().__init__(writer, level, blocksize, device)

    def flush(self):
        # BGZF_EOF after EVERY remaining block, none when nothing is buffered (bgzf.rs:334-343; App. D quirk 3)
        while self._buf:
            k = min(len(self._buf), 65280)
            b = bytes(self._buf[:k])
            del self._buf[:k]
            self.writer.write(self._compress(b))
            self.writer.write(BGZF_EOF)
        if hasattr(self.writer, "flush"):
            self.writer.flush()


class MgzipSyncWriter(_BlockSyncWriter):
    FORMAT = MGZIP

    def __init__(self, writer, level, blocksize=BUFSIZE, device=0):
        super().__init__(writer, level, blocksize, device)

    def _compress(self, block):
        if self._ctx is None or len(block) > self._ctx_cap:     # flush() sends everything buffered as one block (mgzip.rs:309-315)
            if self._ctx is not None:
                self._ctx.close()
            self._ctx_cap = max(len(block), self.blocksize, DICT_SIZE)
            self._ctx = Context(self.FORMAT, self.level, self._device, self._ctx_cap, 4)
        return self._ctx.encode_blocks([(bytes(block), None, False)])[0][0]

    def flush(self):
        if self._buf:
            b = bytes(self._buf)
            self._buf.clear()
            self.writer.write(self._compress(b))
        if hasattr(self.writer, "flush"):
            self.writer.flush()


class SyncZ:
    """SyncZ<W> (syncz.rs:59-87): `write` / `flush` / `finish` over the format's synchronous writer."""

    def __init__(self, inner):
        self.inner = inner

    @staticmethod
    def builder(fmt=Gzip):
        return SyncZBuilder(fmt)

    def write(self, data):
        return self.inner.write(data)

    def flush(self):
        return self.inner.flush()

    def finish(self):
        inner, self.inner = self.inner, None
        return inner.finish()


class SyncZBuilder:
    """SyncZBuilder<F, W> (syncz.rs:12-57).  Bgzf / Mgzip get their block sync writers.  For Gzip / Zlib /
    RawDeflate the reference wraps flate2's streaming encoders (deflate.rs:145-154, 255-264, 334-343) — one
    continuous zlib-ng stream; this engine has no such single-stream encoder, so those formats are served by a
    ParCompress with the format's default block size: same container, decodable by the same readers, not the
    same bytes as flate2's stream."""

    def __init__(self, fmt=Gzip):
        self.format = fmt() if isinstance(fmt, type) else fmt
        self._level = Compression.new(3)                          # syncz.rs:32
        self._device = 0

    @classmethod
    def new(cls, fmt=Gzip):
        return cls(fmt)

    def compression_level(self, level):
        self._level = level if isinstance(level, Compression) else Compression.new(level)
        return self

    def device(self, index):
        self._device = int(index)
        return self

    def from_writer(self, writer):
        if self.format.ID == BGZF:
            return SyncZ(BgzfSyncWriter(writer, self._level, device=self._device))
        if self.format.ID == MGZIP:
            return SyncZ(MgzipSyncWriter(writer, self._level, device=self._device))
        return SyncZ(ParCompress(self.format, writer, self._level, self.format.DEFAULT_BUFSIZE, self._device, 16))


def compress_file(in_path, out_path, fmt=Gzip, level=3, buffer_size=0, devices=(0,), blocks_in_flight=0):
    """File -> compressed file through the native writer (gzpb_compress_file): read(2) lands in the pinned
    slabs, the ordered output is written from pinned memory.  Returns (bytes_in, bytes_out)."""
    fmt = fmt() if isinstance(fmt, type) else fmt
    devs = (C.c_int * len(devices))(*devices)
    bi, bo = C.c_uint64(0), C.c_uint64(0)
    rc = _lib.load().gzpb_compress_file(devs, len(devices), fmt.ID, _lvl(level), buffer_size, blocks_in_flight,
                                        os.fsencode(in_path), os.fsencode(out_path), C.byref(bi), C.byref(bo))
    if rc != 0:
        raise GzpError(rc)
    return bi.value, bo.value


def bgzf_index(stream):
    """.gzi index of a BGZF stream (htslib `bgzip -i` layout); host-only, needs no GPU."""
    stream = bytes(stream)
    L = _lib.load()
    n = C.c_size_t(0)
    rc = L.gzpb_bgzf_index(stream, len(stream), None, 0, C.byref(n))
    if rc != 0:
        raise GzpError(rc)
    out = C.create_string_buffer(n.value)
    rc = L.gzpb_bgzf_index(stream, len(stream), out, n.value, C.byref(n))
    if rc != 0:
        raise GzpError(rc)
    return out.raw[:n.value]


def bgzf_virtual_offset(block_offset, within_block):
    return _lib.load().gzpb_bgzf_virtual_offset(block_offset, within_block)


class Decoder:
    """One device decoder = the per-worker `Decompressor` of the reference
    (BlockFormatSpec::create_decompressor, deflate.rs:372-381, 521-530), for a whole GPU."""

    def __init__(self, fmt, device=0, max_blocks_in_flight=5328):
        self._lib = _lib.load()
        self.fmt = fmt.ID if isinstance(fmt, _Format) or (isinstance(fmt, type) and issubclass(fmt, _Format)) else int(fmt)
        h = C.c_void_p()
        rc = self._lib.gzpb_decoder_create(C.byref(h), device, self.fmt, max_blocks_in_flight)
        if rc != 0:
            raise GzpError(rc)
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gzpb_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def header_size(self):
        return self._lib.gzpb_block_header_size(self.fmt)

    def decode(self, data, partial=False):
        """Decode the whole members in `data`.  Returns (decoded bytes, consumed input bytes).
        partial=True leaves an incomplete trailing member unconsumed; otherwise it is an Io error."""
        data = bytes(data)
        total = C.c_uint64(0)
        self._lib.gzpb_scan_blocks(self.fmt, data, len(data), None, 0, None, None, C.byref(total))
        out = C.create_string_buffer(max(1, total.value))
        olen = C.c_size_t(0)
        used = C.c_size_t(0)
        rc = self._lib.gzpb_decode_stream(self._h, data, len(data), out, total.value, C.byref(olen),
                                          C.byref(used) if partial else None)
        if rc != 0:
            raise self._error(rc)
        return out.raw[:olen.value], (used.value if partial else len(data))

    def _error(self, rc):
        if rc == -12:
            f, e = C.c_uint32(0), C.c_uint32(0)
            self._lib.gzpb_decoder_last_check(self._h, C.byref(f), C.byref(e), None)
            err = GzpError(rc, f"Invalid checksum, found {f.value}, expected {e.value}")      # lib.rs:139-140
            err.found, err.expected = f.value, e.value
            return err
        return GzpError(rc)

    def launch_count(self):
        return self._lib.gzpb_decoder_launch_count(self._h)


class ParDecompressBuilder:
    """ParDecompressBuilder<F: BlockFormatSpec> (par/decompress.rs:16-111).  `num_threads` is kept for API
    parity; the device batch (`blocks_in_flight`) replaces the 2*num_threads channel bound (:73)."""

    def __init__(self, fmt=Bgzf):
        self.format = fmt() if isinstance(fmt, type) else fmt
        if self.format.ID not in (BGZF, MGZIP):
            raise GzpError(-9, "ParDecompress needs a BlockFormatSpec format (Mgzip or Bgzf)")
        self._buffer_size = BUFSIZE                              # par/decompress.rs:34
        self._num_threads = 1
        self._pin = None
        self._device = 0
        self._blocks_in_flight = 5328

    @classmethod
    def new(cls, fmt=Bgzf):
        return cls(fmt)

    def buffer_size(self, n):
        if n < DICT_SIZE:                                         # par/decompress.rs:41-47
            raise GzpError(-1, f"Invalid buffer size ({n}), must be >= {DICT_SIZE}")
        self._buffer_size = int(n)
        return self

    def num_threads(self, n):
        if n == 0:                                                # par/decompress.rs:50-56
            raise GzpError(-2, "Invalid number of threads (0) selected.")
        self._num_threads = int(n)
        return self

    def maybe_num_threads(self, n):                               # par/decompress.rs:90-93
        self._num_threads = int(n)
        return self

    def pin_threads(self, first_core):
        self._pin = first_core
        return self

    def device(self, index):
        self._device = int(index)
        return self

    def blocks_in_flight(self, n):
        self._blocks_in_flight = int(n)
        return self

    def native(self, on=True):
        """Use the C reader object (gzpb_reader_*) instead of the Python reader loop."""
        self._native = bool(on)
        return self

    def from_reader(self, reader):
        if getattr(self, "_native", False):
            return NativeParDecompress(self.format, reader, self._device, self._blocks_in_flight)
        return ParDecompress(self.format, reader, self._buffer_size, self._device, self._blocks_in_flight)

    def maybe_par_from_reader(self, reader):
        # the reference falls back to a single-threaded MultiGzDecoder for 0 threads (:96-102);
        # this engine has no CPU decoder, so the GPU path serves both cases
        return self.from_reader(reader)


class NativeParDecompress:
    """ParDecompress<F> on the C reader object (gzpb_reader_*, include/gzpb.h): the reader loop, the pinned
    buffers and the member-parallel GPU decode live in libgzpb.so; Python forwards `read` (par/decompress.rs:238-287)."""

    def __init__(self, fmt, reader, device=0, blocks_in_flight=0, chunk_bytes=0):
        self._lib = _lib.load()
        self.format = fmt
        self.reader = reader
        self._src_error = None

        def _source(_user, buf, cap):
            try:
                b = self.reader.read(cap)
                if not b:
                    return 0
                C.memmove(buf, bytes(b), len(b))
                return len(b)
            except Exception as e:
                self._src_error = e
                return -1

        self._cb = _lib.SOURCE_FN(_source)
        h = C.c_void_p()
        rc = self._lib.gzpb_reader_create(C.byref(h), device, fmt.ID, blocks_in_flight, chunk_bytes, C.cast(self._cb, C.c_void_p), None)
        if rc != 0:
            raise GzpError(rc)
        self._h = h

    def _error(self, rc):
        if rc == -12:
            f, e = C.c_uint32(0), C.c_uint32(0)
            self._lib.gzpb_reader_last_check(self._h, C.byref(f), C.byref(e))
            err = GzpError(rc, f"Invalid checksum, found {f.value}, expected {e.value}")      # lib.rs:139-140
            err.found, err.expected = f.value, e.value
            return err
        err = GzpError(rc)
        err.__cause__ = self._src_error
        return err

    def read(self, n=-1):
        if self._h is None:
            raise GzpError(-7)
        out = bytearray()
        want = n if n is not None and n >= 0 else None
        while want is None or len(out) < want:
            k = (1 << 22) if want is None else min(want - len(out), 1 << 26)
            buf = C.create_string_buffer(k)
            got = self._lib.gzpb_reader_read(self._h, buf, k)
            if got < 0:
                raise self._error(got)
            if got == 0:
                break
            out += buf.raw[:got]
        return bytes(out)

    def finish(self):
        """Close things in such a way as to get errors (par/decompress.rs:222-236)."""
        if self._h is None:
            return
        rc = self._lib.gzpb_reader_finish(self._h)
        err = self._error(rc) if rc != 0 else None
        self.close()
        if err:
            raise err

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gzpb_reader_destroy(self._h)
            self._h = None
            self._cb = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ParDecompress:
    """ParDecompress<F> as a Python `read` object (par/decompress.rs:113-352): the reader loop
    (:190-207) pulls whole members from `reader`, the GPU decodes them in batches, `read`
    hands the decoded bytes out in stream order."""

    _CHUNK = 8 << 20

    def __init__(self, fmt, reader, buffer_size=BUFSIZE, device=0, blocks_in_flight=5328):
        self.format = fmt
        self.reader = reader
        self.buffer_size = buffer_size
        self._dec = Decoder(fmt.ID, device, blocks_in_flight)
        self._in = bytearray()
        self._out = bytearray()
        self._eof = False
        self._error = None

    @classmethod
    def builder(cls, fmt=Bgzf):
        return ParDecompressBuilder(fmt)

    def _fill(self):
        chunk = self.reader.read(self._CHUNK)
        if chunk:
            self._in.extend(chunk)
        else:
            self._eof = True
        try:
            if self._eof:
                # a short trailing header is EOF (:193, 205-206); a truncated member is an Io error (:197)
                hs = self._dec.header_size()
                if len(self._in) >= hs:
                    out, used = self._dec.decode(self._in, partial=False)
                    self._out.extend(out)
                self._in.clear()
            else:
                out, used = self._dec.decode(self._in, partial=True)
                self._out.extend(out)
                del self._in[:used]
        except GzpError as e:
            self._error = e
            raise

    def read(self, n=-1):
        if self._error:
            raise self._error
        while (n < 0 or len(self._out) < n) and not self._eof:
            self._fill()
        k = len(self._out) if n < 0 else min(n, len(self._out))
        b = bytes(self._out[:k])
        del self._out[:k]
        return b

    def finish(self):
        """Close things in such a way as to get errors (:222-236)."""
        self._dec.close()
        if self._error:
            raise self._error

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._dec.close()
        return False
