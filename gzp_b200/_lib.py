"""ctypes binding of libgzpb.so (include/gzpb.h).  Fails loudly when the CUDA
library is missing — there is no CPU fallback in the product path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GZPB_LIB selects another build of the same library for A/B measurements (tests/perf_*.py); the default is the in-tree one
SO_PATH = os.environ.get("GZPB_LIB") or os.path.join(_HERE, "libgzpb.so")

GZIP, ZLIB, RAWDEFLATE, MGZIP, BGZF, SNAP = 0, 1, 2, 3, 4, 5


class BlockIn(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_size_t), ("dict", C.c_void_p), ("dict_len", C.c_size_t),
                ("is_last", C.c_int)]


class BlockOut(C.Structure):
    _fields_ = [("dst", C.c_void_p), ("cap", C.c_size_t), ("out_len", C.c_size_t), ("check_sum", C.c_uint32),
                ("check_amount", C.c_uint32), ("status", C.c_int)]


class BlockDesc(C.Structure):
    _fields_ = [("in_off", C.c_uint64), ("out_off", C.c_uint64), ("in_len", C.c_uint32), ("out_len", C.c_uint32),
                ("crc", C.c_uint32), ("pad", C.c_uint32)]


_SIGS = {
    "gzpb_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t]),
    "gzpb_destroy": (None, [C.c_void_p]),
    "gzpb_encode_batch": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(BlockIn), C.POINTER(BlockOut)]),
    "gzpb_encode_stream": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t)]),
    "gzpb_encode_stream_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t)]),
    "gzpb_encode_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "gzpb_encode_device_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "gzpb_unit_stride": (C.c_size_t, [C.c_void_p]),
    "gzpb_encode_capacity": (C.c_size_t, [C.c_int, C.c_size_t]),
    "gzpb_header": (C.c_size_t, [C.c_int, C.c_int, C.c_void_p]),
    "gzpb_footer": (C.c_size_t, [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]),
    "gzpb_crc32_combine": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint64]),
    "gzpb_adler32_combine": (C.c_uint32, [C.c_uint32, C.c_uint32, C.c_uint64]),
    "gzpb_default_bufsize": (C.c_size_t, [C.c_int]),
    "gzpb_needs_dict": (C.c_int, [C.c_int]),
    "gzpb_level_supported": (C.c_int, [C.c_int, C.c_int]),
    "gzpb_host_alloc": (C.c_void_p, [C.c_size_t]),
    "gzpb_host_free": (None, [C.c_void_p]),
    "gzpb_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "gzpb_kernel_ms": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "gzpb_launch_count": (C.c_uint64, [C.c_void_p]),
    "gzpb_ctx_variant": (C.c_char_p, [C.c_void_p]),
    "gzpb_writer_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "gzpb_writer_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "gzpb_writer_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gzpb_submit": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(BlockIn), C.POINTER(BlockOut), C.POINTER(C.c_uint64)]),
    "gzpb_poll": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int]),
    "gzpb_writer_reserve": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "gzpb_writer_commit": (C.c_int, [C.c_void_p, C.c_size_t]),
    "gzpb_compress_file": (C.c_int, [C.POINTER(C.c_int), C.c_size_t, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_char_p, C.c_char_p,
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gzpb_writer_set_copy_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "gzpb_writer_write": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "gzpb_writer_flush": (C.c_int, [C.c_void_p]),
    "gzpb_writer_finish": (C.c_int, [C.c_void_p]),
    "gzpb_writer_destroy": (None, [C.c_void_p]),
    "gzpb_block_header_size": (C.c_size_t, [C.c_int]),
    "gzpb_block_size": (C.c_long, [C.c_int, C.c_void_p, C.c_size_t]),
    "gzpb_scan_blocks": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.POINTER(BlockDesc), C.c_size_t, C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]),
    "gzpb_decoder_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_size_t]),
    "gzpb_decoder_destroy": (None, [C.c_void_p]),
    "gzpb_decode_stream": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t)]),
    "gzpb_decode_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gzpb_reader_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "gzpb_reader_read": (C.c_long, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "gzpb_reader_last_check": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "gzpb_reader_finish": (C.c_int, [C.c_void_p]),
    "gzpb_reader_destroy": (None, [C.c_void_p]),
    "gzpb_decoder_last_check": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "gzpb_decoder_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "gzpb_decoder_kernel_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "gzpb_decoder_launch_count": (C.c_uint64, [C.c_void_p]),
    "gzpb_bgzf_index": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gzpb_writer_bgzf_index": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gzpb_bgzf_virtual_offset": (C.c_uint64, [C.c_uint64, C.c_uint32]),
    "gzpb_strerror": (C.c_char_p, [C.c_int]),
    "gzpb_version": (C.c_char_p, []),
}

SINK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t)
SOURCE_FN = C.CFUNCTYPE(C.c_long, C.c_void_p, C.c_void_p, C.c_size_t)
EXPORTS = tuple(_SIGS)
_lib = None


def load():
    """Load libgzpb.so; raises ImportError (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -m gzp_b200.build` (nvcc, sm_100a). "
                "gzp_b200 has no CPU fallback.")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            if os.environ.get("GZPB_LIB") and not hasattr(lib, name):
                continue                                        # an older A/B build may lack the newest entry points
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
