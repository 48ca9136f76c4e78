"""tests/emu — CPU SIMT emulator build of the kernels.  TEST INFRASTRUCTURE ONLY.

The .cu sources of gzp_b200/csrc are compiled with g++ against tests/emu/shim
(every CUDA thread a fiber, see shim/cuda_runtime.h) into tests/emu/_build/
libgzpb_emu.so, which exports the same C ABI as the product library.  It lets
the CPU test suite run the *kernel logic* bit-for-bit against the oracle in a
container without a GPU.  Nothing under gzp_b200/ can load this library: the
product path fails loudly when libgzpb.so / an sm_100 device is missing.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "gzp_b200", "csrc")
# GZPB_EMU_ASAN=1: AddressSanitizer build in its own directory (run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
# and ASAN_OPTIONS=detect_leaks=0): "device" memory is host heap here, so an out-of-bounds access of a kernel or of the
# host runtime is reported with the kernel's source line — compute-sanitizer's memcheck without a GPU.
_ASAN = os.environ.get("GZPB_EMU_ASAN") == "1"
# GZPB_EMU_DEFS="-DGZPB_X=1 ...": compile-time kernel variants (the A/B candidates of a GPU session) in their own directory
_DEFS = os.environ.get("GZPB_EMU_DEFS", "").split()
_BUILD = os.path.join(_HERE, ("_build_asan" if _ASAN else "_build") +
                      ("_" + "".join(c if c.isalnum() else "_" for c in "".join(_DEFS)) if _DEFS else ""))
SO = os.path.join(_BUILD, "libgzpb_emu.so")
CXX = os.environ.get("CXX", "g++")
if os.environ.get("GZPB_EMU_ASAN") == "1" and os.path.exists("/usr/bin/g++"):
    CXX = "/usr/bin/g++"           # the distribution compiler knows where its libasan lives
FLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-fno-strict-aliasing", "-Wno-unused", "-I", os.path.join(_HERE, "shim"),
         "-I", _CSRC, "-DGZPB_EMU=1"] + _DEFS + (["-fsanitize=address", "-fno-omit-frame-pointer"] if _ASAN else [])


def build(force=False):
    os.makedirs(_BUILD, exist_ok=True)
    srcs = sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    deps += [os.path.join(_HERE, "shim", "cuda_runtime.h"), os.path.join(_HERE, "emu_runtime.cpp"),
             os.path.join(_ROOT, "include", "gzpb.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    objs = []
    procs = []
    for s in srcs + [os.path.join(_HERE, "emu_runtime.cpp")]:
        o = os.path.join(_BUILD, os.path.basename(s).rsplit(".", 1)[0] + ".o")
        procs.append(subprocess.Popen([CXX] + FLAGS + ["-x", "c++", "-c", s, "-o", o]))
        objs.append(o)
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emulator build failed")
    subprocess.check_call([CXX, "-shared", "-o", SO] + objs + ["-lpthread"] + (["-fsanitize=address"] if _ASAN else []))
    return SO


_lib = None


def lib():
    """ctypes handle of the emulated library with the product's prototypes."""
    global _lib
    if _lib is None:
        from gzp_b200 import _lib as product
        L = C.CDLL(build())
        for name, (res, args) in product._SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class EmuContext:
    """The slice of gzp_b200.Context the emulator tests need, bound to the emulated library."""

    def __init__(self, fmt, level, max_block_bytes=0, max_blocks_in_flight=8):
        self.L = lib()
        self.fmt, self.level = int(fmt), int(level)
        h = C.c_void_p()
        rc = self.L.gzpb_create(C.byref(h), 0, self.fmt, self.level, max_block_bytes, max_blocks_in_flight)
        if rc != 0:
            raise RuntimeError("gzpb_create (emu): %d %s" % (rc, self.L.gzpb_strerror(rc).decode()))
        self.h = h

    def close(self):
        if self.h:
            self.L.gzpb_destroy(self.h)
            self.h = None

    def encode_stream(self, data, buffer_size=0):
        data = bytes(data)
        n = len(data)
        bs = buffer_size or self.L.gzpb_default_bufsize(self.fmt)
        nblocks = max(1, (n + bs - 1) // bs)
        cap = 64 + nblocks * self.L.gzpb_encode_capacity(self.fmt, bs)
        out = C.create_string_buffer(cap)
        olen = C.c_size_t(0)
        src = C.create_string_buffer(data, n) if n else C.create_string_buffer(1)
        rc = self.L.gzpb_encode_stream(self.h, src, n, buffer_size, out, cap, C.byref(olen))
        if rc != 0:
            raise RuntimeError("gzpb_encode_stream (emu): %d %s" % (rc, self.L.gzpb_strerror(rc).decode()))
        return out.raw[:olen.value]


class EmuDecoder:
    """gzpb_decoder bound to the emulated library."""

    def __init__(self, fmt, max_blocks_in_flight=64):
        self.L = lib()
        self.fmt = int(fmt)
        h = C.c_void_p()
        rc = self.L.gzpb_decoder_create(C.byref(h), 0, self.fmt, max_blocks_in_flight)
        if rc != 0:
            raise RuntimeError("gzpb_decoder_create (emu): %d" % rc)
        self.h = h

    def close(self):
        if self.h:
            self.L.gzpb_decoder_destroy(self.h)
            self.h = None

    def decode_stream(self, data, out_cap=None):
        """Returns (status, bytes, found_crc, expected_crc)."""
        data = bytes(data)
        total = C.c_uint64(0)
        self.L.gzpb_scan_blocks(self.fmt, data, len(data), None, 0, None, None, C.byref(total))
        cap = out_cap if out_cap is not None else total.value + 64
        out = C.create_string_buffer(max(cap, 1))
        olen = C.c_size_t(0)
        rc = self.L.gzpb_decode_stream(self.h, data, len(data), out, cap, C.byref(olen), None)
        f, e = C.c_uint32(0), C.c_uint32(0)
        self.L.gzpb_decoder_last_check(self.h, C.byref(f), C.byref(e), None)
        return rc, out.raw[:olen.value], f.value, e.value
