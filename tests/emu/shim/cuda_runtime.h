// tests/emu/shim/cuda_runtime.h — TEST INFRASTRUCTURE ONLY.
//
// A minimal CPU SIMT emulator ("compute-sanitizer without a GPU") used by
// tests/test_emu_kernels.py: the .cu files under gzp_b200/csrc are compiled with
// g++ against this header instead of the CUDA toolkit, every CUDA thread becomes a
// fiber, one CTA runs at a time, and the warp/CTA collectives (__shfl_sync,
// __ballot_sync, __match_any_sync, __syncwarp, __syncthreads, mbarrier + bulk copy)
// are rendezvous points between fibers.  It checks the LOGIC of the kernels
// bit-for-bit against the oracle in a container that has no GPU; it says nothing
// about memory-model races or speed.  The product (gzp_b200/libgzpb.so) is never
// built from, linked against, or able to load anything in tests/emu.
#pragma once
#define GZPB_EMU 1
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

// ---- qualifiers -------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// ---- vector / dim types -----------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// ---- runtime API subset -----------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };
enum { cudaHostAllocPortable = 1, cudaHostAllocMapped = 2, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p);
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned flags);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *src, size_t n) { memcpy((void *)&sym, src, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMemcpyFromSymbol(void *dst, const T &sym, size_t n) { memcpy(dst, (const void *)&sym, n); return cudaSuccess; }

// ---- the SIMT engine (tests/emu/emu_runtime.cpp) ----------------------------
namespace gzpb_emu {
struct Self { uint3 tid, bid, bdim, gdim; unsigned lane, warp, linear; };
extern Self *g_self;
enum Op { OP_SYNCWARP, OP_BALLOT, OP_SHFL, OP_MATCH_ANY, OP_ANY, OP_ALL, OP_RED_ADD, OP_RED_OR, OP_RED_AND, OP_RED_XOR, OP_RED_MIN, OP_RED_MAX };
uint64_t collective(int op, unsigned mask, uint64_t val, uint64_t *all32);   // all32 (optional): every lane's value
void syncthreads();
void wait_yield();          // a fiber waiting on something another fiber must do
void note_progress();
uint8_t *dyn_smem();
long long clock();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void launch_async(void *stream, dim3 grid, dim3 block, size_t smem, std::function<void()> body);   // queued on the (lazy) stream
[[noreturn]] void trap(const char *why);
// Bulk copies (cp.async.bulk) land LATE by default: the bytes are moved only when a thread waits on the copy's
// mbarrier, i.e. at the last moment the hardware could deliver them.  A kernel that reads a tile before waiting for
// it, reuses the destination as scratch while the copy is in flight, or exits with a copy pending is wrong here every
// time, as it is sometimes on a GPU.  GZPB_EMU_TMA=eager copies at issue time.
bool tma_late();
void tma_defer(void *bar, void *dst, const void *src, uint32_t bytes);
uint32_t tma_deliver(void *bar);        // performs every copy pending on `bar`, returns the bytes moved
}  // namespace gzpb_emu

#define threadIdx (gzpb_emu::g_self->tid)
#define blockIdx (gzpb_emu::g_self->bid)
#define blockDim (gzpb_emu::g_self->bdim)
#define gridDim (gzpb_emu::g_self->gdim)

static inline void __syncthreads() { gzpb_emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { gzpb_emu::collective(gzpb_emu::OP_SYNCWARP, mask, 0, nullptr); }
static inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_BALLOT, mask, pred != 0, nullptr); }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
template <class T> static inline uint64_t emu_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    if (width != 32) gzpb_emu::trap("__shfl_sync: width != 32 not emulated");
    uint64_t all[32];
    gzpb_emu::collective(gzpb_emu::OP_SHFL, mask, emu_bits(v), all);
    return emu_unbits<T>(all[src & 31]);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
    if (width != 32) gzpb_emu::trap("__shfl_xor_sync: width != 32 not emulated");
    uint64_t all[32];
    gzpb_emu::collective(gzpb_emu::OP_SHFL, mask, emu_bits(v), all);
    return emu_unbits<T>(all[(gzpb_emu::g_self->lane ^ (unsigned)lanemask) & 31]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    if (width != 32) gzpb_emu::trap("__shfl_up_sync: width != 32 not emulated");
    uint64_t all[32];
    gzpb_emu::collective(gzpb_emu::OP_SHFL, mask, emu_bits(v), all);
    const unsigned l = gzpb_emu::g_self->lane;
    return l >= delta ? emu_unbits<T>(all[l - delta]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    if (width != 32) gzpb_emu::trap("__shfl_down_sync: width != 32 not emulated");
    uint64_t all[32];
    gzpb_emu::collective(gzpb_emu::OP_SHFL, mask, emu_bits(v), all);
    const unsigned l = gzpb_emu::g_self->lane;
    return l + delta < 32 ? emu_unbits<T>(all[l + delta]) : v;
}
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_MATCH_ANY, mask, emu_bits(v), nullptr); }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_RED_ADD, mask, v, nullptr); }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_RED_OR, mask, v, nullptr); }
static inline unsigned __reduce_and_sync(unsigned mask, unsigned v) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_RED_AND, mask, v, nullptr); }
static inline unsigned __reduce_xor_sync(unsigned mask, unsigned v) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_RED_XOR, mask, v, nullptr); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_RED_MIN, mask, v, nullptr); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return (unsigned)gzpb_emu::collective(gzpb_emu::OP_RED_MAX, mask, v, nullptr); }

// ---- bit intrinsics -----------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __ffsll(unsigned long long v) { return __builtin_ffsll((long long)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __clzll(unsigned long long v) { return v ? __builtin_clzll(v) : 64; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s)
{
    uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
    for (int i = 0; i < 4; i++) { unsigned sel = (s >> (4 * i)) & 0xF; unsigned byte = (unsigned)(v >> (8 * (sel & 7))) & 0xFF; if (sel & 8) byte = (byte & 0x80) ? 0xFF : 0; r |= byte << (8 * i); }
    return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline long long clock64() { return gzpb_emu::clock(); }
static inline void __trap() { gzpb_emu::trap("__trap()"); }
static inline void __nanosleep(unsigned) { gzpb_emu::wait_yield(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

// ---- atomics (one OS thread: plain read-modify-write) ---------------------------
#define EMU_ATOMIC(name, T, expr) static inline T name(T *p, T v) { T old = *p; *p = (expr); return old; }
EMU_ATOMIC(atomicAdd, unsigned, old + v)
EMU_ATOMIC(atomicAdd, int, old + v)
EMU_ATOMIC(atomicAdd, unsigned long long, old + v)
EMU_ATOMIC(atomicSub, unsigned, old - v)
EMU_ATOMIC(atomicOr, unsigned, old | v)
EMU_ATOMIC(atomicOr, int, old | v)
EMU_ATOMIC(atomicOr, unsigned long long, old | v)
EMU_ATOMIC(atomicAnd, unsigned, old & v)
EMU_ATOMIC(atomicXor, unsigned, old ^ v)
EMU_ATOMIC(atomicMax, unsigned, old > v ? old : v)
EMU_ATOMIC(atomicMax, int, old > v ? old : v)
EMU_ATOMIC(atomicMax, unsigned long long, old > v ? old : v)
EMU_ATOMIC(atomicMin, unsigned, old < v ? old : v)
EMU_ATOMIC(atomicMin, int, old < v ? old : v)
EMU_ATOMIC(atomicExch, unsigned, v)
EMU_ATOMIC(atomicExch, int, v)
#undef EMU_ATOMIC
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) { unsigned old = *p; if (old == cmp) *p = v; return old; }
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v) { unsigned long long old = *p; if (old == cmp) *p = v; return old; }

// ---- CUDA's integer min/max overload set ----------------------------------------
#define EMU_MINMAX(T) static inline T min(T a, T b) { return a < b ? a : b; } static inline T max(T a, T b) { return a > b ? a : b; }
EMU_MINMAX(int)
EMU_MINMAX(unsigned)
EMU_MINMAX(long)
EMU_MINMAX(unsigned long)
EMU_MINMAX(long long)
EMU_MINMAX(unsigned long long)
#undef EMU_MINMAX
static inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
static inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
static inline unsigned max(unsigned a, int b) { return a > (unsigned)b ? a : (unsigned)b; }
static inline unsigned max(int a, unsigned b) { return (unsigned)a > b ? (unsigned)a : b; }
