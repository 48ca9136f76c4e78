// gzpb_common.cuh — shared device helpers and layout constants for the sm_100a
// per-block encode kernels (the GPU replacement of gzp's worker loop,
// /root/reference/src/par/compress.rs:279-294).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gzpb {

// ---- DEFLATE constants (RFC 1951) and the libdeflate-style encoder limits ----
constexpr int kMinMatch = 3;
constexpr int kMaxMatch = 258;
constexpr int kNumLitlen = 288;
constexpr int kNumOffset = 32;
constexpr int kNumPrecode = 19;
constexpr int kEndOfBlock = 256;
constexpr int kFirstLenSym = 257;
constexpr int kMaxLitlenCw = 14;
constexpr int kMaxOffsetCw = 15;
constexpr int kMaxPreCw = 7;
constexpr int kWindow = 32768;
constexpr int kMinBlockLength = 5000;
constexpr int kSoftMaxBlockLength = 300000;
constexpr int kSeqStoreLength = 50000;
constexpr int kFastSoftMaxBlockLength = 65535;   // level 1 (deflate_compress_fastest)
constexpr int kFastSeqStoreLength = 8192;
constexpr int kObsPerCheck = 512;

// ---- HBM layout of one encode unit (one gzp block) ---------------------------
// A unit owns fixed-stride slots in every per-batch device array.
constexpr int kMaxUnitBytes = 65536;      // positions per unit the single-window kernels handle
constexpr int kInStride = 65536 + 64;     // input slot (padded so 16-byte bulk copies never run off the end)
constexpr int kOutPayloadOff = 32;        // payload starts 4-byte aligned; the container header sits right before it
constexpr int kOutStride = 73728;         // 32 + 65280 + 6528 + 8 + 28 rounded up
constexpr int kTokStride = 65536 + 64;    // tokens (u32) per unit

// match-table entry (u64) produced by the match kernel, consumed by the parser
//   [ 0, 8)  lenA-3 (0 = none)   depth D result
//   [ 8,23)  offA
//   [23,31)  lenB-3 (0 = none)   depth D>>1 result
//   [31,46)  offB
//   [46]     hash3 predecessor inside the window
//   [47,61)  off3 (0 = no usable 3-byte match; only offsets <= 8192 are kept)
__host__ __device__ inline uint64_t pack_entry(uint32_t lenA, uint32_t offA, uint32_t lenB, uint32_t offB,
                                               uint32_t n3ok, uint32_t off3)
{
    return (uint64_t)(lenA ? lenA - 3 : 0) | ((uint64_t)offA << 8) | ((uint64_t)(lenB ? lenB - 3 : 0) << 23) |
           ((uint64_t)offB << 31) | ((uint64_t)n3ok << 46) | ((uint64_t)off3 << 47);
}

struct LevelParams {
    int mode;   // 0 greedy, 1 lazy, 2 lazy2, -1 stored only
    int depth;  // max_search_depth
    int nice;   // nice_match_length
    int level;
    int ht;     // 1 = level 1: ht_matchfinder (15-bit hash4, 2-entry buckets = a depth-2 chain), fastest parser
};

__host__ __device__ inline bool level_params(int level, LevelParams *lp)
{
    lp->level = level;
    lp->ht = 0;
    switch (level) {
    case 0: lp->mode = -1; lp->depth = 0; lp->nice = 0; return true;
    case 1: lp->mode = 0; lp->depth = 2; lp->nice = 32; lp->ht = 1; return true;
    case 2: lp->mode = 0; lp->depth = 6; lp->nice = 10; return true;
    case 3: lp->mode = 0; lp->depth = 12; lp->nice = 14; return true;
    case 4: lp->mode = 0; lp->depth = 16; lp->nice = 30; return true;
    case 5: lp->mode = 1; lp->depth = 16; lp->nice = 30; return true;
    case 6: lp->mode = 1; lp->depth = 35; lp->nice = 65; return true;
    case 7: lp->mode = 1; lp->depth = 100; lp->nice = 130; return true;
    case 8: lp->mode = 2; lp->depth = 300; lp->nice = 258; return true;
    case 9: lp->mode = 2; lp->depth = 600; lp->nice = 258; return true;
    default: return false;
    }
}

__device__ __forceinline__ uint32_t lz_hash(uint32_t seq, int bits) { return (seq * 0x1E35A7BDu) >> (32 - bits); }

// unaligned little-endian 32-bit load from a word-aligned byte array
__device__ __forceinline__ uint32_t ld32u(const uint32_t *words, uint32_t byte_pos)
{
    uint32_t w = byte_pos >> 2, s = (byte_pos & 3) * 8;
    uint32_t lo = words[w], hi = words[w + 1];
    return __funnelshift_r(lo, hi, s);
}

// Kernel launch and dynamic shared memory go through two macros so that the same
// sources also build against tests/emu (a CPU SIMT emulator used only by the tests).
#ifndef GZPB_EMU
#define GZPB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define GZPB_DYN_SMEM(name) extern __shared__ __align__(128) uint8_t name[]
#else
#define GZPB_LAUNCH(kernel, grid, block, smem, stream, ...) gzpb_emu::launch_async((void *)(stream), dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
#define GZPB_DYN_SMEM(name) uint8_t *name = gzpb_emu::dyn_smem()
#endif

__device__ __forceinline__ uint32_t lanemask_lt()
{
#ifndef GZPB_EMU
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
#else
    return (1u << (threadIdx.x & 31)) - 1u;
#endif
}

// ---- CRC-32 (reflected 0xEDB88320) GF(2) helpers ---------------------------
constexpr uint32_t kCrcPoly = 0xEDB88320u;
constexpr uint32_t kCrc32cPoly = 0x82F63B78u;

// a(x) * b(x) mod P, reflected representation (bit 31 = x^0)
__host__ __device__ inline uint32_t gf2_mulmod(uint32_t a, uint32_t b, uint32_t poly)
{
    uint32_t p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b & 1) ? (b >> 1) ^ poly : (b >> 1);
    }
    return p;
}

// x^(8*nbytes) mod P
__host__ __device__ inline uint32_t gf2_xpow8(uint64_t nbytes, uint32_t poly)
{
    uint32_t r = 0x80000000u;          // x^0
    uint32_t sq = 0x00800000u;         // x^8
    while (nbytes) {
        if (nbytes & 1) r = gf2_mulmod(r, sq, poly);
        sq = gf2_mulmod(sq, sq, poly);
        nbytes >>= 1;
    }
    return r;
}

#ifndef GZPB_EMU
// ---- TMA 1-D bulk copy (cp.async.bulk) + mbarrier ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
// Bounded wait: a TMA transfer that never lands becomes a trap (a launch error the
// host reports as GZPB_ECUDA) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, phase)) {
        if (++spins > (1u << 22)) __trap();
    }
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
#else
// ---- tests/emu: the mbarrier is a small state machine, the bulk copy a memcpy ----------
struct EmuMbar { int32_t tx; uint16_t pending; uint16_t count_phase; };   // overlays the 64-bit barrier word
static_assert(sizeof(EmuMbar) == 8, "EmuMbar must overlay a uint64_t");
inline void emu_mbar_check(EmuMbar *b)
{
    if (b->pending == 0 && b->tx == 0) { b->count_phase ^= 0x8000u; b->pending = b->count_phase & 0x7FFFu; gzpb_emu::note_progress(); }
}
inline void mbar_init(uint64_t *bar, uint32_t count) { EmuMbar *b = (EmuMbar *)bar; b->tx = 0; b->pending = (uint16_t)count; b->count_phase = (uint16_t)count; }
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { EmuMbar *b = (EmuMbar *)bar; b->tx += (int32_t)bytes; b->pending--; emu_mbar_check(b); }
inline bool mbar_try_wait(uint64_t *bar, uint32_t phase)
{
    EmuMbar *b = (EmuMbar *)bar;
    if (gzpb_emu::tma_late()) { const uint32_t moved = gzpb_emu::tma_deliver(bar); if (moved) { b->tx -= (int32_t)moved; emu_mbar_check(b); } }
    return (uint32_t)(b->count_phase >> 15) != (phase & 1u);
}
inline void mbar_wait(uint64_t *bar, uint32_t phase) { while (!mbar_try_wait(bar, phase)) gzpb_emu::wait_yield(); }
inline void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    if (((uintptr_t)smem_dst & 15) || ((uintptr_t)gsrc & 15) || (bytes & 15)) gzpb_emu::trap("cp.async.bulk: dst/src/size must be 16-byte aligned");
    if (gzpb_emu::tma_late()) { gzpb_emu::tma_defer(bar, smem_dst, gsrc, bytes); return; }     // lands when somebody waits for it
    memcpy(smem_dst, gsrc, bytes);
    EmuMbar *b = (EmuMbar *)bar; b->tx -= (int32_t)bytes; emu_mbar_check(b);
}
inline void fence_proxy_async() {}
inline void fence_mbar_init() {}
inline void prefetch_l2(const void *) {}
#endif

}  // namespace gzpb
