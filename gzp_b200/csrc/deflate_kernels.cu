// deflate_kernels.cu — sm_100a kernels for the DEFLATE-family per-block encode
// path (Bgzf / Mgzip / Gzip / Zlib / RawDeflate), i.e. the GPU replacement of
// `format.encode(chunk, ...)` + `check.update(chunk)` in gzp's worker loop
// (/root/reference/src/par/compress.rs:279-294 -> src/bgzf.rs:204-237,
// src/mgzip.rs:187-218).  One thread block per gzp block.
//
// Pipeline per batch of units (a unit = one gzp block, <= 65536 positions):
//   k_check  : CRC-32 / Adler-32 of the unit's data (Check::update)
//   k_split + k_link : hash-chain links of every position (hash4 -> next4[], hash3 -> prev3[])
//   k_match  : TMA-staged input + chains in shared memory; per-position longest
//              match for search depth D and D/2 (parse-independent)
//   k_emit   : sequential lazy/greedy parse over the match table, block
//              splitting, Huffman construction, parallel bit packing, container
//              header/footer
//   k_scan / k_gather : exclusive scan of block sizes + compaction into stream order
// Results are bit-identical to oracle/deflate_oracle.c (tests/test_gpu_parity.py).
#include <stdio.h>
#include <stdlib.h>

#include "gzpb_common.cuh"
#include "deflate_kernels.cuh"
#include <type_traits>

namespace gzpb {

__constant__ uint32_t c_crc_tab[4][256];      // slicing-by-4, reflected 0xEDB88320
__constant__ uint32_t c_xpow512[1024];        // x^(8*512*j) mod P
__constant__ uint32_t c_xpow512k[16];         // x^(8*512*1024*t) mod P: units longer than 512 KiB (up to 4 MiB + dictionary)
__device__ uint32_t g_xpow64[1024];            // x^(8*64*j) mod P: k_split's 64-byte CRC slices (indexed per thread: global, not __constant__)
__constant__ uint16_t c_static_litlen_cw[288];
__constant__ uint8_t c_static_litlen_len[288];
__constant__ uint8_t c_min_lens[80];
__device__ unsigned long long g_phase[32];   // SM-cycle accounting of kernel phases (debug/profiling aid)

static uint32_t h_bitrev(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i); return r; }

void upload_deflate_constants()
{
    static uint32_t tab[4][256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (kCrcPoly & (0u - (c & 1)));
        tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int s = 1; s < 4; s++) tab[s][i] = (tab[s - 1][i] >> 8) ^ tab[0][tab[s - 1][i] & 0xFF];
    cudaMemcpyToSymbol(c_crc_tab, tab, sizeof tab);
    static uint32_t xp[1024];
    for (int j = 0; j < 1024; j++) xp[j] = gf2_xpow8((uint64_t)512 * j, kCrcPoly);
    cudaMemcpyToSymbol(c_xpow512, xp, sizeof xp);
    static uint32_t xpk[16];
    for (int t = 0; t < 16; t++) xpk[t] = gf2_xpow8((uint64_t)512 * 1024 * t, kCrcPoly);
    cudaMemcpyToSymbol(c_xpow512k, xpk, sizeof xpk);
    static uint32_t xp64[1024];
    for (int j = 0; j < 1024; j++) xp64[j] = gf2_xpow8((uint64_t)64 * j, kCrcPoly);
    cudaMemcpyToSymbol(g_xpow64, xp64, sizeof xp64);
    uint16_t cw[288]; uint8_t ln[288];
    for (int s = 0; s < 288; s++) {
        uint32_t code; int len;
        if (s < 144) { len = 8; code = 0x30 + s; }
        else if (s < 256) { len = 9; code = 0x190 + (s - 144); }
        else if (s < 280) { len = 7; code = s - 256; }
        else { len = 8; code = 0xC0 + (s - 280); }
        cw[s] = (uint16_t)h_bitrev(code, len); ln[s] = (uint8_t)len;
    }
    cudaMemcpyToSymbol(c_static_litlen_cw, cw, sizeof cw);
    cudaMemcpyToSymbol(c_static_litlen_len, ln, sizeof ln);
    static const uint8_t min_lens[80] = {9,9,9,9,9,9,8,8,7,7,6,6,6,6,6,6,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
                                         5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
                                         4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4};
    cudaMemcpyToSymbol(c_min_lens, min_lens, sizeof min_lens);
}

// Sub-unit geometry.  A unit (one gzp block, with its optional 32 KiB dictionary in
// front) longer than 65 536 positions is searched in segments of `seg` new
// positions; each segment's chain/match kernels see a sub-unit = [32 KiB halo |
// seg new positions | 262 bytes of look-ahead] <= 65 536 bytes.  Links beyond the
// 32 KiB window are never followed, so the halo reproduces the full-history chains
// exactly (bit-identical results, tests/test_gpu_parity.py).
struct Geo {
    const uint8_t *in;         // unit slots
    const uint32_t *unit_len;  // dict + data bytes
    const uint32_t *unit_dict; // dictionary bytes in front of the data
    uint32_t in_stride, m_stride, tok_stride, out_stride;
    uint32_t spu, seg;         // sub-units per unit, new positions per sub-unit
};
struct Sub { uint32_t u, h, len, nb, ne, quirk; bool valid; };

__device__ __forceinline__ Sub sub_geometry(const Geo &g, uint32_t sub)
{
    Sub r;
    r.u = sub / g.spu;
    const uint32_t k = sub % g.spu;
    const uint32_t n = g.unit_len[r.u], dict = g.unit_dict[r.u];
    if (g.spu == 1) {
        r.h = 0; r.len = n; r.nb = dict; r.ne = n; r.quirk = (dict == 0); r.valid = n > dict;
        return r;
    }
    const uint32_t a = dict + k * g.seg;
    r.valid = a < n;
    const uint32_t b = min(a + g.seg, n);
    r.h = a > (uint32_t)kWindow ? ((a - kWindow) & ~15u) : 0u;
    const uint32_t t = min(n, b + 262u);
    r.len = r.valid ? t - r.h : 0; r.nb = a - r.h; r.ne = b - r.h; r.quirk = (r.h == 0 && dict == 0);
    return r;
}

constexpr uint32_t kNone16 = 0xFFFFu;

__device__ __forceinline__ uint32_t ldg32u(const uint32_t *__restrict__ words, uint32_t byte_pos)
{
    uint32_t w = byte_pos >> 2, s = (byte_pos & 3) * 8;
    return __funnelshift_r(__ldg(words + w), __ldg(words + w + 1), s);
}

// =============================================================================
// k_split + k_link: hash-chain links (hash4 -> next4[], hash3 -> prev3[]) of every position — the insertion side of
// libdeflate's hc_matchfinder restated: every position p <= n-5 is inserted, position 0 under hash 0 (next_hashes
// starts at {0,0}); next4[p] = distance to the previous position with the same 16-bit hash (0 = none / outside the
// 32 KiB window), prev3[p] likewise for the 15-bit hash of 3 bytes.
// A fully parallel pass (k_split) partitions the positions of a sub-unit stably by bucket range into position-ordered
// lists (16 ranges of the hash4 space, 4 of the hash3 space); each list is linked by its own single-warp job (k_link)
// that needs only a 12-16 KiB bucket table, so 13 (hash3) / 17 (hash4) jobs share an SM.  (The first design, one CTA per
// unit with the whole 192 KiB of bucket tables in shared memory, kept two latency-bound warps per SM busy: 6.7 ms per
// batch of 3256 units; lists + MATCH.ANY grouping 4.7 ms; grouping through the bucket table + ballots, below, 3.9 ms.)
// =============================================================================
constexpr int kSplitThreads = 1024;
constexpr int kBits4 = 12, kBits3 = 13;  // buckets per link job: hash4 8 KiB table + 4 KiB counts, hash3 16 KiB table
constexpr int kQ4Bits = 16 - kBits4, kQ3Bits = 15 - kBits3;
constexpr int kL4 = 1 << kQ4Bits, kL3 = 1 << kQ3Bits;   // lists: equal ranges of the hash4 / hash3 bucket space
constexpr int kSplitLists = kL4 + kL3;
constexpr int kLsStride = 64;            // list_start entries per sub-unit: [0..kL4] hash4 bounds, [kL4+1..kL4+1+kL3] hash3 bounds

// Lanes of the warp whose (active) list id equals this lane's: the id has only BITS bits, so BITS ballots do what a
// MATCH.ANY does at a fraction of its latency (k_split ranks every position twice per pass).
// Strong (relaxed, CTA scope) 16-bit shared-memory accesses: several lanes of a warp storing different values to one
// address is k_link's grouping mechanism — conflicting STRONG writes are ordered by coherence and are no data race in the
// PTX memory model (conflicting weak ones would be); same STS.U16 / LDS.U16 instructions either way.
__device__ __forceinline__ void st_relaxed_shared_u16(uint16_t *p, uint32_t v)
{
#ifndef GZPB_EMU
    asm volatile("st.relaxed.cta.shared.u16 [%0], %1;" ::"r"(smem_u32(p)), "h"((uint16_t)v) : "memory");
#else
    *p = (uint16_t)v;
#endif
}
__device__ __forceinline__ uint32_t ld_relaxed_shared_u16(const uint16_t *p)
{
#ifndef GZPB_EMU
    uint16_t v;
    asm volatile("ld.relaxed.cta.shared.u16 %0, [%1];" : "=h"(v) : "r"(smem_u32(p)) : "memory");
    return v;
#else
    return *p;
#endif
}

template <int BITS, bool kAllActive = false>
__device__ __forceinline__ uint32_t same_list_mask(uint32_t id, bool act)
{
    uint32_t m = kAllActive ? 0xFFFFFFFFu : __ballot_sync(0xFFFFFFFFu, act);
#ifndef GZPB_EMU
    // four instructions per bit (bit test with predicate, VOTE, two predicated LOP3); the C++ form below compiles to seven
#pragma unroll
    for (int b = 0; b < BITS; b++)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t, v;\n\t"
                     "and.b32 t, %1, %2;\n\t"
                     "setp.ne.u32 p, t, 0;\n\t"
                     "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"
                     "@p lop3.b32 %0, %0, v, 0, 0xC0;\n\t"      // m & v
                     "@!p lop3.b32 %0, %0, v, 0, 0x30;\n\t"     // m & ~v
                     "}" : "+r"(m) : "r"(id), "r"(1u << b));
#else
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const uint32_t v = __ballot_sync(0xFFFFFFFFu, (id >> b) & 1u);
        m &= ((id >> b) & 1u) ? v : ~v;
    }
#endif
    return m;
}

__global__ void __launch_bounds__(kSplitThreads, 2)
k_split(const __grid_constant__ Geo g, uint32_t *__restrict__ lists, uint32_t *__restrict__ list_start,
        uint16_t *__restrict__ next4, uint16_t *__restrict__ prev3, int ht, uint32_t *__restrict__ sum_part, int check_kind)
{
    __shared__ uint32_t s_w[32][kSplitLists];     // per-warp member counts, then running bases
    __shared__ uint32_t s_start[kSplitLists];
    __shared__ uint32_t s_tab[4][256];            // CRC-32 slicing tables (check_kind 0)
    __shared__ uint32_t s_crc;
    __shared__ unsigned long long s_a, s_b;
    const Sub sb = sub_geometry(g, blockIdx.x);
    if (!sb.valid) {                               // no data in this sub-unit: the identity of Check::combine
        if (sum_part && threadIdx.x == 0) sum_part[blockIdx.x] = (check_kind == 1) ? 1u : 0u;
        return;
    }
    // The positions this sub-unit LINKS, in unit coordinates: its new positions [a, b) — sub-unit 0 also takes the
    // dictionary in front of the data (set_dictionary inserts it).  The 32 KiB halo that k_match stages in front of a
    // later sub-unit is not linked again: links are distances, and k_link carries its bucket heads from one sub-unit
    // of a unit to the next.
    const uint32_t un = g.unit_len[sb.u], udict = g.unit_dict[sb.u];
    const uint32_t ksub = blockIdx.x % g.spu;
    const uint32_t ua = udict + ksub * g.seg;
    const uint32_t lo = ksub == 0 ? 0u : ua, hi_new = g.spu == 1 ? un : min(ua + g.seg, un);
    const uint32_t uinsert = un >= 5 ? un - 4 : 0;                 // positions p <= n-5 are inserted
    const uint32_t hi = min(hi_new, uinsert), ninsert = hi > lo ? hi - lo : 0;
    const uint32_t *uw = (const uint32_t *)(g.in + (size_t)sb.u * g.in_stride);
    const bool quirk0 = (ksub == 0 && udict == 0);                 // position 0 goes under hash 0 (next_hashes starts at {0,0})
    const uint32_t *inw = (const uint32_t *)(g.in + (size_t)sb.u * g.in_stride + sb.h);   // checksum pass below: sub-unit coordinates
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lt = lanemask_lt();
    const uint32_t ntiles = (ninsert + 31) / 32, tpw = (ntiles + 31) / 32;
    const uint32_t t0 = warp * tpw, t1 = min(ntiles, t0 + tpw);
    uint32_t *arr4 = lists + (size_t)blockIdx.x * 2 * kMaxUnitBytes;
    uint32_t *arr3 = arr4 + kMaxUnitBytes;

    for (uint32_t i = lane; i < (uint32_t)kSplitLists; i += 32) s_w[warp][i] = 0;
    __syncwarp();
    // pass 1: every warp counts the list members of its contiguous range of tiles
    // (tiles in front of the last one are full: their code carries no activity predicate)
    const uint32_t tfull = min(t1, ninsert / 32);
    auto count_tile = [&](uint32_t t, auto full) {
        constexpr bool kFull = decltype(full)::value;
        const uint32_t p = t * 32 + lane;
        const bool act = kFull || p < ninsert;
        const uint32_t v = act ? ldg32u(uw, lo + p) : 0;
        uint32_t h4 = ht ? (lz_hash(v, 15) << 1) : lz_hash(v, 16), h3 = lz_hash(v & 0xFFFFFFu, 15);   // level 1: 15-bit buckets (even slots)
        if (p == 0 && quirk0) { h4 = 0; h3 = 0; }
        const uint32_t q4 = h4 >> kBits4, q3 = h3 >> kBits3;
        const uint32_t m4 = same_list_mask<kQ4Bits, kFull>(q4, act), m3 = same_list_mask<kQ3Bits, kFull>(q3, act);
        if (act && (m4 & lt) == 0) s_w[warp][q4] += __popc(m4);          // group leader
        if (act && (m3 & lt) == 0) s_w[warp][kL4 + q3] += __popc(m3);
        __syncwarp();
    };
    for (uint32_t t = t0; t < max(t0, tfull); t++) count_tile(t, std::true_type{});
    for (uint32_t t = max(t0, tfull); t < t1; t++) count_tile(t, std::false_type{});
    __syncthreads();
    if (tid < kSplitLists) {
        // exclusive prefix over the warps (position order) for list `tid`
        uint32_t acc = 0;
        for (int w = 0; w < 32; w++) { uint32_t c = s_w[w][tid]; s_w[w][tid] = acc; acc += c; }
        s_start[tid] = acc;   // total
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t *ls = list_start + (size_t)blockIdx.x * kLsStride;
        uint32_t a = 0;
        for (int k = 0; k < kL4; k++) { uint32_t c = s_start[k]; s_start[k] = a; ls[k] = a; a += c; }
        ls[kL4] = a;
        a = 0;
        for (int k = 0; k < kL3; k++) { uint32_t c = s_start[kL4 + k]; s_start[kL4 + k] = a; ls[kL4 + 1 + k] = a; a += c; }
        ls[kL4 + 1 + kL3] = a;
    }
    __syncthreads();
    // pass 2: scatter (hash, position - lo) entries in position order
    auto scatter_tile = [&](uint32_t t, auto full) {
        constexpr bool kFull = decltype(full)::value;
        const uint32_t p = t * 32 + lane;
        const bool act = kFull || p < ninsert;
        const uint32_t v = act ? ldg32u(uw, lo + p) : 0;
        uint32_t h4 = ht ? (lz_hash(v, 15) << 1) : lz_hash(v, 16), h3 = lz_hash(v & 0xFFFFFFu, 15);   // level 1: 15-bit buckets (even slots)
        if (p == 0 && quirk0) { h4 = 0; h3 = 0; }
        const uint32_t q4 = h4 >> kBits4, q3 = h3 >> kBits3;
        const uint32_t m4 = same_list_mask<kQ4Bits, kFull>(q4, act), m3 = same_list_mask<kQ3Bits, kFull>(q3, act);
        uint32_t b4 = 0, b3 = 0;
        if (act) { b4 = s_w[warp][q4]; b3 = s_w[warp][kL4 + q3]; }
        __syncwarp();
        if (act) {
            arr4[s_start[q4] + b4 + __popc(m4 & lt)] = h4 | (p << 16);
            arr3[s_start[kL4 + q3] + b3 + __popc(m3 & lt)] = h3 | (p << 16);
            if ((m4 & lt) == 0) s_w[warp][q4] = b4 + __popc(m4);
            if ((m3 & lt) == 0) s_w[warp][kL4 + q3] = b3 + __popc(m3);
        }
        __syncwarp();
    };
    for (uint32_t t = t0; t < max(t0, tfull); t++) scatter_tile(t, std::true_type{});
    for (uint32_t t = max(t0, tfull); t < t1; t++) scatter_tile(t, std::false_type{});
    // positions that are never inserted carry no link
    for (uint32_t p = max(lo, uinsert) + tid; p < hi_new; p += kSplitThreads) {
        next4[(size_t)sb.u * g.m_stride + p] = 0;
        prev3[(size_t)sb.u * g.m_stride + p] = 0;
    }
    // ---- Check::update folded into this pass (the reference's second pass over the block, bgzf.rs:224-225): the
    // bytes are in L1 from the hashing above.  Every thread takes one 64-byte slice of the sub-unit's NEW bytes
    // [nb, ne) — slices are counted from the end so that only the first one is short — and the slices recombine
    // with x^(8*64*j) mod P (CRC-32) or with the Adler-32 sum formulas.  One partial sum per sub-unit; k_emit
    // folds the sub-units of a long unit.
    if (sum_part) {
        const uint32_t R = sb.ne - sb.nb;          // <= 65536 = 1024 slices
        if (tid == 0) { s_crc = 0; s_a = 0; s_b = 0; }
        if (check_kind == 0) {
            for (uint32_t i = tid; i < 1024; i += kSplitThreads) s_tab[i >> 8][i & 255] = c_crc_tab[i >> 8][i & 255];
            __syncthreads();
            uint32_t acc = 0;
            if (tid * 64 < R) {
                const uint32_t end = sb.nb + (R - 64 * tid), beg = (R - 64 * tid) >= 64 ? end - 64 : sb.nb;
                uint32_t c = ~0u, pos = beg;
                for (; pos + 4 <= end; pos += 4) {
                    c ^= ldg32u(inw, pos);
                    c = s_tab[3][c & 0xFF] ^ s_tab[2][(c >> 8) & 0xFF] ^ s_tab[1][(c >> 16) & 0xFF] ^ s_tab[0][c >> 24];
                }
                const uint8_t *in8 = (const uint8_t *)inw;
                for (; pos < end; pos++) c = (c >> 8) ^ s_tab[0][(c ^ __ldg(in8 + pos)) & 0xFF];
                c = ~c;
                acc = tid == 0 ? c : gf2_mulmod(c, g_xpow64[tid], kCrcPoly);
            }
            for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
            if (lane == 0 && acc) atomicXor(&s_crc, acc);
            __syncthreads();
            if (tid == 0) sum_part[blockIdx.x] = s_crc;
        } else {
            __syncthreads();
            unsigned long long sa = 0, sb2 = 0;
            if (tid * 64 < R) {
                const uint32_t beg = sb.nb + 64 * tid, end = min(sb.ne, beg + 64);
                const uint8_t *in8 = (const uint8_t *)inw;
                uint32_t a = 0, b = 0;
                for (uint32_t pos = beg; pos < end; pos++) { a += __ldg(in8 + pos); b += a; }
                sa = a;
                sb2 = (b + (unsigned long long)a * (sb.ne - end)) % 65521ull;
            }
            for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xFFFFFFFFu, sa, o); sb2 += __shfl_xor_sync(0xFFFFFFFFu, sb2, o); }
            if (lane == 0) { atomicAdd(&s_a, sa); atomicAdd(&s_b, sb2); }
            __syncthreads();
            if (tid == 0) {
                const uint32_t a = (uint32_t)((1ull + s_a) % 65521ull), b = (uint32_t)(((unsigned long long)R + s_b) % 65521ull);
                sum_part[blockIdx.x] = (b << 16) | a;
            }
        }
    }
}

// kMulti: units of several sub-units (the loop over them and the head sweep exist only there)
// kHash4: the hash4 lists (12 KiB of bucket state: 17 jobs per SM) or the hash3 lists (16 KiB: 13 per SM) — two launches
template <bool kMulti, bool kHash4>
__global__ void __launch_bounds__(32)
k_link(const __grid_constant__ Geo g, const uint32_t *__restrict__ lists, const uint32_t *__restrict__ list_start,
       uint16_t *__restrict__ next4, uint16_t *__restrict__ prev3, uint8_t *__restrict__ clen_g)
{
    // One job = one list of one UNIT, walked sub-unit by sub-unit with the bucket heads carried along.
    // hash4 job: u16 head[4096] + u8 cnt[4096] (occurrences so far per bucket = chain-length estimate);
    // hash3 job: u16 head[8192].  Heads hold unit positions mod 65536; a link is valid below 32768, so at every
    // sub-unit boundary the heads that fell out of the window are parked on a sentinel 32768 positions back.
    constexpr int kStateSlots = kHash4 ? (1 << kBits4) + (1 << kBits4) / 2 : (1 << kBits3);      // u16 slots of bucket state
    __shared__ __align__(16) uint16_t head[kStateSlots];
    uint8_t *cnt = (uint8_t *)(head + (1 << kBits4));
    constexpr uint32_t kJobs = kHash4 ? kL4 : kL3;
    const uint32_t u = blockIdx.x / kJobs, job = blockIdx.x % kJobs + (kHash4 ? 0 : kL4);
    const uint32_t un = g.unit_len[u], udict = g.unit_dict[u];
    if (un <= udict) return;
    const uint32_t lane = threadIdx.x, lt = lanemask_lt();
    constexpr bool is4 = kHash4;
    const uint32_t nb16 = is4 ? (2u << kBits4) / 16 : (2u << kBits3) / 16;     // uint4 words of the head table
    uint16_t *out = (is4 ? next4 : prev3) + (size_t)u * g.m_stride;
    uint8_t *clen = clen_g + (size_t)u * g.m_stride;
    {
        uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u), zero = make_uint4(0, 0, 0, 0);
        uint4 *h = (uint4 *)head, *c4 = (uint4 *)cnt;
        for (uint32_t i = lane; i < nb16; i += 32) h[i] = ones;
        if (is4) for (uint32_t i = lane; i < (1u << kBits4) / 16; i += 32) c4[i] = zero;
    }
    __syncwarp();
    for (uint32_t ksub = 0; ksub < (kMulti ? g.spu : 1u); ksub++) {
        const uint32_t ua = udict + ksub * g.seg;
        if (ksub && ua >= un) break;
        const uint32_t lo = ksub == 0 ? 0u : ua;
        const uint32_t sub = u * g.spu + ksub;
        if (kMulti && ksub) {
            // heads older than the window (and, once, the never-used ones) -> sentinel; occurrence counts decay
            const uint32_t sentinel = (lo - (uint32_t)kWindow) & 0xFFFFu;
            uint32_t *h32 = (uint32_t *)head;
            for (uint32_t i = lane; i < nb16 * 4; i += 32) {
                uint32_t w = h32[i], r = 0;
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    uint32_t hv = (w >> (16 * hf)) & 0xFFFFu;
                    const uint32_t age = (lo - hv) & 0xFFFFu;
                    if ((ksub == 1 && hv == kNone16) || age >= (uint32_t)kWindow || age == 0) hv = sentinel;
                    r |= hv << (16 * hf);
                }
                h32[i] = r;
            }
            if (is4) { uint32_t *c32 = (uint32_t *)cnt; for (uint32_t i = lane; i < (1u << kBits4) / 4; i += 32) c32[i] = (c32[i] >> 1) & 0x7F7F7F7Fu; }
            __syncwarp();
        }
        const uint32_t *ls = list_start + (size_t)sub * kLsStride;
        const uint32_t *arr = lists + (size_t)sub * 2 * kMaxUnitBytes + (is4 ? 0 : kMaxUnitBytes);
        const uint32_t beg = is4 ? ls[job] : ls[kL4 + 1 + (job - kL4)];
        const uint32_t end = is4 ? ls[job + 1] : ls[kL4 + 2 + (job - kL4)];
        constexpr int G = 8;
        uint32_t en[G];
#pragma unroll
        for (int k = 0; k < G; k++) { uint32_t i = beg + 32 * k + lane; en[k] = i < end ? __ldg(arr + i) : 0; }
        for (uint32_t base0 = beg; base0 < end; base0 += 32 * G) {
            uint32_t e[G];
#pragma unroll
            for (int k = 0; k < G; k++) e[k] = en[k];
#pragma unroll
            for (int k = 0; k < G; k++) { uint32_t i = base0 + 32 * (G + k) + lane; en[k] = i < end ? __ldg(arr + i) : 0; }
            // one tile of 32 entries; kFull: every lane has an entry (all tiles but a list's last), no activity predicates
            auto tile = [&](const uint32_t ek, const uint32_t i, auto full) {
                constexpr bool kFull = decltype(full)::value;
                const bool act = kFull || i < end;
                const uint32_t b = ek & (is4 ? ((1u << kBits4) - 1) : ((1u << kBits3) - 1)), p = lo + (ek >> 16);   // unit position
                // Entries of one tile that share a bucket are ordered by lane (= position order).  Every lane needs its
                // group (the lanes with the same bucket): the predecessor is the nearest lower member, or the bucket head
                // for the group's first member; the group's last member becomes the new head.
                // The bucket table itself names the groups: every lane reads its bucket's state, then writes its LANE id
                // there; whichever member's store survives, all members read the same 5-bit id back, and five ballots
                // turn equal ids into the group mask (ALU work instead of the ADU pipe).
                uint32_t hv = 0, c0 = 0;
                if (act) { hv = head[b]; if (is4) c0 = cnt[b]; }
                __syncwarp();
                if (act) st_relaxed_shared_u16(&head[b], lane);
                __syncwarp();
                const uint32_t grp = same_list_mask<5, kFull>(act ? ld_relaxed_shared_u16(&head[b]) : 0u, act);   // (the ballots order these reads before the stores below)
                const uint32_t lower = grp & lt;
                const uint32_t pl = __shfl_sync(0xFFFFFFFFu, p, lower ? 31 - __clz(lower) : lane);
                uint32_t dist = 0, occ = 0;
                if (act) {
                    if (lower) dist = p - pl;
                    else dist = (!kMulti || ksub == 0) ? (hv == kNone16 ? 0u : p - hv) : ((p - hv) & 0xFFFFu);
                    if (is4) occ = c0 + (uint32_t)__popc(lower);
                }
                if (act) {
                    if ((grp >> lane) == 1u) {
                        head[b] = (uint16_t)p;
                        if (is4) cnt[b] = (uint8_t)min(occ + 1u, 255u);   // occurrences so far
                    }
                    if (dist >= (uint32_t)kWindow) dist = 0;
                    out[p] = (uint16_t)dist;
                    if (is4) clen[p] = (uint8_t)min(occ, 127u);
                }
                __syncwarp();
            };
            if (base0 + 32 * G <= end) {
#pragma unroll
                for (int k = 0; k < G; k++) tile(e[k], base0 + 32 * k + lane, std::true_type{});
            } else {
#pragma unroll
                for (int k = 0; k < G; k++) {
                    if (base0 + 32 * k >= end) break;
                    tile(e[k], base0 + 32 * k + lane, std::false_type{});
                }
            }
        }
    }
}

// =============================================================================
// k_match: per-position longest-match search.  1 CTA (1024 threads) per unit.
// The unit's bytes and its next4[] chain links are staged into shared memory
// with two TMA bulk copies (cp.async.bulk + mbarrier); each thread then owns
// positions and walks its hash chain exactly like hc_matchfinder_longest_match,
// recording the best match over the first D nodes (A) and the first D/2 nodes
// (B).  Because every position is inserted, the chain of a position does not
// depend on parsing decisions, so all positions are searched in parallel.
// =============================================================================
constexpr int kMatchThreads = 1024;
// Extend a match of `len` bytes between positions p and q, 4 bytes per step: both sides step through aligned words
// with their own fixed shift, so a step costs two shared-memory loads (not four) and no address arithmetic.
__device__ __forceinline__ uint32_t lz_extend(const uint32_t *s_in, uint32_t p, uint32_t q, uint32_t len, uint32_t maxlen)
{
    const uint8_t *b = (const uint8_t *)s_in;
    uint32_t wp = (p + len) >> 2, wq = (q + len) >> 2;
    const uint32_t shp = ((p + len) & 3) * 8, shq = ((q + len) & 3) * 8;
    uint32_t lo_p = s_in[wp], lo_q = s_in[wq];
    while (len + 8 <= maxlen) {                                    // 8 bytes per trip: one exit test for two words
        const uint32_t m_p = s_in[wp + 1], m_q = s_in[wq + 1], hi_p = s_in[wp + 2], hi_q = s_in[wq + 2];
        const uint32_t x0 = __funnelshift_r(lo_p, m_p, shp) ^ __funnelshift_r(lo_q, m_q, shq);
        const uint32_t x1 = __funnelshift_r(m_p, hi_p, shp) ^ __funnelshift_r(m_q, hi_q, shq);
        if (x0 | x1) return x0 ? len + ((__ffs(x0) - 1) >> 3) : len + 4 + ((__ffs(x1) - 1) >> 3);
        len += 8; wp += 2; wq += 2; lo_p = hi_p; lo_q = hi_q;
    }
    if (len + 4 <= maxlen) {
        const uint32_t hi_p = s_in[wp + 1], hi_q = s_in[wq + 1];
        const uint32_t x = __funnelshift_r(lo_p, hi_p, shp) ^ __funnelshift_r(lo_q, hi_q, shq);
        if (x) return len + ((__ffs(x) - 1) >> 3);
        len += 4;
    }
    while (len < maxlen && b[p + len] == b[q + len]) len++;
    return len;
}

__global__ void __launch_bounds__(kMatchThreads, 1)
k_match(const __grid_constant__ Geo g, const uint16_t *__restrict__ next4g, const uint16_t *__restrict__ prev3g,
        uint64_t *__restrict__ mtab, uint32_t *__restrict__ mtab2, uint8_t *__restrict__ clen_g, uint16_t *__restrict__ order_g,
        int depth, int nice, int lazy, int ht)
{
    GZPB_DYN_SMEM(smem);
    uint32_t *s_in = (uint32_t *)smem;
    uint16_t *s_next = (uint16_t *)(smem + kInStride);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_whist[32][128];
    __shared__ uint32_t s_tot[128];
    const uint32_t tid = threadIdx.x;
    const Sub sb = sub_geometry(g, blockIdx.x);
    if (!sb.valid) return;
    const uint32_t n = sb.len;
    const uint8_t *in = g.in + (size_t)sb.u * g.in_stride + sb.h;
    uint64_t *M = mtab + (size_t)sb.u * g.m_stride + sb.h;
    uint32_t *M2 = (lazy == 2) ? mtab2 + (size_t)sb.u * g.m_stride + sb.h : nullptr;   // lazy2: depth/4 column

    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (n >= 5) {
        if (tid == 0) {
            uint32_t bin = (n + 15u) & ~15u, bnx = (n * 2 + 15u) & ~15u;
            mbar_expect_tx(&bar, bin + bnx);
            tma_load_1d(s_in, in, bin, &bar);
            tma_load_1d(s_next, next4g + (size_t)sb.u * g.m_stride + sb.h, bnx, &bar);
        }
        mbar_wait(&bar, 0);
    }
    const uint16_t *p3 = prev3g + (size_t)sb.u * g.m_stride + sb.h;
    const uint32_t depthB = (uint32_t)depth >> 1, depthC = (uint32_t)depth >> 2;

    // ---- phase 1: order the positions by chain length -------------------------
    // Chains differ wildly in length (average ~10 nodes, cap D), so 32 consecutive
    // positions walked in lock step keep only ~8/32 lanes busy.  A links-only
    // pre-walk (no byte compares) gives every position's chain length; a counting
    // sort (per-warp histograms in shared memory) then lets each warp of phase 2
    // process 32 positions of EQUAL chain length.  Ordering only changes the
    // schedule, never a result.
    uint8_t *clen = clen_g + (size_t)sb.u * g.m_stride + sb.h;
    uint16_t *order = order_g + (size_t)blockIdx.x * kMaxUnitBytes;
    const uint32_t warp = tid >> 5;
    for (uint32_t i = tid; i < 32 * 128; i += kMatchThreads) (&s_whist[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t p = sb.nb + tid; p < sb.ne; p += kMatchThreads) {
        uint32_t len = 0;
        if (n - p >= 5) {
            len = min((uint32_t)clen[p], (uint32_t)depth);   // occurrence index in the bucket (k_link): the chain-length estimate
        }
        len = min(len, 127u);
        clen[p] = (uint8_t)len;
        atomicAdd(&s_whist[warp][len], 1u);
    }
    __syncthreads();
    if (tid < 128) {
        // column `tid` (one chain length): exclusive prefix over the 32 warps
        uint32_t acc = 0;
        for (int w = 0; w < 32; w++) { uint32_t v = s_whist[w][tid]; s_whist[w][tid] = acc; acc += v; }
        s_tot[tid] = acc;
    }
    __syncthreads();
    if (tid == 0) { uint32_t acc = 0; for (int l = 127; l >= 0; l--) { uint32_t v = s_tot[l]; s_tot[l] = acc; acc += v; } }   // longest first
    __syncthreads();
    for (uint32_t p = sb.nb + tid; p < sb.ne; p += kMatchThreads) {
        uint32_t len = clen[p];
        uint32_t slot = s_tot[len] + atomicAdd(&s_whist[warp][len], 1u);
        order[slot] = (uint16_t)p;
    }
    __syncthreads();

    // ---- phase 2: the searches, 32 positions of equal chain length per warp ----
    const uint32_t npos = sb.ne - sb.nb;
    // The two global loads of a position — its slot in the sorted order and its hash3 link, a scattered 2-byte load that
    // misses L1 — are issued one position ahead (ncu, profiles/r2w: 22 % of k_match's warp time was long-scoreboard stall):
    // the next position's order entry at the top of this position's walk, its prev3 link right after the walk.
    uint32_t p_cur = tid < npos ? order[tid] : 0u;
    uint32_t d3_cur = (tid < npos && !ht) ? p3[p_cur] : 1u;
    for (uint32_t i = tid; i < npos; i += kMatchThreads) {
        const uint32_t p = p_cur;
        const uint32_t d3 = d3_cur;            // level 1 (ht_matchfinder) has no hash3 table: the parser must see "bucket usable, no 3-byte match"
        const bool have_next = i + kMatchThreads < npos;
        const uint32_t p_next = have_next ? order[i + kMatchThreads] : 0u;
        const uint32_t maxlen = min((uint32_t)kMaxMatch, n - p);
        if (maxlen < 5) { M[p] = 0; if (M2) M2[p] = 0; p_cur = p_next; d3_cur = (have_next && !ht) ? p3[p_next] : 1u; continue; }
        const uint32_t nicep = min((uint32_t)nice, maxlen);
        // bytes 0..3 (the quick reject) and 4..11 (the inline extension) of this position: four words, three funnel shifts
        const uint32_t pw = p >> 2, psh = (p & 3) * 8;
        const uint32_t pa1 = s_in[pw + 1], pa2 = s_in[pw + 2];
        const uint32_t seq4 = __funnelshift_r(s_in[pw], pa1, psh);
        const uint32_t w1 = __funnelshift_r(pa1, pa2, psh), w2 = __funnelshift_r(pa2, s_in[pw + 3], psh);

        uint32_t best = 3, boff = 0, lenB = 0, offB = 0, lenC = 0, offC = 0;
        bool haveB = !lazy, haveC = (lazy != 2);
        uint32_t q = p, visited = 0;
        uint32_t pbest = 0;                    // the byte of p a longer match must reach: b[p + best], kept in a register
        const uint8_t *b = (const uint8_t *)s_in;
        // the visit count at which the next snapshot (depth/4, depth/2) or the depth limit falls: one compare per node
        uint32_t snap = (!haveC && depthC) ? depthC : (!haveB && depthB) ? depthB : (uint32_t)depth;
        uint32_t d = s_next[q];
        for (;;) {
            if (d == 0) break;
            q -= d;
            if (p - q >= (uint32_t)kWindow) break;
            visited++;
            d = s_next[q];                     // the next link leaves now: its latency hides behind this node's compares
            // the bytes at q, q + 4, q + 8 share their word offset and their shift: every word is loaded once
            const uint32_t wq = q >> 2, sh = (q & 3) * 8;
            uint32_t a1 = 0;
            bool cand = (best == 3) || (b[q + best] == pbest);
            if (cand) { a1 = s_in[wq + 1]; cand = __funnelshift_r(s_in[wq], a1, sh) == seq4; }
            if (cand) {
                // most matches are short: first 8 bytes of the extension inline, the rest in lz_extend
                uint32_t len;
                const uint32_t a2 = s_in[wq + 2];
                uint32_t x = __funnelshift_r(a1, a2, sh) ^ w1;
                if (x) len = 4 + ((__ffs(x) - 1) >> 3);
                else {
                    x = __funnelshift_r(a2, s_in[wq + 3], sh) ^ w2;
                    if (x) len = 8 + ((__ffs(x) - 1) >> 3);
                    else len = (maxlen > 12) ? lz_extend(s_in, p, q, 12, maxlen) : 12;
                }
                len = min(len, maxlen);
                if (len > best) {
                    best = len; boff = p - q;
                    if (len >= nicep) break;
                    pbest = b[p + best];
                }
            }
            if (visited == snap) {
                if (!haveC && visited == depthC) { haveC = true; lenC = best > 3 ? best : 0; offC = boff; }
                if (!haveB && visited == depthB) { haveB = true; lenB = best > 3 ? best : 0; offB = boff; }
                if (visited == (uint32_t)depth) break;
                snap = (!haveC && depthC > visited) ? depthC : (!haveB && depthB > visited) ? depthB : (uint32_t)depth;
            }
        }
        p_cur = p_next; d3_cur = (have_next && !ht) ? p3[p_next] : 1u;      // (the next position's hash3 link leaves now)
        uint32_t off3 = 0;
        if (!ht && d3 && d3 <= 8192u && ((ld32u(s_in, p - d3) ^ seq4) & 0xFFFFFFu) == 0) off3 = d3;
        uint32_t lenA = best > 3 ? best : 0;
        if (!haveB) { lenB = lenA; offB = boff; }
        if (M2) { if (!haveC) { lenC = lenA; offC = boff; } M2[p] = (lenC ? lenC - 3 : 0) | (offC << 8); }
        if (!lazy) { lenB = 0; offB = 0; }
        M[p] = pack_entry(lenA, boff, lenB, offB, d3 != 0, off3);
    }
}

// =============================================================================
// k_emit: parse + Huffman + bit packing + container.  1 CTA (one warp) per unit, 32 CTAs per SM.  The warp runs the
// reference's parser as a windowed parse over match-table tiles streamed into a 3-slot shared-memory ring by TMA bulk
// copies (every lane evaluates the parser's two states at its position, the warp follows the path through the window
// and stops exactly at state-changing events: min_len re-calculation, block-split checks, sequence-store limit); at
// each DEFLATE block boundary it builds the Huffman codes (libdeflate's sort / in-place tree / length limiting /
// canonical codewords restated), picks the cheapest of dynamic / static / stored, and packs the tokens in parallel
// (prefix scan of code lengths, OR-merge in a shared staging buffer, coalesced 32-bit stores).
// =============================================================================
constexpr int kEmitThreads = 32;
constexpr int kTile = 128;                 // positions per streamed tile
constexpr int kRing = 3;                   // tiles resident in the shared-memory ring
constexpr int kEmitUnitsPerSM = 32;        // 7 KiB of shared memory and 64 registers per unit: the CTA limit of an SM
constexpr int kTokPerThread = 8;
constexpr int kChunkTok = kEmitThreads * kTokPerThread;
constexpr int kStageWords = (kChunkTok * 48) / 32 + 8;

struct EmitShared {
    alignas(16) uint64_t mt[kRing][kTile];
    alignas(16) uint8_t inb[kRing][kTile];
    alignas(8) uint64_t bar[kRing];
    uint32_t fl[kNumLitlen];
    uint32_t fo[kNumOffset];
    uint32_t obs[10], new_obs[10];
    uint8_t lens[kNumLitlen + kNumOffset];   // litlen then offset lens (contiguous like libdeflate)
    uint16_t lcw[kNumLitlen];
    uint16_t ocw[kNumOffset];
    uint8_t plen[kNumPrecode];
    uint16_t pcw[kNumPrecode];
    uint32_t pfreq[kNumPrecode];
    // Scratch that is only live while a DEFLATE block is being finished borrows memory that is idle then (32 instead of
    // 22 resident units per SM): the Huffman sort array and the bit-packing staging words live in the tile ring (the
    // parser drains it at the block end and issues its tiles again afterwards: a few tiles per ~50 KB block), the
    // precode items in the literal/length frequencies (dead once the symbol costs are summed).
    static_assert(sizeof(uint64_t) * kRing * kTile >= sizeof(uint32_t) * kStageWords && kStageWords >= kNumLitlen, "scratch must fit the tile ring");
    static_assert(sizeof(uint32_t) * kNumLitlen >= sizeof(uint16_t) * (kNumLitlen + kNumOffset), "items must fit the frequencies");
    __device__ __forceinline__ uint32_t *Ap() { return (uint32_t *)&mt[0][0]; }
    __device__ __forceinline__ uint32_t *stagep() { return (uint32_t *)&mt[0][0]; }
    __device__ __forceinline__ uint16_t *itemsp() { return (uint16_t *)fl; }
    uint32_t scan[kEmitThreads / 32];
    uint32_t used[8];
    // control block written by thread 0
    uint32_t blk_begin, blk_end, ntok, is_final, min_len;
    uint32_t nused, nitems, nlit, noff, nexpl, btype;
    uint32_t cost_dyn, cost_static;
    uint32_t G;        // bit position in the payload
    uint32_t carry;    // partial word at G>>5
    int32_t status;
};

__device__ __forceinline__ uint32_t bsr32(uint32_t v) { return 31 - __clz(v); }

__device__ __forceinline__ void len_slot(uint32_t len, uint32_t &slot, uint32_t &ebits, uint32_t &eval)
{
    uint32_t l = len - 3;
    if (l < 8) { slot = l; ebits = 0; eval = 0; }
    else if (l == 255) { slot = 28; ebits = 0; eval = 0; }
    else {
        uint32_t k = 31 - __clz(l);
        slot = 4 * k - 4 + ((l >> (k - 2)) & 3);
        ebits = k - 2; eval = l & ((1u << (k - 2)) - 1);
    }
}
__device__ __forceinline__ void off_slot(uint32_t off, uint32_t &slot, uint32_t &ebits, uint32_t &eval)
{
    uint32_t d = off - 1;
    if (d < 4) { slot = d; ebits = 0; eval = 0; }
    else {
        uint32_t k = 31 - __clz(d);
        slot = 2 * k + ((d >> (k - 1)) & 1);
        ebits = k - 1; eval = d & ((1u << (k - 1)) - 1);
    }
}
__device__ __forceinline__ uint32_t len_slot_only(uint32_t len) { uint32_t s, a, b; len_slot(len, s, a, b); return s; }
__device__ __forceinline__ uint32_t off_slot_only(uint32_t off) { uint32_t s, a, b; off_slot(off, s, a, b); return s; }

__device__ __forceinline__ void stage_put(uint32_t *st, uint32_t bitpos, uint64_t bits, uint32_t nbits)
{
    if (nbits == 0) return;
    uint32_t w = bitpos >> 5, s = bitpos & 31;
    uint64_t lo = bits << s;
    uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32);
    uint32_t w2 = s ? (uint32_t)(bits >> (64 - s)) : 0;
    if (w0) atomicOr(&st[w], w0);
    if (w1) atomicOr(&st[w + 1], w1);
    if (w2) atomicOr(&st[w + 2], w2);
}

__device__ __forceinline__ uint32_t choose_min_match_len(uint32_t num_used, uint32_t depth)
{
    if (num_used >= 80) return 3;
    uint32_t m = c_min_lens[num_used];
    if (depth < 16) {
        uint32_t cap = depth < 5 ? 4 : depth < 10 ? 5 : 7;
        m = min(m, cap);
    }
    return m;
}

// Cooperative restatement of libdeflate's deflate_make_huffman_code().
// freqs/lens/cw live in shared memory; all kEmitThreads threads call it.
template <typename CW>
__device__ void make_huffman_code(EmitShared &S, const uint32_t *freqs, uint32_t num_syms, uint32_t max_len,
                                  uint8_t *lens, CW *cw)
{
    const uint32_t tid = threadIdx.x;
    if (tid == 0) S.nused = 0;
    __syncthreads();
    // rank sort by (freq, sym)
    for (uint32_t s = tid; s < num_syms; s += kEmitThreads) {
        uint32_t f = freqs[s];
        lens[s] = 0; cw[s] = 0;
        if (f) {
            uint32_t key = s | (f << 10), rank = 0;
            for (uint32_t t = 0; t < num_syms; t++) {
                uint32_t ft = freqs[t];
                rank += (ft && (t | (ft << 10)) < key);
            }
            S.Ap()[rank] = key;
            atomicAdd(&S.nused, 1u);
        }
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t *A = S.Ap();
        const uint32_t n = S.nused;
        if (n < 2) {
            uint32_t sym = n ? (A[0] & 1023u) : 0;
            uint32_t nz = sym ? sym : 1;
            cw[0] = 0; lens[0] = 1; cw[nz] = 1; lens[nz] = 1;
        } else {
            const uint32_t FM = ~1023u;
            const uint32_t last = n - 1;
            uint32_t i = 0, b = 0, e = 0;
            do {
                uint32_t nf;
                if (i + 1 <= last && (b == e || (A[i + 1] & FM) <= (A[b] & FM))) {
                    nf = (A[i] & FM) + (A[i + 1] & FM); i += 2;
                } else if (b + 2 <= e && (i > last || (A[b + 1] & FM) < (A[i] & FM))) {
                    nf = (A[b] & FM) + (A[b + 1] & FM);
                    A[b] = (e << 10) | (A[b] & 1023u);
                    A[b + 1] = (e << 10) | (A[b + 1] & 1023u);
                    b += 2;
                } else {
                    nf = (A[i] & FM) + (A[b] & FM);
                    A[b] = (e << 10) | (A[b] & 1023u);
                    i++; b++;
                }
                A[e] = nf | (A[e] & 1023u);
            } while (++e < last);

            uint32_t len_counts[16];
            for (uint32_t l = 0; l <= max_len; l++) len_counts[l] = 0;
            len_counts[1] = 2;
            int root = (int)n - 2;
            A[root] &= 1023u;
            for (int node = root - 1; node >= 0; node--) {
                uint32_t parent = A[node] >> 10;
                uint32_t dpt = (A[parent] >> 10) + 1;
                A[node] = (A[node] & 1023u) | (dpt << 10);
                if (dpt >= max_len) {
                    dpt = max_len;
                    do { dpt--; } while (len_counts[dpt] == 0);
                }
                len_counts[dpt]--;
                len_counts[dpt + 1] += 2;
            }
            uint32_t k = 0;
            for (uint32_t len = max_len; len >= 1; len--) {
                uint32_t c = len_counts[len];
                while (c--) lens[A[k++] & 1023u] = (uint8_t)len;
            }
            uint32_t next_cw[16];
            next_cw[0] = 0; next_cw[1] = 0;
            for (uint32_t len = 2; len <= max_len; len++) next_cw[len] = (next_cw[len - 1] + len_counts[len - 1]) << 1;
            for (uint32_t s = 0; s < num_syms; s++) {
                uint32_t l = lens[s];
                if (l) cw[s] = (CW)(__brev(next_cw[l]++) >> (32 - l));
            }
        }
    }
    __syncthreads();
}

// flush `nbits` staged bits starting at S.G to the payload; all threads call it.
__device__ __forceinline__ void stage_commit(EmitShared &S, uint32_t *__restrict__ payload, uint32_t nbits)
{
    __syncthreads();
    const uint32_t g = S.G, bw = g >> 5, nfull = ((g & 31) + nbits) >> 5;
    for (uint32_t i = threadIdx.x; i < nfull; i += kEmitThreads) payload[bw + i] = S.stagep()[i];
    __syncthreads();
    if (threadIdx.x == 0) { S.carry = S.stagep()[nfull]; S.G = g + nbits; }
    __syncthreads();
}
// zero the staging area and seed word 0 with the carry; all threads call it.
__device__ __forceinline__ void stage_begin(EmitShared &S, uint32_t words)
{
    for (uint32_t i = threadIdx.x; i < words; i += kEmitThreads) S.stagep()[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) S.stagep()[0] = S.carry;
    __syncthreads();
}

// Match-table / input tiles streamed into a kRing-slot shared-memory ring by TMA
// bulk copies.  All methods are warp-collective (called by every lane of warp 0
// with uniform arguments); lane 0 issues the copies.
struct Parser {
    const uint64_t *mt_g;
    const uint8_t *in_g;
    EmitShared *S;
    uint32_t n, ntiles, issued, ready;
    uint32_t phase;    // bit s: the parity slot s's barrier completes next (every issued tile is waited for exactly once)
    __device__ __forceinline__ void issue(uint32_t t)
    {
        uint32_t slot = t % kRing, pos = t * kTile;
        uint32_t cnt = min((uint32_t)kTile, n - pos);
        uint32_t bm = (cnt * 8 + 15u) & ~15u, bi = (cnt + 15u) & ~15u;
        mbar_expect_tx(&S->bar[slot], bm + bi);
        tma_load_1d(S->mt[slot], mt_g + pos, bm, &S->bar[slot]);
        tma_load_1d(S->inb[slot], in_g + pos, bi, &S->bar[slot]);
    }
    __device__ __forceinline__ void wait_next()
    {
        const uint32_t slot = ready % kRing;
        mbar_wait(&S->bar[slot], (phase >> slot) & 1u);
        phase ^= 1u << slot;
        ready++;
    }
    // tiles below p's tile are dead: refill their slots
    __device__ __forceinline__ void advance(uint32_t p)
    {
        uint32_t want = min(ntiles, p / kTile + kRing);
        if (issued < want) {
            // tiles a long match jumped over were issued but never consumed: observe their
            // completion before their slots (and mbarrier phases) are reused
            const uint32_t first = p / kTile;
            while (ready < issued && ready < first) wait_next();
            if (issued < first) issued = ready = first;     // tiles a long match jumped over entirely are never loaded
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {
                fence_proxy_async();
                for (uint32_t t = issued; t < want; t++) issue(t);
            }
            issued = want;
        }
    }
    __device__ __forceinline__ void need(uint32_t p)
    {
        uint32_t t = p / kTile;
        while (ready <= t) wait_next();
    }
    // the ring memory is about to be used as scratch: nothing may be in flight
    __device__ __forceinline__ void drain()
    {
        while (ready < issued) wait_next();
        __syncwarp();
    }
    // the scratch use is over: the tiles from p's on are issued again by the next advance()
    __device__ __forceinline__ void restart(uint32_t p) { issued = ready = p / kTile; }
    __device__ __forceinline__ uint64_t M(uint32_t p) { return S->mt[(p / kTile) % kRing][p % kTile]; }
    __device__ __forceinline__ uint32_t B(uint32_t p) { return S->inb[(p / kTile) % kRing][p % kTile]; }
};

// hc_matchfinder_longest_match() answered from the match table
__device__ __forceinline__ void table_search(uint64_t e, uint32_t b, bool useB, uint32_t maxlen, uint32_t &len, uint32_t &off)
{
    len = b; off = 0;
    if (maxlen < 5) return;
    uint32_t lx = useB ? (uint32_t)(e >> 23) & 0xFF : (uint32_t)e & 0xFF;
    uint32_t ox = useB ? (uint32_t)(e >> 31) & 0x7FFF : (uint32_t)(e >> 8) & 0x7FFF;
    if (lx) lx += 3;
    if (b < 4) {
        if (!((e >> 46) & 1)) return;
        uint32_t off3 = (uint32_t)(e >> 47) & 0x3FFF;
        if (b < 3 && off3) { len = 3; off = off3; }
        if (lx) { len = lx; off = ox; }
    } else if (lx > b) { len = lx; off = ox; }
}

__device__ __forceinline__ uint32_t adler_combine_dev(uint32_t a1, uint32_t a2, uint64_t len2);

// Check::combine over the sub-units of one unit (k_split left one partial sum per sub-unit); warp-collective.
__device__ __forceinline__ uint32_t fold_unit_sum(const Geo &g, uint32_t u, const uint32_t *__restrict__ part, int kind, uint32_t lane)
{
    if (g.spu == 1) return part[u];
    const uint32_t n = g.unit_len[u], dict = g.unit_dict[u];
    if (kind == 0) {
        uint32_t acc = 0;
        for (uint32_t k = lane; k < g.spu; k += 32) {
            const uint32_t a = dict + k * g.seg;
            if (a >= n) break;
            const uint32_t b = min(a + g.seg, n);
            const uint32_t c = part[(size_t)u * g.spu + k];
            acc ^= (b == n) ? c : gf2_mulmod(c, gf2_xpow8((uint64_t)(n - b), kCrcPoly), kCrcPoly);   // x^(8 * bytes behind sub-unit k)
        }
        for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        return acc;
    }
    uint32_t s = 1u;
    if (lane == 0)
        for (uint32_t k = 0; k < g.spu; k++) {
            const uint32_t a = dict + k * g.seg;
            if (a >= n) break;
            s = adler_combine_dev(s, part[(size_t)u * g.spu + k], (uint64_t)(min(a + g.seg, n) - a));
        }
    return __shfl_sync(0xFFFFFFFFu, s, 0);
}

template <bool kLazy2>     // levels 8-9: the parser looks two positions ahead (states HC / X below); compiled out otherwise
__global__ void __launch_bounds__(kEmitThreads, kEmitUnitsPerSM)
k_emit(const __grid_constant__ Geo g, const uint32_t *__restrict__ unit_flags,
       const uint64_t *__restrict__ mtab, const uint32_t *__restrict__ mtab2, uint32_t *__restrict__ crc_io, const uint32_t *__restrict__ sum_part, int check_kind,
       uint32_t *__restrict__ tok_base,
       uint8_t *__restrict__ out_base, uint32_t *__restrict__ out_len, int32_t *__restrict__ out_status,
       int mode, int depth, int nice, int level, int format)
{
    __shared__ EmitShared S;
    const uint32_t u = blockIdx.x, tid = threadIdx.x;
    const uint32_t n = g.unit_len[u];          // dictionary + data
    const uint32_t dict = g.unit_dict[u];      // preset dictionary in front of the data (never emitted)
    const uint32_t dl = n - dict;              // bytes to encode
    const uint32_t flags = unit_flags[u];
    const uint8_t *in = g.in + (size_t)u * g.in_stride;
    // the unit's Check::sum: folded from k_split's per-sub-unit partial sums (written back for the host / the batch
    // combine), or what k_check left (level 0 has no k_split pass)
    uint32_t unit_sum = 0;
    if (check_kind >= 0) {
        if (sum_part) { unit_sum = fold_unit_sum(g, u, sum_part, check_kind, tid); if (tid == 0) crc_io[u] = unit_sum; }
        else unit_sum = crc_io[u];
    }
    uint32_t *tok = tok_base + (size_t)u * g.tok_stride;
    uint8_t *slot = out_base + (size_t)u * g.out_stride;
    uint32_t *payload = (uint32_t *)(slot + kOutPayloadOff);
    const bool sync_flush = (flags & 2u) != 0;    // Gzip/Zlib non-last and RawDeflate: no BFINAL, sync marker
    const bool final_block = !sync_flush;
    // level 1 = deflate_compress_fastest: greedy over the depth-2 table, min match 4, no split statistics,
    // blocks end at 65535 bytes or 8192 matches
    const bool fast = (level == 1);
    const uint32_t soft_max = fast ? (uint32_t)kFastSoftMaxBlockLength : (uint32_t)kSoftMaxBlockLength;
    const uint32_t seq_limit = fast ? (uint32_t)kFastSeqStoreLength : (uint32_t)kSeqStoreLength;

    Parser P;
    P.mt_g = mtab + (size_t)u * g.m_stride; P.in_g = in; P.S = &S; P.n = n;
    P.ntiles = (n + kTile - 1) / kTile; P.issued = dict / kTile; P.ready = dict / kTile; P.phase = 0;

    if (tid == 0) {
        for (int i = 0; i < kRing; i++) mbar_init(&S.bar[i], 1);
        fence_mbar_init();
        S.G = 0; S.carry = 0; S.status = 0;
    }
    __syncthreads();

    long long t_prev = clock64();
#define PHASE(idx_) do { if (tid == 0) { long long t_ = clock64(); atomicAdd(&g_phase[idx_], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
    const uint32_t passthrough = (level == 0) ? 0xFFFFFFFFu : (uint32_t)(55 - level * 4);
    if (dl <= passthrough && !(sync_flush && dl == 0)) {
        // deflate_compress_none(): stored blocks of <= 65535 bytes
        uint32_t pos = dict;
        do {
            uint32_t len = min(n - pos, 65535u);
            uint32_t bfinal = (n - pos <= 65535u) ? (final_block ? 1u : 0u) : 0u;
            uint32_t tot_bits = 8 + 32 + 8 * len;
            for (uint32_t cb = 0; cb < tot_bits;) {
                uint32_t chunk = min(tot_bits - cb, (uint32_t)(kStageWords - 4) * 32);
                stage_begin(S, kStageWords);
                uint32_t g0 = S.G & 31;
                for (uint32_t bit = cb + tid * 8; bit < cb + chunk; bit += kEmitThreads * 8) {
                    uint32_t v;
                    if (bit == 0) v = bfinal;
                    else if (bit < 24) v = (len >> (bit - 8)) & 0xFF;
                    else if (bit < 40) v = ((~len & 0xFFFF) >> (bit - 24)) & 0xFF;
                    else v = in[pos + (bit - 40) / 8];
                    stage_put(S.stagep(), g0 + (bit - cb), v, 8);
                }
                stage_commit(S, payload, chunk);
                cb += chunk;
            }
            pos += len;
        } while (pos != n);
    } else if (dl > 0) {
        uint32_t p = dict;           // parser position (warp 0 is authoritative)
        uint32_t next_recalc = 0, min_len = 3;
        while (true) {
            // ---------------- block start (all threads) ----------------
            if (tid == 0) S.blk_begin = p;
            __syncthreads();
            const uint32_t bb = S.blk_begin;
            if (bb >= n) break;
            const uint32_t max_block_end = (n - bb < soft_max + (uint32_t)kMinBlockLength) ? n : bb + soft_max;
            for (uint32_t i = tid; i < kNumLitlen; i += kEmitThreads) S.fl[i] = 0;
            if (tid < kNumOffset) S.fo[tid] = 0;
            if (tid < 10) { S.obs[tid] = 0; S.new_obs[tid] = 0; }
            if (tid < 8) S.used[tid] = 0;
            __syncthreads();
            {   // calculate_min_match_len(): distinct byte values in the first <= 4096 bytes
                uint32_t span = min(max_block_end - bb, 4096u);
                if (max_block_end - bb >= 512) {
                    for (uint32_t i = tid; i < span; i += kEmitThreads) { uint32_t c = in[bb + i]; atomicOr(&S.used[c >> 5], 1u << (c & 31)); }
                }
                __syncthreads();
            }
            PHASE(0);
            // ---------------- windowed parse (warp 0) ----------------
            // Each lane evaluates one main-loop iteration of the reference parser
            // ("step") as if the parser were in its fresh state at P0+lane; the
            // warp then follows the actual path through the window, commits the
            // tokens of the steps on the path and stops exactly at the events that
            // change parser state (min_len recalculation, block-split checks).
            if (tid < 32) {
                const uint32_t lane = tid;
                if (fast) min_len = 4;
                else if (max_block_end - bb < 512) min_len = 3;
                else {
                    uint32_t nu = 0;
                    for (int i = 0; i < 8; i++) nu += __popc(S.used[i]);
                    min_len = choose_min_match_len(nu, depth);
                }
                next_recalc = bb + min(n - bb, 10000u);
                uint32_t ntok = 0, nmatch = 0, num_obs = 0, num_new_obs = 0, in_h = 0;
                bool end_block = false;
                const uint32_t *M2 = kLazy2 ? mtab2 + (size_t)u * g.m_stride : nullptr;   // lazy2: depth/4 column
                do {
                    P.advance(p);
                    P.need(min(n - 1, p + 33 + (kLazy2 ? 1u : 0u)));
                    // ---- per-lane transitions at q = p + lane ----
                    // F(q)  : parser in its fresh state at q (top of the reference's main loop).
                    // HB(q) : parser at `have_cur_match` at q, the current match being the depth/2 search result at q
                    //         (reached through the look-ahead at the next position).
                    // HC(q) : the same with the depth/4 result at q (lazy2: reached through the look-ahead two positions on).
                    // X(q)  : lazy2, second of the two literals in front of an HC match: literal at q, then HC(q + 1).
                    // A transition emits ONE token and names the next state:
                    //   word = advance (bits 0-8) | next state (bits 9-10: 0 F, 1 HB, 2 HC, 3 X) | token-is-match (bit 11)
                    constexpr uint32_t kToHB = 1u | (1u << 9), kToHC = 1u | (2u << 9), kToX = 1u | (3u << 9), kIsM = 1u << 11;
                    const uint32_t q = p + lane;
                    uint32_t wF = 1, wHB = 1, wHC = 1, lenF = 0, offF = 0, lenHB = 0, offHB = 0, lenHC = 0, offHC = 0;
                    if (q < max_block_end) {
                        const uint64_t e0 = P.M(q);
                        const uint32_t maxlen = min((uint32_t)kMaxMatch, n - q);
                        const uint32_t maxlen1 = (q + 1 < n) ? min((uint32_t)kMaxMatch, n - (q + 1)) : 0u;
                        const uint64_t e1 = maxlen1 >= 5 ? P.M(q + 1) : 0ull;
                        const uint32_t nice_q = min((uint32_t)nice, maxlen);
                        const uint32_t maxlen2 = (kLazy2 && q + 2 < n) ? min((uint32_t)kMaxMatch, n - (q + 2)) : 0u;
                        const uint64_t e2 = (kLazy2 && maxlen2 >= 5) ? P.M(q + 2) : 0ull;
                        const uint32_t c2 = (kLazy2 && maxlen2 >= 5) ? M2[q + 2] : 0u;
                        // decide(): with the current match (cl, co) at q — emit it, or a literal and move on to the better
                        // match one (lazy) or two (lazy2) positions ahead
                        auto decide = [&](uint32_t cl, uint32_t co) -> uint32_t {
                            if (cl >= nice_q) return cl | kIsM;
                            uint32_t nl, no;
                            table_search(e1, cl - 1, true, maxlen1, nl, no);
                            if (nl >= cl && 4 * (int)(nl - cl) + ((int)bsr32(co) - (int)bsr32(no)) > 2) return kToHB;
                            if (kLazy2) {
                                // longest_match(q + 2, cl - 1, depth >> 2) answered from the depth/4 column
                                nl = cl - 1; no = 0;
                                if (maxlen2 >= 5) {
                                    uint32_t lx = c2 & 0xFF, ox = (c2 >> 8) & 0x7FFF;
                                    const uint32_t bl = cl - 1;
                                    if (lx) lx += 3;
                                    if (bl < 4) {
                                        if ((e2 >> 46) & 1) {
                                            const uint32_t off3 = (uint32_t)(e2 >> 47) & 0x3FFF;
                                            if (bl < 3 && off3) { nl = 3; no = off3; }
                                            if (lx) { nl = lx; no = ox; }
                                        }
                                    } else if (lx > bl) { nl = lx; no = ox; }
                                }
                                if (nl >= cl && 4 * (int)(nl - cl) + ((int)bsr32(co) - (int)bsr32(no)) > 6) return kToX;
                            }
                            return cl | kIsM;
                        };
                        uint32_t cl, co;
                        table_search(e0, min_len - 1, false, maxlen, cl, co);
                        if (mode == 0) {
                            if (cl >= min_len && (cl > 3 || co <= 4096)) { lenF = cl; offF = co; wF = cl | kIsM; }
                        } else {
                            if (!(cl < min_len || (cl == 3 && co > 8192))) { lenF = cl; offF = co; wF = decide(cl, co); }
                            const uint32_t off3q = (uint32_t)(e0 >> 47) & 0x3FFF;
                            // HB(q): current match = depth/2 result at q (or the 3-byte match)
                            uint32_t hl = (uint32_t)(e0 >> 23) & 0xFF, ho = (uint32_t)(e0 >> 31) & 0x7FFF;
                            if (hl) hl += 3;
                            else { ho = off3q; hl = ho ? 3 : 0; }
                            if (hl && maxlen >= 5) { lenHB = hl; offHB = ho; wHB = decide(hl, ho); }
                            if (kLazy2) {
                                // HC(q): current match = depth/4 result at q (or the 3-byte match)
                                const uint32_t c0 = maxlen >= 5 ? M2[q] : 0u;
                                hl = c0 & 0xFF; ho = (c0 >> 8) & 0x7FFF;
                                if (hl) hl += 3;
                                else { ho = off3q; hl = ho ? 3 : 0; }
                                if (hl && maxlen >= 5) { lenHC = hl; offHC = ho; wHC = decide(hl, ho); }
                            }
                        }
                    }
                    // ---- follow the path through the window (one token per hop) ----
                    // (every lane walks the same path: the F and HB words of a position travel in one shuffle)
                    const uint32_t wlimit = min(32u, max_block_end - p);
                    uint32_t vis = 0, c = 0, st = in_h, myst;
                    if (!kLazy2) {
                        // ten instructions per hop: the state travels as the shift (0 / 16) that selects its transition word,
                        // the lane a hop lands on notes the state itself, one ballot afterwards gives the path
                        const uint32_t w2 = wF | (wHB << 16);
                        uint32_t sh = in_h << 4, mine = 0xFFu;
                        while (c < wlimit) {
                            const uint32_t a = __shfl_sync(0xFFFFFFFFu, w2, c);
                            mine = (lane == c) ? sh : mine;
                            const uint32_t w = a >> sh;
                            c += w & 0x1FF; sh = (w >> 5) & 16u;
                        }
                        vis = __ballot_sync(0xFFFFFFFFu, mine != 0xFFu);
                        myst = (mine >> 4) & 1u; st = sh >> 4;
                    } else
                    {
                        const uint32_t w2 = wF | (wHB << 12);
                        uint32_t sm0 = 0, sm1 = 0;
                        while (c < wlimit) {
                            vis |= 1u << c; sm0 |= (kLazy2 ? (st & 1u) : st) << c;
                            const uint32_t a = __shfl_sync(0xFFFFFFFFu, w2, c);
                            uint32_t w = (st & 1u) ? (a >> 12) : a;
                            if (kLazy2) {
                                sm1 |= (st >> 1) << c;
                                const uint32_t b = __shfl_sync(0xFFFFFFFFu, wHC, c);
                                if (st == 2u) w = b; else if (st == 3u) w = kToHC;
                            }
                            c += w & 0x1FF; st = (w >> 9) & 3u;
                        }
                        myst = ((sm0 >> lane) & 1u) | (kLazy2 ? (((sm1 >> lane) & 1u) << 1) : 0u);
                    }
                    const bool onpath = (vis >> lane) & 1u;
                    const uint32_t myw = kLazy2 ? (myst == 0 ? wF : myst == 1 ? wHB : myst == 2 ? wHC : kToHC) : (myst ? wHB : wF);
                    const bool myM = (myw & kIsM) != 0;                             // my token is a match
                    const uint32_t incl = __popc(vis & (lane == 31 ? 0xFFFFFFFFu : ((2u << lane) - 1)));   // tokens up to and including mine
                    const uint32_t e_l = q + (myw & 0x1FF);
                    const bool ends_iter = ((myw >> 9) & 3u) == 0;     // next state is F: a main-loop iteration ends here
                    // ---- events ----
                    // (cheap uniform pre-test: most windows cannot contain any event)
                    uint32_t commit_mask = vis, next_p = p + c, next_h = st;
                    int event = 0;   // 1 = recalc before lane Lr, 2 = block check after lane Lc, 3 = sequence store full after lane Ls
                    uint32_t mmask = __ballot_sync(0xFFFFFFFFu, onpath && myM);
                    const bool may_check = !fast && (num_new_obs + 32 >= (uint32_t)kObsPerCheck) && (p + 32 + 258 - bb >= (uint32_t)kMinBlockLength) && (n - p > (uint32_t)kMinBlockLength);
                    const bool may_recalc = (mode != 0) && (p + 32 > next_recalc);
                    const bool may_seq = nmatch + 32 >= seq_limit;
                    if (may_check || may_recalc || may_seq) {
                        uint32_t rmask = may_recalc ? __ballot_sync(0xFFFFFFFFu, onpath && myst == 0 && q >= next_recalc) : 0u;
                        uint32_t cmask = __ballot_sync(0xFFFFFFFFu, !fast && onpath && ends_iter && (num_new_obs + incl >= (uint32_t)kObsPerCheck) &&
                                                                         (e_l - bb >= (uint32_t)kMinBlockLength) && (n - e_l >= (uint32_t)kMinBlockLength));
                        // sequence store full (SEQ_STORE_LENGTH matches in this DEFLATE block): the block ends
                        const uint32_t mincl = __popc(mmask & (lane == 31 ? 0xFFFFFFFFu : ((2u << lane) - 1)));
                        uint32_t smask = __ballot_sync(0xFFFFFFFFu, onpath && ends_iter && (nmatch + mincl >= seq_limit));
                        int Lr = rmask ? __ffs(rmask) - 1 : 64, Lc = cmask ? __ffs(cmask) - 1 : 64, Ls = smask ? __ffs(smask) - 1 : 64;
                        if (Lr <= Lc && Lr <= Ls && Lr < 64) { event = 1; commit_mask = vis & ((1u << Lr) - 1); next_p = p + Lr; next_h = 0; }
                        else if (Ls <= Lc && Ls < 64) { event = 3; commit_mask = vis & (Ls == 31 ? 0xFFFFFFFFu : ((2u << Ls) - 1)); next_p = __shfl_sync(0xFFFFFFFFu, e_l, Ls); next_h = 0; }
                        else if (Lc < 64) { event = 2; commit_mask = vis & (Lc == 31 ? 0xFFFFFFFFu : ((2u << Lc) - 1)); next_p = __shfl_sync(0xFFFFFFFFu, e_l, Lc); next_h = 0; }
                    }
                    // ---- commit ----
                    if ((commit_mask >> lane) & 1u) {
                        const uint32_t ti = ntok + incl - 1;
                        const bool isM = myM;
                        const uint32_t mlen = kLazy2 ? (myst == 0 ? lenF : myst == 1 ? lenHB : lenHC) : (myst ? lenHB : lenF);
                        const uint32_t moff = kLazy2 ? (myst == 0 ? offF : myst == 1 ? offHB : offHC) : (myst ? offHB : offF);
                        const uint32_t lit = isM ? 0u : P.B(q);
                        const uint32_t lsym = isM ? kFirstLenSym + len_slot_only(mlen) : lit;
                        const uint32_t ocls = isM ? 8 + (mlen >= 9) : (((lit >> 5) & 6) | (lit & 1));
                        atomicAdd(&S.fl[lsym], 1u);
                        atomicAdd(&S.new_obs[ocls], 1u);
                        if (isM) atomicAdd(&S.fo[off_slot_only(moff)], 1u);
                        tok[ti] = isM ? (0x80000000u | (mlen << 16) | moff) : lit;
                    }
                    {
                        uint32_t added = __popc(commit_mask);
                        ntok += added; num_new_obs += added;
                        nmatch += __popc(commit_mask & mmask);
                    }
                    in_h = next_h;
                    p = next_p;
                    __syncwarp();
                    if (event == 1) {
                        // recalculate_min_match_len() from the literal frequencies so far
                        uint32_t total = 0;
                        for (int i = 0; i < 8; i++) total += S.fl[lane * 8 + i];
                        for (int o = 16; o; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
                        uint32_t cutoff = total >> 10, nu = 0;
                        for (int i = 0; i < 8; i++) nu += (S.fl[lane * 8 + i] > cutoff);
                        for (int o = 16; o; o >>= 1) nu += __shfl_xor_sync(0xFFFFFFFFu, nu, o);
                        min_len = choose_min_match_len(nu, depth);
                        next_recalc += min(n - next_recalc, p - bb);
                    } else if (event == 2) {
                        // should_end_block() -> do_end_block_check()
                        uint32_t block_length = p - bb;
                        if (num_obs > 0) {
                            uint32_t d = 0;
                            if (lane < 10) {
                                uint32_t expected = S.obs[lane] * num_new_obs, actual = S.new_obs[lane] * num_obs;
                                d = actual > expected ? actual - expected : expected - actual;
                            }
                            for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o);
                            uint32_t num_items = num_obs + num_new_obs;
                            uint32_t cutoff = num_new_obs * 200 / 512 * num_obs;
                            if (block_length < 10000 && num_items < 8192)
                                cutoff += (uint32_t)((uint64_t)cutoff * (8192 - num_items) / 8192);
                            if (d + (block_length / 4096) * num_obs >= cutoff) end_block = true;
                        }
                        if (!end_block) {
                            if (lane < 10) { S.obs[lane] += S.new_obs[lane]; S.new_obs[lane] = 0; }
                            num_obs += num_new_obs; num_new_obs = 0;
                        }
                        __syncwarp();
                    } else if (event == 3) {
                        end_block = true;
                    }
                } while (p < max_block_end && !end_block);
                // nothing stays in flight across the block end: the CTA must not exit under a pending copy (tiles a final
                // long match jumped over), and the ring doubles as the block-finishing scratch
                P.drain();
                if (lane == 0) {
                    S.blk_end = p; S.ntok = ntok; S.is_final = (final_block && p == n) ? 1u : 0u;
                    S.fl[kEndOfBlock] += 1;
                }
            }
            PHASE(1);
            __syncthreads();
            PHASE(2);
            // ---------------- finish block (all threads) ----------------
            const uint32_t be = S.blk_end, ntok = S.ntok, block_len = be - bb, is_final = S.is_final;
            uint8_t *ll = S.lens, *ol = S.lens + kNumLitlen;
            make_huffman_code<uint16_t>(S, S.fl, kNumLitlen, kMaxLitlenCw, ll, S.lcw);
            PHASE(3);
            make_huffman_code<uint16_t>(S, S.fo, kNumOffset, kMaxOffsetCw, ol, S.ocw);
            PHASE(4);
            {   // symbol costs (parallel) — before the precode items are written: those reuse the frequency array
                uint32_t dyn = 0, stat = 0;
                for (uint32_t s = tid; s < kNumLitlen; s += kEmitThreads) {
                    uint32_t f = S.fl[s];
                    if (s < 256) { dyn += f * ll[s]; stat += f * (s < 144 ? 8 : 9); }
                    else if (s == 256) { dyn += ll[s]; stat += 7; }   // one end-of-block symbol
                    else if (s < 286) {
                        uint32_t k = s - 257;
                        uint32_t extra = (k < 8 || k == 28) ? 0 : (k - 4) / 4;
                        dyn += f * (extra + ll[s]); stat += f * (extra + c_static_litlen_len[s]);
                    }
                }
                if (tid < 30) {
                    uint32_t extra = tid < 4 ? 0 : (tid - 2) / 2;
                    dyn += S.fo[tid] * (extra + ol[tid]); stat += S.fo[tid] * (extra + 5);
                }
                static_assert(kEmitThreads == 32, "one warp sums the symbol costs");
                for (int o = 16; o; o >>= 1) { dyn += __shfl_xor_sync(0xFFFFFFFFu, dyn, o); stat += __shfl_xor_sync(0xFFFFFFFFu, stat, o); }
                if (tid == 0) { S.cost_dyn = dyn; S.cost_static = stat; }
            }
            __syncthreads();
            if (tid == 0) {
                // deflate_precompute_huffman_header(): RLE of the code lengths
                uint32_t nlit = kNumLitlen, noff = kNumOffset;
                while (nlit > 257 && ll[nlit - 1] == 0) nlit--;
                while (noff > 1 && ol[noff - 1] == 0) noff--;
                S.nlit = nlit; S.noff = noff;
                for (int i = 0; i < kNumPrecode; i++) S.pfreq[i] = 0;
                const uint32_t num_lens = nlit + noff;
                uint32_t ni = 0, run_start = 0;
#define LENS_AT(i_) ((i_) < nlit ? ll[(i_)] : ol[(i_) - nlit])
                do {
                    uint32_t len = LENS_AT(run_start);
                    uint32_t run_end = run_start;
                    do { run_end++; } while (run_end != num_lens && len == LENS_AT(run_end));
                    if (len == 0) {
                        while (run_end - run_start >= 11) {
                            uint32_t eb = min(run_end - run_start - 11, 0x7Fu);
                            S.pfreq[18]++; S.itemsp()[ni++] = (uint16_t)(18 | (eb << 5)); run_start += 11 + eb;
                        }
                        if (run_end - run_start >= 3) {
                            uint32_t eb = min(run_end - run_start - 3, 7u);
                            S.pfreq[17]++; S.itemsp()[ni++] = (uint16_t)(17 | (eb << 5)); run_start += 3 + eb;
                        }
                    } else if (run_end - run_start >= 4) {
                        S.pfreq[len]++; S.itemsp()[ni++] = (uint16_t)len; run_start++;
                        do {
                            uint32_t eb = min(run_end - run_start - 3, 3u);
                            S.pfreq[16]++; S.itemsp()[ni++] = (uint16_t)(16 | (eb << 5)); run_start += 3 + eb;
                        } while (run_end - run_start >= 3);
                    }
                    while (run_start != run_end) { S.pfreq[len]++; S.itemsp()[ni++] = (uint16_t)len; run_start++; }
                } while (run_start != num_lens);
#undef LENS_AT
                S.nitems = ni;
            }
            __syncthreads();
            PHASE(5);
            make_huffman_code<uint16_t>(S, S.pfreq, kNumPrecode, kMaxPreCw, S.plen, S.pcw);
            PHASE(6);
            if (tid == 0) {
                const uint8_t perm[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
                uint32_t nexpl = kNumPrecode;
                while (nexpl > 4 && S.plen[perm[nexpl - 1]] == 0) nexpl--;
                S.nexpl = nexpl;
                uint32_t dyn = 3 + 5 + 5 + 4 + 3 * nexpl, stat = 3;
                for (int s = 0; s < kNumPrecode; s++) {
                    uint32_t extra = s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0;
                    dyn += S.pfreq[s] * (extra + S.plen[s]);
                }
                S.cost_dyn += dyn; S.cost_static += stat;
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t bitcount = S.G & 7;
                uint32_t unc = 3 + ((0u - (bitcount + 3)) & 7) + 32 + 40 * ((block_len + 65534) / 65535 - 1) + 8 * block_len;
                uint32_t best = min(S.cost_dyn, min(S.cost_static, unc));
                S.btype = (best == unc) ? 0 : (best == S.cost_static) ? 1 : 2;
            }
            __syncthreads();
            const uint32_t btype = S.btype;
            PHASE(7);
            if (btype == 0) {
                uint32_t pos = bb;
                do {
                    uint32_t len = min(be - pos, 65535u);
                    uint32_t bfinal = (be - pos <= 65535u) ? is_final : 0u;
                    // 3 header bits, pad to a byte boundary
                    stage_begin(S, 4);
                    uint32_t g0 = S.G & 31;
                    uint32_t hb = 3 + ((0u - ((S.G & 7) + 3)) & 7);
                    if (tid == 0) stage_put(S.stagep(), g0, bfinal, 3);
                    stage_commit(S, payload, hb);
                    uint32_t tot_bits = 32 + 8 * len;
                    for (uint32_t cb = 0; cb < tot_bits;) {
                        uint32_t chunk = min(tot_bits - cb, (uint32_t)(kStageWords - 4) * 32);
                        stage_begin(S, kStageWords);
                        g0 = S.G & 31;
                        for (uint32_t bit = cb + tid * 8; bit < cb + chunk; bit += kEmitThreads * 8) {
                            uint32_t v;
                            if (bit < 16) v = (len >> bit) & 0xFF;
                            else if (bit < 32) v = ((~len & 0xFFFF) >> (bit - 16)) & 0xFF;
                            else v = in[pos + (bit - 32) / 8];
                            stage_put(S.stagep(), g0 + (bit - cb), v, 8);
                        }
                        stage_commit(S, payload, chunk);
                        cb += chunk;
                    }
                    pos += len;
                } while (pos != be);
            } else {
                // ---- block header (thread 0) ----
                stage_begin(S, kStageWords);
                uint32_t hbits = 0;
                if (tid == 0) {
                    uint32_t g0 = S.G & 31, bp = g0;
                    stage_put(S.stagep(), bp, is_final | (btype << 1), 3); bp += 3;
                    if (btype == 2) {
                        const uint8_t perm[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
                        stage_put(S.stagep(), bp, S.nlit - 257, 5); bp += 5;
                        stage_put(S.stagep(), bp, S.noff - 1, 5); bp += 5;
                        stage_put(S.stagep(), bp, S.nexpl - 4, 4); bp += 4;
                        for (uint32_t i = 0; i < S.nexpl; i++) { stage_put(S.stagep(), bp, S.plen[perm[i]], 3); bp += 3; }
                        for (uint32_t i = 0; i < S.nitems; i++) {
                            uint32_t it = S.itemsp()[i], sym = it & 0x1F;
                            uint32_t pl = S.plen[sym];
                            stage_put(S.stagep(), bp, S.pcw[sym], pl); bp += pl;
                            uint32_t extra = sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0;
                            stage_put(S.stagep(), bp, it >> 5, extra); bp += extra;
                        }
                    }
                    S.scan[0] = bp - g0;
                }
                __syncthreads();
                hbits = S.scan[0];
                stage_commit(S, payload, hbits);
                PHASE(8);
                // ---- tokens (parallel) ----
                const bool dynamic = (btype == 2);
                for (uint32_t t0 = 0; t0 < ntok + 1; t0 += kChunkTok) {   // +1: the end-of-block symbol rides as a token
                    stage_begin(S, kStageWords);
                    uint64_t code[kTokPerThread]; uint32_t cl[kTokPerThread];
                    uint32_t sum = 0;
#pragma unroll
                    for (int k = 0; k < kTokPerThread; k++) {
                        uint32_t ti = t0 + tid * kTokPerThread + k;
                        uint64_t c = 0; uint32_t l = 0;
                        if (ti < ntok) {
                            uint32_t t = tok[ti];
                            if (!(t & 0x80000000u)) {
                                c = dynamic ? S.lcw[t] : c_static_litlen_cw[t];
                                l = dynamic ? ll[t] : c_static_litlen_len[t];
                            } else {
                                uint32_t len = (t >> 16) & 0x1FF, off = t & 0xFFFF;
                                uint32_t ls, le, lv, os, oe, ov;
                                len_slot(len, ls, le, lv); off_slot(off, os, oe, ov);
                                uint32_t lsym = kFirstLenSym + ls;
                                uint32_t lc = dynamic ? S.lcw[lsym] : c_static_litlen_cw[lsym];
                                uint32_t lcl = dynamic ? ll[lsym] : c_static_litlen_len[lsym];
                                uint32_t oc = dynamic ? S.ocw[os] : (__brev(os) >> 27);
                                uint32_t ocl = dynamic ? ol[os] : 5;
                                c = lc; l = lcl;
                                c |= (uint64_t)lv << l; l += le;
                                c |= (uint64_t)oc << l; l += ocl;
                                c |= (uint64_t)ov << l; l += oe;
                            }
                        } else if (ti == ntok) {
                            c = dynamic ? S.lcw[kEndOfBlock] : c_static_litlen_cw[kEndOfBlock];
                            l = dynamic ? ll[kEndOfBlock] : c_static_litlen_len[kEndOfBlock];
                        }
                        code[k] = c; cl[k] = l; sum += l;
                    }
                    // block-wide exclusive scan of `sum`
                    uint32_t incl = sum;
                    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if ((tid & 31) >= (uint32_t)o) incl += v; }
                    if ((tid & 31) == 31) S.scan[tid >> 5] = incl;
                    __syncthreads();
                    uint32_t woff = 0, total = 0;
                    for (int w = 0; w < kEmitThreads / 32; w++) { uint32_t v = S.scan[w]; if (w < (int)(tid >> 5)) woff += v; total += v; }
                    uint32_t bp = (S.G & 31) + woff + incl - sum;
#pragma unroll
                    for (int k = 0; k < kTokPerThread; k++) { stage_put(S.stagep(), bp, code[k], cl[k]); bp += cl[k]; }
                    stage_commit(S, payload, total);
                }
            }
            PHASE(9);
            p = be;
            P.restart(be);
            if (be >= n) break;
        }
    }

    PHASE(10);
    // ---- stream tail: sync-flush marker, final partial byte, container ----
    if (sync_flush) {
        stage_begin(S, 8);
        uint32_t hb = 3 + ((0u - ((S.G & 7) + 3)) & 7);
        stage_commit(S, payload, hb);            // 3 zero bits + padding
        stage_begin(S, 8);
        if (tid == 0) stage_put(S.stagep(), S.G & 31, 0xFFFF0000ull, 32);
        stage_commit(S, payload, 32);
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t G = S.G;
        payload[G >> 5] = S.carry;               // upper bits are zero
        uint32_t nbytes = (G + 7) >> 3;
        uint32_t total = nbytes;
        uint32_t hdr_off = kOutPayloadOff;
        int32_t st = S.status;
        if (format == 4 || format == 3) {        // Bgzf / Mgzip member per unit
            const uint32_t hs = format == 4 ? 18 : 20;
            uint8_t *h = slot + kOutPayloadOff - hs;
            hdr_off = kOutPayloadOff - hs;
            uint32_t xfl = level >= 9 ? 2 : level <= 1 ? 4 : 0;
            h[0] = 31; h[1] = 139; h[2] = 8; h[3] = 4; h[4] = h[5] = h[6] = h[7] = 0; h[8] = (uint8_t)xfl; h[9] = 255;
            if (format == 4) {
                h[10] = 6; h[11] = 0; h[12] = 'B'; h[13] = 'C'; h[14] = 2; h[15] = 0;
                uint32_t bs = (uint16_t)((uint16_t)nbytes + 26 - 1);
                h[16] = (uint8_t)bs; h[17] = (uint8_t)(bs >> 8);
                // GZPB_EBLOCKSIZE (bgzf.rs:218-223 tests the payload against 65536; the member — payload + 26 — must fit, or the
                // u16 BSIZE above wraps and a corrupt block would be reported as success: deliberate deviation)
                if (nbytes + 26 > 65536) st = -3;
            } else {
                h[10] = 8; h[11] = 0; h[12] = 'I'; h[13] = 'G'; h[14] = 4; h[15] = 0;
                uint32_t bs = nbytes + 28;
                h[16] = (uint8_t)bs; h[17] = (uint8_t)(bs >> 8); h[18] = (uint8_t)(bs >> 16); h[19] = (uint8_t)(bs >> 24);
            }
            uint8_t *f = slot + kOutPayloadOff + nbytes;
            uint32_t crc = unit_sum;
            f[0] = (uint8_t)crc; f[1] = (uint8_t)(crc >> 8); f[2] = (uint8_t)(crc >> 16); f[3] = (uint8_t)(crc >> 24);
            f[4] = (uint8_t)dl; f[5] = (uint8_t)(dl >> 8); f[6] = (uint8_t)(dl >> 16); f[7] = (uint8_t)(dl >> 24);
            total = hs + nbytes + 8;
            if (format == 4 && (flags & 1u)) {   // is_last: append BGZF_EOF (deflate.rs:622-624)
                const uint8_t eof[28] = {0x1f,0x8b,0x08,0x04,0,0,0,0,0,0xff,0x06,0,0x42,0x43,0x02,0,0x1b,0,0x03,0,0,0,0,0,0,0,0,0};
                for (int i = 0; i < 28; i++) f[8 + i] = eof[i];
                total += 28;
            }
            uint32_t avail = dl + max(128u, (uint32_t)((double)dl * 0.1)) + 8;
            if (nbytes > avail) st = -4;          // GZPB_ECOMPRESS: libdeflate would have returned 0
        } else {
            uint32_t avail = dl + max(128u, (uint32_t)((double)dl * 0.1));
            if (nbytes > avail) st = -4;
        }
        out_len[u * 2] = total;
        out_len[u * 2 + 1] = hdr_off;
        out_status[u] = st;
    }
}

// =============================================================================
// k_scan: exclusive scan of the per-unit output sizes (single CTA).
// k_gather: compaction of the fixed-stride unit slots into stream order.
// =============================================================================
__global__ void __launch_bounds__(1024)
k_scan(const uint32_t *__restrict__ out_len, uint64_t *__restrict__ offsets, uint32_t nunits,
       const uint64_t *__restrict__ base_ptr, uint64_t cap, int32_t *__restrict__ overflow, uint64_t *end_mirror)
{
    __shared__ uint64_t part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (nunits + 1023) / 1024;
    const uint64_t base = base_ptr ? *base_ptr : 0;
    uint64_t s = 0;
    for (uint32_t i = tid * per; i < min(nunits, (tid + 1) * per); i++) s += out_len[2 * i];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        uint64_t acc = base;
        for (int i = 0; i < 1024; i++) { uint64_t v = part[i]; part[i] = acc; acc += v; }
        offsets[nunits] = acc;
        if (end_mirror) { *(volatile uint64_t *)end_mirror = acc; __threadfence_system(); }
        *overflow = (acc > cap) ? 1 : 0;
    }
    __syncthreads();
    uint64_t acc = part[tid];
    for (uint32_t i = tid * per; i < min(nunits, (tid + 1) * per); i++) { offsets[i] = acc; acc += out_len[2 * i]; }
}

__global__ void __launch_bounds__(256)
k_gather(const uint8_t *__restrict__ out_base, const uint32_t *__restrict__ out_len, const uint64_t *__restrict__ offsets,
         uint8_t *__restrict__ dst, const int32_t *__restrict__ overflow, uint32_t out_stride, uint32_t nunits)
{
    if (*overflow) return;
    // grid-stride over the units: a host-memory target (zero-copy D2H) is launched with a small grid — its stores drain
    // at PCIe speed, and a full grid of stalled CTAs would hold the thread slots the next batch's kernels need
    for (uint32_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        const uint32_t len = out_len[2 * u], hoff = out_len[2 * u + 1];
        const uint8_t *src = out_base + (size_t)u * out_stride + hoff;
        uint8_t *d = dst + offsets[u];
        // head bytes until d is 4-byte aligned
        uint32_t head = min(len, (uint32_t)((4 - ((uintptr_t)d & 3)) & 3));
        if (threadIdx.x < head) d[threadIdx.x] = src[threadIdx.x];
        const uint8_t *s2 = src + head; uint8_t *d2 = d + head;
        uint32_t rem = len - head, nw = rem >> 2;
        const uint32_t *sw = (const uint32_t *)((uintptr_t)s2 & ~(uintptr_t)3);
        uint32_t sh = ((uintptr_t)s2 & 3) * 8;
        uint32_t *dw = (uint32_t *)d2;
        for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) dw[i] = __funnelshift_r(sw[i], sw[i + 1], sh);
        uint32_t tail = rem & 3;
        if (threadIdx.x < tail) d2[nw * 4 + threadIdx.x] = s2[nw * 4 + threadIdx.x];
    }
}

// =============================================================================
// k_check: per-unit checksum of the DATA part of a unit (Check::update):
// kind 0 = CRC-32 (512-byte chunks recombined with x^(8*512*j) mod P),
// kind 1 = Adler-32 (per-chunk partial sums recombined mod 65521).
// HBM-bound: one pass over the input.
// =============================================================================
__global__ void __launch_bounds__(256)
k_check(const __grid_constant__ Geo g, uint32_t *__restrict__ sum_out, int kind)
{
    __shared__ uint32_t s_crc;
    __shared__ unsigned long long s_a, s_b;
    __shared__ uint32_t s_tab[4][256];
    const uint32_t u = blockIdx.x, tid = threadIdx.x;
    const uint32_t dict = g.unit_dict[u];
    const uint32_t n = g.unit_len[u] - dict;
    const uint8_t *in = g.in + (size_t)u * g.in_stride + dict;
    if (tid == 0) { s_crc = 0; s_a = 0; s_b = 0; }
    if (kind == 0) {
        const uint32_t mis = (uint32_t)((uintptr_t)in & 3);
        for (uint32_t i = tid; i < 1024; i += 256) s_tab[i >> 8][i & 255] = c_crc_tab[i >> 8][i & 255];
        __syncthreads();
        uint32_t acc = 0;
        for (uint32_t j = tid; j * 512 < n; j += 256) {
            uint32_t end = n - 512 * j, beg = end >= 512 ? end - 512 : 0;
            uint32_t c = ~0u, pos = beg;
            while (pos < end && ((pos + mis) & 3)) { c = (c >> 8) ^ s_tab[0][(c ^ in[pos]) & 0xFF]; pos++; }
            for (; pos + 4 <= end; pos += 4) {
                c ^= __ldg((const uint32_t *)(in + pos));
                c = s_tab[3][c & 0xFF] ^ s_tab[2][(c >> 8) & 0xFF] ^ s_tab[1][(c >> 16) & 0xFF] ^ s_tab[0][c >> 24];
            }
            while (pos < end) { c = (c >> 8) ^ s_tab[0][(c ^ in[pos]) & 0xFF]; pos++; }
            c = ~c;
            if (j >= 1024) c = gf2_mulmod(c, c_xpow512k[(j >> 10) & 15], kCrcPoly);      // beyond the 512 KiB the first table spans
            acc ^= (j == 0) ? c : gf2_mulmod(c, c_xpow512[j & 1023], kCrcPoly);
        }
        for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if ((tid & 31) == 0 && acc) atomicXor(&s_crc, acc);
        __syncthreads();
        if (tid == 0) sum_out[u] = s_crc;
    } else {
        __syncthreads();
        unsigned long long sa = 0, sb2 = 0;
        for (uint32_t j = tid; j * 512 < n; j += 256) {
            uint32_t beg = 512 * j, end = min(n, beg + 512);
            uint32_t a = 0, b = 0;
            for (uint32_t pos = beg; pos < end; pos++) { a += in[pos]; b += a; }
            sa += a;
            sb2 += (b + (unsigned long long)a * (n - end)) % 65521ull;
        }
        for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(0xFFFFFFFFu, sa, o); sb2 += __shfl_xor_sync(0xFFFFFFFFu, sb2, o); }
        if ((tid & 31) == 0) { atomicAdd(&s_a, sa); atomicAdd(&s_b, sb2); }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = (uint32_t)((1ull + s_a) % 65521ull), b = (uint32_t)(((unsigned long long)n + s_b) % 65521ull);
            sum_out[u] = (b << 16) | a;
        }
    }
}

// =============================================================================
// k_check_combine: Check::combine over a batch (src/check.rs:162, 121-128;
// src/par/compress.rs:308).  (sum_a, len_a) o (sum_b, len_b) is associative:
//   CRC-32 : (x^(8*len_b) * sum_a mod P) xor sum_b        (GF(2) polynomial arithmetic)
//   Adler32: zlib's adler32_combine recurrence
// so the batch folds with a block-wide tree reduction; the host folds ONE value per batch
// into the stream's running check.
// =============================================================================
__device__ __forceinline__ uint32_t adler_combine_dev(uint32_t a1, uint32_t a2, uint64_t len2)
{
    const uint32_t BASE = 65521u;
    uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t sum1 = a1 & 0xFFFF;
    uint32_t sum2 = (uint32_t)(((uint64_t)rem * sum1) % BASE);
    sum1 += (a2 & 0xFFFF) + BASE - 1;
    sum2 += (a1 >> 16) + (a2 >> 16) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return sum1 | (sum2 << 16);
}

__global__ void __launch_bounds__(256)
k_check_combine(const uint32_t *__restrict__ sums, const uint32_t *__restrict__ unit_len, const uint32_t *__restrict__ unit_dict,
                uint32_t nunits, int kind, uint32_t *__restrict__ out /* [0] = sum, [1],[2] = length lo/hi */)
{
    __shared__ uint32_t s_sum[256];
    __shared__ unsigned long long s_len[256];
    const uint32_t tid = threadIdx.x;
    const uint32_t per = (nunits + 255) / 256;
    uint32_t acc = (kind == 1) ? 1u : 0u;       // identity: crc32("") = 0, adler32("") = 1
    unsigned long long alen = 0;
    for (uint32_t i = tid * per; i < min(nunits, (tid + 1) * per); i++) {
        const unsigned long long l = unit_len[i] - unit_dict[i];
        if (l) acc = (kind == 1) ? adler_combine_dev(acc, sums[i], l) : (gf2_mulmod(acc, gf2_xpow8(l, kCrcPoly), kCrcPoly) ^ sums[i]);
        alen += l;
    }
    s_sum[tid] = acc; s_len[tid] = alen;
    __syncthreads();
    for (uint32_t stride = 1; stride < 256; stride <<= 1) {
        if ((tid & (2 * stride - 1)) == 0) {
            const uint32_t b = s_sum[tid + stride];
            const unsigned long long lb = s_len[tid + stride];
            if (lb) s_sum[tid] = (kind == 1) ? adler_combine_dev(s_sum[tid], b, lb) : (gf2_mulmod(s_sum[tid], gf2_xpow8(lb, kCrcPoly), kCrcPoly) ^ b);
            s_len[tid] += lb;
        }
        __syncthreads();
    }
    if (tid == 0) { out[0] = s_sum[0]; out[1] = (uint32_t)s_len[0]; out[2] = (uint32_t)(s_len[0] >> 32); }
}

cudaError_t launch_check_combine(const uint32_t *sums, const uint32_t *unit_len, const uint32_t *unit_dict, uint32_t nunits, int kind,
                                 uint32_t *out3, cudaStream_t st)
{
    GZPB_LAUNCH(k_check_combine, 1, 256, 0, st, sums, unit_len, unit_dict, nunits, kind, out3);
    return cudaGetLastError();
}

void read_phase_counters(unsigned long long *out, bool reset)
{
    cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_phase, z, sizeof z); }
}

// ---------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------
static bool debug_sync() { static int v = -1; if (v < 0) { const char *e = getenv("GZPB_DEBUG_SYNC"); v = (e && *e == '1') ? 1 : 0; } return v == 1; }
#define DBG_SYNC(name)                                                                        \
    do {                                                                                      \
        if (debug_sync()) {                                                                   \
            cudaError_t e_ = cudaStreamSynchronize(st);                                       \
            fprintf(stderr, "[gzpb] %s: %s\n", name, cudaGetErrorString(e_));                 \
            if (e_ != cudaSuccess) return e_;                                                 \
        }                                                                                     \
    } while (0)

static Geo make_geo(const DeflateBatch &b)
{
    Geo g;
    g.in = b.in; g.unit_len = b.unit_len; g.unit_dict = b.unit_dict;
    g.in_stride = b.in_stride; g.m_stride = b.m_stride; g.tok_stride = b.tok_stride; g.out_stride = b.out_stride;
    g.spu = b.spu; g.seg = b.seg;
    return g;
}

#define GZPB_COUNTED_LAUNCH(...) do { GZPB_LAUNCH(__VA_ARGS__); if (b.launch_counter) ++*b.launch_counter; } while (0)
cudaError_t launch_deflate_pipeline(const DeflateBatch &b, cudaStream_t st)
{
    static bool attr_done[64] = {};          // function attributes are per device (several GPUs in one process)
    const int match_smem = kInStride + 65536 * 2;
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (cur_dev < 0 || cur_dev >= 64) return cudaErrorInvalidValue;
    if (!attr_done[cur_dev]) {
        cudaFuncSetAttribute(k_match, cudaFuncAttributeMaxDynamicSharedMemorySize, match_smem);
        attr_done[cur_dev] = true;
    }
    if (b.nunits == 0) return cudaSuccess;
    LevelParams lp;
    if (!level_params(b.level, &lp)) return cudaErrorInvalidValue;
    const Geo g = make_geo(b);
    // Check::update: folded into k_split (one partial sum per sub-unit, k_emit folds them); level 0 has no k_split pass
    const bool fold_check = b.check_kind >= 0 && lp.mode >= 0 && b.sum_part != nullptr;
    if (b.check_kind >= 0 && !fold_check) {
        if (b.timer) b.timer->start(KT_CRC, st);
        GZPB_COUNTED_LAUNCH(k_check, b.nunits, 256, 0, st, g, b.crc, b.check_kind);
        if (b.timer) b.timer->stop(st);
    }
    if (lp.mode >= 0) {
        if (b.timer) b.timer->start(KT_CHAIN, st);
        GZPB_COUNTED_LAUNCH(k_split, b.nunits * b.spu, kSplitThreads, 0, st, g, b.lists, b.list_start, b.next4, b.prev3, lp.ht,
                    fold_check ? b.sum_part : (uint32_t *)nullptr, b.check_kind);
        DBG_SYNC("k_split");
        if (b.spu > 1) {
            GZPB_COUNTED_LAUNCH((k_link<true, false>), b.nunits * kL3, 32, 0, st, g, b.lists, b.list_start, b.next4, b.prev3, b.clen);
            GZPB_COUNTED_LAUNCH((k_link<true, true>), b.nunits * kL4, 32, 0, st, g, b.lists, b.list_start, b.next4, b.prev3, b.clen);
        } else {
            GZPB_COUNTED_LAUNCH((k_link<false, false>), b.nunits * kL3, 32, 0, st, g, b.lists, b.list_start, b.next4, b.prev3, b.clen);
            GZPB_COUNTED_LAUNCH((k_link<false, true>), b.nunits * kL4, 32, 0, st, g, b.lists, b.list_start, b.next4, b.prev3, b.clen);
        }
        DBG_SYNC("k_link");
        if (b.timer) { b.timer->stop(st); b.timer->start(KT_MATCH, st); }
        GZPB_COUNTED_LAUNCH(k_match, b.nunits * b.spu, kMatchThreads, match_smem, st, g, b.next4, b.prev3, b.mtab, b.mtab2, b.clen, b.order, lp.depth, lp.nice, lp.mode, lp.ht);
        DBG_SYNC("k_match");
        if (b.timer) b.timer->stop(st);
    }
    if (b.timer) b.timer->start(KT_EMIT, st);
    if (lp.mode == 2)
        GZPB_COUNTED_LAUNCH(k_emit<true>, b.nunits, kEmitThreads, 0, st, g, b.unit_flags, b.mtab, b.mtab2, b.crc, fold_check ? b.sum_part : (const uint32_t *)nullptr, b.check_kind,
                    b.tokens, b.out, b.out_len, b.status, lp.mode, lp.depth, lp.nice, b.level, b.format);
    else
        GZPB_COUNTED_LAUNCH(k_emit<false>, b.nunits, kEmitThreads, 0, st, g, b.unit_flags, b.mtab, b.mtab2, b.crc, fold_check ? b.sum_part : (const uint32_t *)nullptr, b.check_kind,
                    b.tokens, b.out, b.out_len, b.status, lp.mode, lp.depth, lp.nice, b.level, b.format);
    DBG_SYNC("k_emit");
    if (b.timer) b.timer->stop(st);
    return cudaGetLastError();
}

// scan + gather are launched separately so that the caller can order them
// after the previous batch's scan (stream offsets chain from batch to batch).
#undef GZPB_COUNTED_LAUNCH
cudaError_t launch_pack(const DeflateBatch &b, cudaStream_t st)
{
    if (b.nunits == 0) return cudaSuccess;
    if (b.timer) b.timer->start(KT_GATHER, st);
    GZPB_LAUNCH(k_scan, 1, 1024, 0, st, b.out_len, b.offsets, b.nunits, b.base_ptr, b.packed_cap, b.overflow, b.end_mirror);
    const uint32_t ggrid = (b.packed_on_host && b.nunits > (uint32_t)b.packed_on_host) ? (uint32_t)b.packed_on_host : b.nunits;
    GZPB_LAUNCH(k_gather, ggrid, 256, 0, st, b.out, b.out_len, b.offsets, b.packed, b.overflow, b.out_stride, b.nunits);
    DBG_SYNC("k_scan+k_gather");
    if (b.timer) b.timer->stop(st);
    return cudaGetLastError();
}

}  // namespace gzpb
