// snappy_kernels.cu — sm_100a kernel for gzp's Snap format: every gzp block is a
// complete Snappy *framed* stream (/root/reference/src/snap.rs:61-74 ->
// snap::read::FrameEncoder): stream identifier, then per <= 65 536 source bytes one
// chunk {type, u24 length, masked CRC-32C, raw Snappy block or stored bytes}.
//
// One CTA (4 warps) per 64 KiB chunk.  The chunk is staged into shared memory with one
// TMA bulk copy; warp 0 runs the raw encoder's greedy parse, warps 1-3 compute the
// CRC-32C of the chunk meanwhile.  The parse is sequential only where the format makes
// it so (the u16 hash table is updated at visited positions only): the probe loop — the
// scan for the next 4-byte match, with its accelerating skip — is evaluated 32 probes
// at a time (hashes and candidate compares in parallel, MATCH.ANY orders probes that
// share a bucket, the first hit wins and later probes leave no trace), literals and
// match extension are warp-wide.  Bit-identical to the oracle (oracle/snappy_oracle.c).
#include <stdio.h>

#include "gzpb_common.cuh"
#include "deflate_kernels.cuh"
#include "snappy_kernels.cuh"

namespace gzpb {

__constant__ uint32_t c_crc32c_tab[4][256];
__constant__ uint32_t c_xpow512_c[130];   // x^(8*512*j) mod P (Castagnoli)

void upload_snappy_constants()
{
    static uint32_t tab[4][256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (kCrc32cPoly & (0u - (c & 1)));
        tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int s = 1; s < 4; s++) tab[s][i] = (tab[s - 1][i] >> 8) ^ tab[0][tab[s - 1][i] & 0xFF];
    cudaMemcpyToSymbol(c_crc32c_tab, tab, sizeof tab);
    uint32_t xp[130];
    for (int j = 0; j < 130; j++) xp[j] = gf2_xpow8((uint64_t)512 * j, kCrc32cPoly);
    cudaMemcpyToSymbol(c_xpow512_c, xp, sizeof xp);
}

constexpr uint32_t kSnapChunk = 65536;
constexpr uint32_t kInputMargin = 15;
constexpr uint32_t kMinNonLiteral = 17;
constexpr int kSnapThreads = 128;
constexpr int kSnapSmem = (int)kSnapChunk + 64 + 16384 * 2 + 1024 * 4;     // staged chunk + u16 hash table + CRC tables

__global__ void __launch_bounds__(kSnapThreads)
k_snap(const uint8_t *__restrict__ in_base, const uint32_t *__restrict__ unit_len, uint32_t in_stride, uint32_t cpu,
       uint8_t *__restrict__ out_base, uint32_t out_stride, uint32_t *__restrict__ out_len)
{
    GZPB_DYN_SMEM(smem);
    uint32_t *in_w = (uint32_t *)smem;                                  // the chunk (+ 64 bytes of slack for word loads)
    const uint8_t *in_s = (const uint8_t *)smem;
    uint16_t *table = (uint16_t *)(smem + kSnapChunk + 64);
    uint32_t(*s_tab)[256] = (uint32_t(*)[256])(smem + kSnapChunk + 64 + 16384 * 2);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_crc, s_d;
    const uint32_t chunk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t u = chunk / cpu, k = chunk % cpu;
    const uint32_t n_unit = unit_len[u];
    const uint32_t off = k * kSnapChunk;
    if (off >= n_unit) {
        if (tid == 0) { out_len[2 * chunk] = 0; out_len[2 * chunk + 1] = 32; }
        return;
    }
    const uint32_t n = min(kSnapChunk, n_unit - off);
    const uint8_t *src = in_base + (size_t)u * in_stride + off;
    uint8_t *slot = out_base + (size_t)chunk * out_stride;
    uint8_t *dst = slot + 32;

    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); s_crc = 0; }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (n + 15u) & ~15u;                       // the unit slot is padded: a 16-byte tail read stays inside it
        mbar_expect_tx(&bar, bytes);
        tma_load_1d(smem, src, bytes, &bar);
    }
    for (uint32_t i = tid; i < 1024; i += kSnapThreads) s_tab[i >> 8][i & 255] = c_crc32c_tab[i >> 8][i & 255];
    // table size: next power of two >= n, clamped to [256, 16384]
    uint32_t shift = 32 - 8, tsz = 256;
    while (tsz < 16384 && tsz < n) { shift--; tsz *= 2; }
    for (uint32_t i = tid; i < tsz / 2; i += kSnapThreads) ((uint32_t *)table)[i] = 0;
    mbar_wait(&bar, 0);
    __syncthreads();

    uint32_t d = 0;
    if (warp != 0) {
        // ---- warps 1-3: masked CRC-32C of the uncompressed chunk: 512-byte slices counted from the end, GF(2) recombination ----
        uint32_t acc = 0;
        for (uint32_t j = tid - 32; j * 512 < n; j += kSnapThreads - 32) {
            const uint32_t end = n - 512 * j, beg = end >= 512 ? end - 512 : 0;
            uint32_t c = ~0u, pos = beg;
            for (; pos + 4 <= end; pos += 4) {
                c ^= ld32u(in_w, pos);
                c = s_tab[3][c & 0xFF] ^ s_tab[2][(c >> 8) & 0xFF] ^ s_tab[1][(c >> 16) & 0xFF] ^ s_tab[0][c >> 24];
            }
            for (; pos < end; pos++) c = (c >> 8) ^ s_tab[0][(c ^ in_s[pos]) & 0xFF];
            c = ~c;
            acc ^= (j == 0) ? c : gf2_mulmod(c, c_xpow512_c[j], kCrc32cPoly);
        }
        for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
        if (lane == 0 && acc) atomicXor(&s_crc, acc);
    } else {
        // ---- warp 0: raw Snappy block ----
        if (lane == 0) {   // uvarint(n)
            uint32_t v = n;
            while (v >= 0x80) { dst[d++] = (uint8_t)(v | 0x80); v >>= 7; }
            dst[d++] = (uint8_t)v;
        }
        d = __shfl_sync(0xFFFFFFFFu, d, 0);
        uint32_t next_emit = 0;
#define HASH(x_) (((x_) * 0x1E35A7BDu) >> shift)
        // emit_literal(lit_end): tag by lane 0, bytes by the warp
        auto emit_literal = [&](uint32_t lit_end) {
            const uint32_t len = lit_end - next_emit, nm1 = len - 1;
            uint32_t tagb;
            if (nm1 <= 59) tagb = 1; else if (nm1 < 256) tagb = 2; else tagb = 3;
            if (lane == 0) {
                if (nm1 <= 59) dst[d] = (uint8_t)(nm1 << 2);
                else if (nm1 < 256) { dst[d] = 60 << 2; dst[d + 1] = (uint8_t)nm1; }
                else { dst[d] = 61 << 2; dst[d + 1] = (uint8_t)nm1; dst[d + 2] = (uint8_t)(nm1 >> 8); }
            }
            d += tagb;
            for (uint32_t i = lane; i < len; i += 32) dst[d + i] = in_s[next_emit + i];
            d += len;
        };

        if (n < kMinNonLiteral) {
            emit_literal(n);
        } else {
            const uint32_t s_limit = n - kInputMargin;
            const uint32_t lt = lanemask_lt();
            uint32_t s = 1;
            bool finished = false;
            while (!finished) {
                // ---- the probe loop of the raw encoder, 32 probes per step: probe i looks at P_i (P_0 = s, P_{i+1} = P_i +
                // (skip_i >> 5), skip_{i+1} = skip_i + (skip_i >> 5), skip_0 = 32) and runs only while P_{i+1} <= s_limit; it
                // takes the bucket's entry as candidate, stores P_i there and hits when the 4 bytes at both places agree.
                // In a batch: MATCH.ANY gives the probes of one bucket in order (a later probe's candidate is the nearest
                // earlier probe), the first hit ends the loop and the probes behind it must leave no trace. ----
                uint32_t candidate = 0, found = 0, skip = 32;
                for (;;) {
                    uint32_t P = s, sk = skip;
                    if (skip == 32) { P = s + lane; sk = 32 + lane; }                 // first batch: steps of one
                    else for (uint32_t j = 0; j < lane; j++) { const uint32_t st = sk >> 5; P += st; sk += st; }
                    const uint32_t Pn = P + (sk >> 5), skn = sk + (sk >> 5);            // the next probe's position / skip
                    const bool valid = Pn <= s_limit;
                    const uint32_t v = valid ? ld32u(in_w, P) : 0;
                    const uint32_t h = HASH(v);
                    const uint32_t grp = __match_any_sync(0xFFFFFFFFu, valid ? h : (0x10000u + lane));
                    const uint32_t lower = grp & lt;
                    const uint32_t pl = __shfl_sync(0xFFFFFFFFu, P, lower ? 31 - __clz(lower) : lane);
                    const uint32_t cand = valid ? (lower ? pl : (uint32_t)table[h]) : 0;
                    const bool hit = valid && ld32u(in_w, cand) == v;
                    const uint32_t vmask = __ballot_sync(0xFFFFFFFFu, valid), hmask = __ballot_sync(0xFFFFFFFFu, hit);
                    // probes that really run: up to the first hit, or the valid prefix
                    const uint32_t f = hmask ? (uint32_t)__ffs(hmask) - 1 : 32u;
                    const uint32_t run = hmask ? ((f == 31 ? 0xFFFFFFFFu : ((2u << f) - 1))) : vmask;
                    const uint32_t mine = grp & run;
                    __syncwarp();   // every probe has read its bucket
                    if (valid && ((run >> lane) & 1u) && (mine >> lane) == 1u) table[h] = (uint16_t)P;   // the bucket keeps its last probe
                    __syncwarp();
                    if (hmask) {
                        found = 1;
                        s = __shfl_sync(0xFFFFFFFFu, P, f);
                        candidate = __shfl_sync(0xFFFFFFFFu, cand, f);
                        break;
                    }
                    if (vmask != 0xFFFFFFFFu) break;                                    // s_next ran past s_limit: no further match
                    s = __shfl_sync(0xFFFFFFFFu, Pn, 31);
                    skip = __shfl_sync(0xFFFFFFFFu, skn, 31);
                }
                if (!found) break;
                emit_literal(s);
                for (;;) {
                    // ---- warp: extend the match 32 bytes per step ----
                    const uint32_t base = s;
                    s += 4;
                    uint32_t cand = candidate + 4;
                    for (;;) {
                        bool ok = (s + lane < n) && (in_s[s + lane] == in_s[cand + lane]);
                        uint32_t bad = ~__ballot_sync(0xFFFFFFFFu, ok);
                        if (bad) { uint32_t adv = __ffs(bad) - 1; s += adv; break; }
                        s += 32; cand += 32;
                    }
                    // ---- lane 0: emit_copy, table updates, immediate re-match test ----
                    uint32_t again = 0;
                    if (lane == 0) {
                        uint32_t offs = base - candidate, len = s - base;
                        while (len >= 68) { dst[d] = (uint8_t)((63 << 2) | 2); dst[d + 1] = (uint8_t)offs; dst[d + 2] = (uint8_t)(offs >> 8); d += 3; len -= 64; }
                        if (len > 64) { dst[d] = (uint8_t)((59 << 2) | 2); dst[d + 1] = (uint8_t)offs; dst[d + 2] = (uint8_t)(offs >> 8); d += 3; len -= 60; }
                        if (len <= 11 && offs <= 2047) { dst[d] = (uint8_t)(((offs >> 8) << 5) | ((len - 4) << 2) | 1); dst[d + 1] = (uint8_t)offs; d += 2; }
                        else { dst[d] = (uint8_t)(((len - 1) << 2) | 2); dst[d + 1] = (uint8_t)offs; dst[d + 2] = (uint8_t)(offs >> 8); d += 3; }
                        if (s >= s_limit) again = 2;   // done
                        else {
                            uint32_t x0 = ld32u(in_w, s - 1), x1 = ld32u(in_w, s + 3);
                            uint64_t x = (uint64_t)x0 | ((uint64_t)x1 << 32);
                            table[HASH((uint32_t)x)] = (uint16_t)(s - 1);
                            uint32_t cur = (uint32_t)(x >> 8), ch = HASH(cur);
                            candidate = table[ch];
                            table[ch] = (uint16_t)s;
                            if (cur != ld32u(in_w, candidate)) { s++; again = 0; }
                            else again = 1;
                        }
                    }
                    again = __shfl_sync(0xFFFFFFFFu, again, 0);
                    d = __shfl_sync(0xFFFFFFFFu, d, 0);
                    s = __shfl_sync(0xFFFFFFFFu, s, 0);
                    candidate = __shfl_sync(0xFFFFFFFFu, candidate, 0);
                    next_emit = (again == 0) ? s - 1 : s;
                    if (again == 2) { finished = true; break; }
                    if (again == 0) break;
                }
            }
            if (next_emit < n) emit_literal(n);   // done(): trailing literal
        }
#undef HASH
        if (lane == 0) s_d = d;
    }
    __syncthreads();
    d = s_d;
    const uint32_t crc = ((s_crc >> 15) | (s_crc << 17)) + 0xa282ead8u;

    // ---- frame chunk: stored when compression saved < 12.5 % ----
    const bool stored = d >= n - n / 8;
    if (stored) { for (uint32_t i = tid; i < n; i += kSnapThreads) dst[i] = in_s[i]; d = n; }
    if (tid == 0) {
        uint8_t *h = slot + 24;
        uint32_t chunk_len = 4 + d;
        h[0] = stored ? 1 : 0; h[1] = (uint8_t)chunk_len; h[2] = (uint8_t)(chunk_len >> 8); h[3] = (uint8_t)(chunk_len >> 16);
        h[4] = (uint8_t)crc; h[5] = (uint8_t)(crc >> 8); h[6] = (uint8_t)(crc >> 16); h[7] = (uint8_t)(crc >> 24);
        uint32_t total = 8 + d, hoff = 24;
        if (k == 0) {   // stream identifier in front of the block's first chunk
            const uint8_t ident[10] = {0xff, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};
            for (int i = 0; i < 10; i++) slot[14 + i] = ident[i];
            total += 10; hoff = 14;
        }
        out_len[2 * chunk] = total;
        out_len[2 * chunk + 1] = hoff;
    }
}

cudaError_t launch_snap(const SnapBatch &b, cudaStream_t st)
{
    if (b.nunits == 0) return cudaSuccess;
    if (b.timer) b.timer->start(KT_SNAP, st);
    static bool attr_done[64] = {};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (cur_dev >= 0 && cur_dev < 64 && !attr_done[cur_dev]) { cudaFuncSetAttribute(k_snap, cudaFuncAttributeMaxDynamicSharedMemorySize, kSnapSmem); attr_done[cur_dev] = true; }
    GZPB_LAUNCH(k_snap, b.nunits * b.cpu, kSnapThreads, kSnapSmem, st, b.in, b.unit_len, b.in_stride, b.cpu, b.out, b.out_stride, b.out_len);
    if (b.timer) b.timer->stop(st);
    return cudaGetLastError();
}

}  // namespace gzpb
