// inflate_kernels.cu — sm_100a kernel for the block-parallel DECODE path of gzp
// (SURVEY.md §8(f) rank 1): what a ParDecompress worker does per block,
//   /root/reference/src/par/decompress.rs:163-187  (get_footer_values, decode_block,
//   Check::update, the InvalidCheck comparison), with decode_block =
//   /root/reference/src/deflate.rs:384-405 (Mgzip) / :532-553 (Bgzf) -> raw DEFLATE inflate
//   into a buffer of exactly ISIZE bytes.
//
// One warp per gzp block (a BGZF / Mgzip member), four warps per CTA, thousands of
// blocks in flight.  DEFLATE symbol decoding is inherently sequential, so lane 0 of
// a warp decodes up to 32 tokens at a time (64-bit bit buffer refilled with 32-bit
// loads, 10-bit / 8-bit direct lookup tables in shared memory, canonical bit-by-bit
// fallback for longer codes) and the whole warp then writes them out: a prefix scan
// of the token lengths places every token, literals are stored in parallel, matches that only read
// bytes from before the batch are copied four at a time (loads before stores), the
// few that read the batch's own output follow in token order.  The warp finally computes the
// CRC-32 of its output (per-lane slices recombined in GF(2)) = `Check::update`, and
// compares it with the member's footer.  Decoded bytes land directly at their final
// stream offset (the host computed the offsets from the ISIZE fields), so there is
// no gather pass.  Bit-exact by construction: output must equal the original input
// (tests/test_gpu_inflate.py round trips + streams produced by zlib / the oracle).
#include <stdio.h>

#include "gzpb_common.cuh"
#include "deflate_kernels.cuh"
#include "inflate_kernels.cuh"

namespace gzpb {

__constant__ uint32_t c_icrc_tab[4][256];      // slicing-by-4, reflected 0xEDB88320
__constant__ uint16_t c_len_base[32] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 0, 0, 0};
__constant__ uint8_t c_len_extra[32] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0, 0};
__constant__ uint16_t c_dist_base[32] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 0, 0};
__constant__ uint8_t c_dist_extra[32] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 0, 0};

void upload_inflate_constants()
{
    static uint32_t tab[4][256];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ (kCrcPoly & (0u - (c & 1)));
        tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int s = 1; s < 4; s++) tab[s][i] = (tab[s - 1][i] >> 8) ^ tab[0][tab[s - 1][i] & 0xFF];
    cudaMemcpyToSymbol(c_icrc_tab, tab, sizeof tab);
}

constexpr int kInfWarps = 4;          // gzp blocks per CTA (one warp each)
constexpr int kLitBits = 10;          // direct-lookup bits of the litlen table
constexpr int kDistBits = 8;          // ... of the offset table (also holds the 7-bit precode table)
constexpr int kBatch = 32;            // tokens decoded by lane 0 per write-out round

// decode status of one block (device side); the host maps them onto GzpError
enum : int32_t { INF_OK = 0, INF_BAD_DATA = 1, INF_OUT_OVERRUN = 2, INF_IN_OVERRUN = 3, INF_BAD_CHECK = 4 };

struct InfWarp {
    uint16_t lit_tab[1 << kLitBits];   // (sym << 4) | len; 0 = longer code (or unused) -> canonical fallback
    uint16_t dist_tab[1 << kDistBits];
    uint16_t sorted[kNumLitlen + kNumOffset];   // symbols ordered by (len, sym): litlen part, then offset part
    uint16_t lcnt[16], dcnt[16];       // codes per length
    uint16_t first[16], start[16];     // builder scratch: first canonical code / first sorted index per length
    uint8_t lens[kNumLitlen + kNumOffset + 4];
    uint8_t plens[kNumPrecode + 1];
    uint32_t tok[kBatch];              // literal byte, or 0x80000000 | len << 16 | dist
    uint32_t pos[kBatch];              // output position of every token of the batch
    uint8_t idx[kBatch];               // lanes holding batch-independent matches, compacted
};

// unaligned little-endian 32-bit global load (two aligned loads; may touch up to 7 bytes past p)
__device__ __forceinline__ uint32_t gld32(const uint8_t *p)
{
    const uint32_t *w = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
    const uint32_t s = ((uintptr_t)p & 3) * 8;
    return __funnelshift_r(__ldg(w), __ldg(w + 1), s);
}

struct BitReader {
    uint64_t bb;          // bit buffer, next bit = bit 0
    uint32_t bc;          // valid bits
    const uint8_t *ip;    // next byte to load
    const uint8_t *lim;   // end of the member's payload: loads start below it (a truncated or hostile member never pulls
                          // the reader through memory; past `lim` the reader is fed zeros and the callers report INF_IN_OVERRUN)
    __device__ __forceinline__ void refill() { if (bc < 32) { const uint32_t v = ip < lim ? gld32(ip) : 0u; bb |= (uint64_t)v << bc; ip += 4; bc += 32; } }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(uint32_t n) { bb >>= n; bc -= n; }
    __device__ __forceinline__ uint32_t take(uint32_t n) { uint32_t v = peek(n); drop(n); return v; }
};

// Warp-cooperative canonical-Huffman table build (RFC 1951 §3.2.2): `lens[0..n)` ->
// direct table of 2^bits entries + (cnt, sorted) for the bit-by-bit fallback.
// Returns false (uniformly) when the code is over-subscribed.
__device__ bool build_table(const uint8_t *lens, uint32_t n, uint16_t *cnt, uint16_t *sorted, uint16_t *tab, uint32_t bits,
                            uint16_t *first, uint16_t *start, uint32_t lane)
{
    for (uint32_t i = lane; i < (1u << bits); i += 32) tab[i] = 0;
    if (lane < 16) cnt[lane] = 0;
    __syncwarp();
    uint32_t ok = 1;
    if (lane == 0) {
        for (uint32_t s = 0; s < n; s++) cnt[lens[s]]++;
        cnt[0] = 0;
        int left = 1;
        uint32_t code = 0, idx = 0;
        for (uint32_t l = 1; l <= 15; l++) {
            left = (left << 1) - (int)cnt[l];
            if (left < 0) ok = 0;
            code = (code + (l > 1 ? cnt[l - 1] : 0)) << 1;
            first[l] = (uint16_t)code; start[l] = (uint16_t)idx;
            idx += cnt[l];
        }
        if (ok) {
            uint16_t off[16];
            for (uint32_t l = 1; l <= 15; l++) off[l] = start[l];
            for (uint32_t s = 0; s < n; s++) { const uint32_t l = lens[s]; if (l) sorted[off[l]++] = (uint16_t)s; }
        }
    }
    ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
    if (!ok) return false;
    __syncwarp();
    uint32_t nsyms = 0;
    for (uint32_t l = 1; l <= 15; l++) nsyms += cnt[l];
    for (uint32_t i = lane; i < nsyms; i += 32) {
        const uint32_t sym = sorted[i], l = lens[sym];
        if (l <= bits) {
            const uint32_t code = (uint32_t)first[l] + (i - (uint32_t)start[l]);
            const uint32_t rev = __brev(code) >> (32 - l);
            const uint16_t e = (uint16_t)((sym << 4) | l);
            for (uint32_t k = rev; k < (1u << bits); k += 1u << l) tab[k] = e;
        }
    }
    __syncwarp();
    return true;
}

// one symbol: direct lookup, else the canonical walk (code lengths bits+1..15); returns -1 on an invalid code
__device__ __forceinline__ int decode_sym(BitReader &br, const uint16_t *tab, uint32_t bits, const uint16_t *cnt, const uint16_t *sorted)
{
    const uint32_t e = tab[br.peek(bits)];
    if (e) { br.drop(e & 15u); return (int)(e >> 4); }
    uint32_t code = 0, first = 0, index = 0;
    const uint32_t v = (uint32_t)br.bb;
    for (uint32_t l = 1; l <= 15; l++) {
        code |= (v >> (l - 1)) & 1u;
        const uint32_t c = cnt[l];
        if (code < first + c) { br.drop(l); return (int)sorted[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

__global__ void __launch_bounds__(kInfWarps * 32)
k_inflate(const uint8_t *__restrict__ comp, const InflateDesc *__restrict__ desc, uint32_t nblocks, uint8_t *out,
          int32_t *__restrict__ status, uint32_t *__restrict__ crc_found)
{
    __shared__ InfWarp s_w[kInfWarps];
    __shared__ uint32_t s_crc[4][256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 1024; i += kInfWarps * 32) s_crc[i >> 8][i & 255] = c_icrc_tab[i >> 8][i & 255];
    __syncthreads();
    const uint32_t blk = blockIdx.x * kInfWarps + warp;
    if (blk >= nblocks) return;
    InfWarp &S = s_w[warp];
    const InflateDesc d = desc[blk];
    const uint8_t *in0 = comp + d.in_off;
    const uint8_t *in_end = in0 + d.in_len;
    uint8_t *dst = out + d.out_off;
    const uint32_t out_len = d.out_len;

    BitReader br;
    br.bb = 0; br.bc = 0; br.ip = in0; br.lim = in_end;
    uint32_t opos = 0;          // bytes written so far (uniform across the warp)
    int32_t err = INF_OK;
    uint32_t bfinal = 0;

    while (!bfinal && err == INF_OK && out_len != 0) {
        // ---- block header (lane 0) ----
        uint32_t btype = 0, nlit = 0, ndist = 0;
        if (lane == 0) {
            if (br.ip > in_end + 8) err = INF_IN_OVERRUN;
            br.refill();
            bfinal = br.take(1);
            btype = br.take(2);
            if (btype == 3) err = INF_BAD_DATA;
            else if (btype == 2) {
                nlit = br.take(5) + 257; ndist = br.take(5) + 1;
                const uint32_t ncode = br.take(4) + 4;
                const uint8_t perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                for (uint32_t i = 0; i < 19; i++) S.plens[i] = 0;
                for (uint32_t i = 0; i < ncode; i++) { br.refill(); S.plens[perm[i]] = (uint8_t)br.take(3); }
                if (nlit > 286 || ndist > 30) err = INF_BAD_DATA;
            }
        }
        bfinal = __shfl_sync(0xFFFFFFFFu, bfinal, 0);
        btype = __shfl_sync(0xFFFFFFFFu, btype, 0);
        err = __shfl_sync(0xFFFFFFFFu, err, 0);
        if (err != INF_OK) break;

        if (btype == 0) {
            // ---- stored block: byte-align, LEN / NLEN, warp copy ----
            uint32_t len = 0;
            const uint8_t *src = nullptr;
            if (lane == 0) {
                br.drop(br.bc & 7u);
                br.refill();
                len = br.take(16);
                const uint32_t nlen = br.take(16);
                if ((len ^ nlen) != 0xFFFFu) err = INF_BAD_DATA;
                src = br.ip - (br.bc >> 3);            // whole bytes still buffered belong to the stored data
                if (src + len > in_end) err = INF_IN_OVERRUN;
                else if (opos + len > out_len) err = INF_OUT_OVERRUN;
                br.ip = src + len; br.bb = 0; br.bc = 0;
            }
            err = __shfl_sync(0xFFFFFFFFu, err, 0);
            if (err != INF_OK) break;
            len = __shfl_sync(0xFFFFFFFFu, len, 0);
            src = (const uint8_t *)__shfl_sync(0xFFFFFFFFu, (unsigned long long)(uintptr_t)src, 0);
            for (uint32_t i = lane; i < len; i += 32) dst[opos + i] = __ldg(src + i);
            opos += len;
            __syncwarp();
            continue;
        }

        // ---- code lengths ----
        if (btype == 1) {
            for (uint32_t s = lane; s < kNumLitlen; s += 32) S.lens[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
            S.lens[kNumLitlen + lane] = (lane < 30) ? 5 : 0;
            nlit = kNumLitlen; ndist = 30;
            __syncwarp();
        } else {
            nlit = __shfl_sync(0xFFFFFFFFu, nlit, 0);
            ndist = __shfl_sync(0xFFFFFFFFu, ndist, 0);
            if (!build_table(S.plens, kNumPrecode, S.dcnt, S.sorted, S.dist_tab, 7, S.first, S.start, lane)) { err = INF_BAD_DATA; break; }
            if (lane == 0) {
                uint32_t i = 0;
                const uint32_t total = nlit + ndist;
                while (i < total) {
                    br.refill();
                    const int sym = decode_sym(br, S.dist_tab, 7, S.dcnt, S.sorted);
                    if (sym < 0) { err = INF_BAD_DATA; break; }
                    if (sym < 16) { S.lens[i++] = (uint8_t)sym; continue; }
                    uint32_t rep, val = 0;
                    if (sym == 16) { if (i == 0) { err = INF_BAD_DATA; break; } val = S.lens[i - 1]; rep = 3 + br.take(2); }
                    else if (sym == 17) rep = 3 + br.take(3);
                    else rep = 11 + br.take(7);
                    if (i + rep > total) { err = INF_BAD_DATA; break; }
                    while (rep--) S.lens[i++] = (uint8_t)val;
                }
                if (err == INF_OK && br.ip > in_end + 8) err = INF_IN_OVERRUN;  // the header ran past the payload (fed zeros since)
                if (err == INF_OK && S.lens[256] == 0) err = INF_BAD_DATA;    // no end-of-block code
            }
            err = __shfl_sync(0xFFFFFFFFu, err, 0);
            if (err != INF_OK) break;
            __syncwarp();
        }
        if (!build_table(S.lens, nlit, S.lcnt, S.sorted, S.lit_tab, kLitBits, S.first, S.start, lane)) { err = INF_BAD_DATA; break; }
        if (!build_table(S.lens + nlit, ndist, S.dcnt, S.sorted + kNumLitlen, S.dist_tab, kDistBits, S.first, S.start, lane)) { err = INF_BAD_DATA; break; }

        // ---- symbols: lane 0 decodes a batch, the warp writes it out ----
        bool eob = false;
        while (!eob && err == INF_OK) {
            uint32_t ntok = 0, nbytes = 0;
            if (lane == 0) {
                while (ntok < (uint32_t)kBatch) {
                    if (br.ip > in_end + 8) { err = INF_IN_OVERRUN; break; }
                    br.refill();
                    const int sym = decode_sym(br, S.lit_tab, kLitBits, S.lcnt, S.sorted);
                    if (sym < 0) { err = INF_BAD_DATA; break; }
                    if (sym < 256) {
                        if (opos + nbytes + 1 > out_len) { err = INF_OUT_OVERRUN; break; }
                        S.tok[ntok++] = (uint32_t)sym; nbytes++;
                        continue;
                    }
                    if (sym == 256) { eob = true; break; }
                    if (sym > 285) { err = INF_BAD_DATA; break; }
                    const uint32_t ls = (uint32_t)sym - 257;
                    const uint32_t len = c_len_base[ls] + br.take(c_len_extra[ls]);
                    br.refill();
                    const int ds = decode_sym(br, S.dist_tab, kDistBits, S.dcnt, S.sorted + kNumLitlen);
                    if (ds < 0 || ds > 29) { err = INF_BAD_DATA; break; }
                    const uint32_t dist = c_dist_base[ds] + br.take(c_dist_extra[ds]);
                    if (dist > opos + nbytes) { err = INF_BAD_DATA; break; }         // reaches before the start of the block
                    if (opos + nbytes + len > out_len) { err = INF_OUT_OVERRUN; break; }
                    S.tok[ntok++] = 0x80000000u | (len << 16) | dist; nbytes += len;
                }
            }
            __syncwarp();
            ntok = __shfl_sync(0xFFFFFFFFu, ntok, 0);
            eob = __shfl_sync(0xFFFFFFFFu, (uint32_t)eob, 0) != 0;
            err = __shfl_sync(0xFFFFFFFFu, err, 0);
            // ---- write-out (tokens decoded before an error are still valid output) ----
            const uint32_t t = lane < ntok ? S.tok[lane] : 0u;
            const bool is_match = (lane < ntok) && (t & 0x80000000u);
            const uint32_t mylen = lane < ntok ? (is_match ? (t >> 16) & 0x1FFu : 1u) : 0u;
            const uint32_t mydist = t & 0xFFFFu;
            uint32_t incl = mylen;
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= (uint32_t)o) incl += v; }
            const uint32_t mypos = opos + incl - mylen;
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            // Matches whose source lies entirely before this batch depend on nothing written in this
            // round: they are copied first, four at a time with all loads issued before the stores, so
            // their L2 round trips overlap.  The rest (sources inside the batch) follow in token order.
            const bool indep = is_match && (mypos - mydist + min(mylen, mydist) <= opos);
            const uint32_t ind = __ballot_sync(0xFFFFFFFFu, indep);
            uint32_t dep = __ballot_sync(0xFFFFFFFFu, is_match && !indep);
            S.pos[lane] = mypos;
            if (indep) S.idx[__popc(ind & lanemask_lt())] = (uint8_t)lane;
            if (lane < ntok && !is_match) dst[mypos] = (uint8_t)t;
            __syncwarp();
            const uint32_t nind = __popc(ind);
            for (uint32_t k = 0; k < nind; k += 4) {
                uint32_t ml[4], mp[4], ms[4];
                uint8_t v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    ml[u] = 0; mp[u] = 0; ms[u] = 0;
                    if (k + u < nind) {
                        const uint32_t j = S.idx[k + u], tj = S.tok[j];
                        const uint32_t lj = (tj >> 16) & 0x1FFu, dj = tj & 0xFFFFu;
                        ml[u] = lj; mp[u] = S.pos[j];
                        ms[u] = mp[u] - dj + (dj >= lj ? lane : lane % dj);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = (lane < ml[u]) ? dst[ms[u]] : (uint8_t)0;
#pragma unroll
                for (int u = 0; u < 4; u++) if (lane < ml[u]) dst[mp[u] + lane] = v[u];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (ml[u] > 32) {     // long match: the remaining bytes
                        const uint32_t j = S.idx[k + u], dj = S.tok[j] & 0xFFFFu;
                        const uint8_t *from = dst + mp[u] - dj;
                        if (dj >= ml[u]) { for (uint32_t b = lane + 32; b < ml[u]; b += 32) dst[mp[u] + b] = from[b]; }
                        else { for (uint32_t b = lane + 32; b < ml[u]; b += 32) dst[mp[u] + b] = from[b % dj]; }
                    }
                }
            }
            __syncwarp();
            while (dep) {
                const int j = __ffs(dep) - 1;
                dep &= dep - 1;
                const uint32_t tj = S.tok[j], pj = S.pos[j];
                const uint32_t lj = (tj >> 16) & 0x1FFu, dj = tj & 0xFFFFu;
                const uint8_t *from = dst + pj - dj;
                if (dj >= lj) { for (uint32_t b = lane; b < lj; b += 32) dst[pj + b] = from[b]; }
                else { for (uint32_t b = lane; b < lj; b += 32) dst[pj + b] = from[b % dj]; }    // overlapping copy = periodic extension
                __syncwarp();
            }
            opos += total;
        }
    }

    // ---- tail: consumed-input check, zero fill of a short output (decode_block's vec![0; orig_size]) ----
    if (lane == 0 && err == INF_OK && out_len != 0) {
        const uint8_t *used = br.ip - (br.bc >> 3);
        if (used > in_end) err = INF_IN_OVERRUN;
    }
    err = __shfl_sync(0xFFFFFFFFu, err, 0);
    for (uint32_t i = opos + lane; i < out_len; i += 32) dst[i] = 0;
    __syncwarp();

    // ---- Check::update over the decoded block: CRC-32, one contiguous slice per lane ----
    uint32_t crc = 0;
    if (out_len) {
        const uint32_t per = ((out_len + 31) / 32 + 3) & ~3u;
        const uint32_t beg = min(out_len, lane * per), end = min(out_len, beg + per);
        uint32_t c = ~0u, pos = beg;
        while (pos < end && ((uintptr_t)(dst + pos) & 3)) { c = (c >> 8) ^ s_crc[0][(c ^ dst[pos]) & 0xFF]; pos++; }
        for (; pos + 4 <= end; pos += 4) {
            c ^= *(const uint32_t *)(dst + pos);
            c = s_crc[3][c & 0xFF] ^ s_crc[2][(c >> 8) & 0xFF] ^ s_crc[1][(c >> 16) & 0xFF] ^ s_crc[0][c >> 24];
        }
        while (pos < end) { c = (c >> 8) ^ s_crc[0][(c ^ dst[pos]) & 0xFF]; pos++; }
        c = ~c;
        // crc(A || B) = crc(A) * x^(8 |B|) + crc(B): shift this slice's CRC by the bytes that follow it
        uint32_t part = (end > beg) ? ((out_len - end) ? gf2_mulmod(c, gf2_xpow8(out_len - end, kCrcPoly), kCrcPoly) : c) : 0u;
        for (int o = 16; o; o >>= 1) part ^= __shfl_xor_sync(0xFFFFFFFFu, part, o);
        crc = part;
    }
    if (lane == 0) {
        if (err == INF_OK && crc != d.crc) err = INF_BAD_CHECK;
        status[blk] = err;
        crc_found[blk] = crc;
    }
}

cudaError_t launch_inflate(const InflateBatch &b, cudaStream_t st)
{
    if (b.nblocks == 0) return cudaSuccess;
    if (b.timer) b.timer->start(KT_INFLATE, st);
    GZPB_LAUNCH(k_inflate, (b.nblocks + kInfWarps - 1) / kInfWarps, kInfWarps * 32, 0, st, b.comp, b.desc, b.nblocks, b.out, b.status, b.crc_found);
    if (b.timer) b.timer->stop(st);
    return cudaGetLastError();
}

}  // namespace gzpb
