/*
 * oracle/snappy_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates the encoder gzp's Snap format calls:
 * `snap::read::FrameEncoder::new(input).read_to_end()`
 * (/root/reference/src/snap.rs:61-74).  The arithmetic lives in the `snap`
 * crate 1.1.1 (/root/reference/Cargo.lock:818-820), not vendored under
 * /root/reference; this file restates its published algorithm (the Go/C++
 * Snappy block encoder it ports: greedy parse, u16 hash table of up to 2^14
 * entries updated only at visited positions, accelerating skip, 15-byte input
 * margin; framing per the Snappy framing-format description).
 * PARITY UNPINNED against the crate's bytes; anchored on stock Snappy decoders
 * (pyarrow raw codec) and the CRC-32C known-answer vector.
 */
#include <string.h>
#include "oracle.h"

#define MAX_BLOCK 65536u
#define INPUT_MARGIN 15u
#define MIN_NON_LITERAL 17u
#define MAX_TABLE 16384u

size_t oracle_snappy_max_compress_len(size_t n) { return 32 + n + n / 6; }

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t ld64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

typedef struct { const uint8_t *src; size_t n; uint8_t *dst; size_t d; size_t s, next_emit; } blk_t;

static void emit_literal(blk_t *b, size_t lit_end)
{
    size_t start = b->next_emit, len = lit_end - start, n = len - 1;
    if (n <= 59) { b->dst[b->d++] = (uint8_t)(n << 2); }
    else if (n < 256) { b->dst[b->d++] = 60 << 2; b->dst[b->d++] = (uint8_t)n; }
    else { b->dst[b->d++] = 61 << 2; b->dst[b->d++] = (uint8_t)n; b->dst[b->d++] = (uint8_t)(n >> 8); }
    memcpy(b->dst + b->d, b->src + start, len);
    b->d += len;
}

static void emit_copy2(blk_t *b, size_t off, size_t len)
{
    b->dst[b->d++] = (uint8_t)(((len - 1) << 2) | 2);
    b->dst[b->d++] = (uint8_t)off; b->dst[b->d++] = (uint8_t)(off >> 8);
}

static void emit_copy(blk_t *b, size_t off, size_t len)
{
    while (len >= 68) { emit_copy2(b, off, 64); len -= 64; }
    if (len > 64) { emit_copy2(b, off, 60); len -= 60; }
    if (len <= 11 && off <= 2047) {
        b->dst[b->d++] = (uint8_t)(((off >> 8) << 5) | ((len - 4) << 2) | 1);
        b->dst[b->d++] = (uint8_t)off;
    } else {
        emit_copy2(b, off, len);
    }
}

static void done(blk_t *b) { if (b->next_emit < b->n) emit_literal(b, b->n); }

static void compress_block(blk_t *b, uint16_t *table, unsigned shift)
{
#define HASH(x) ((uint32_t)((x) * 0x1E35A7BDu) >> shift)
    const uint8_t *src = b->src;
    size_t s_limit = b->n - INPUT_MARGIN;
    b->s = 1;
    uint32_t next_hash = HASH(ld32(src + b->s));
    for (;;) {
        uint32_t skip = 32;
        size_t candidate, s_next = b->s;
        for (;;) {
            b->s = s_next;
            uint32_t step = skip >> 5;
            s_next = b->s + step;
            skip += step;
            if (s_next > s_limit) { done(b); return; }
            candidate = table[next_hash];
            table[next_hash] = (uint16_t)b->s;
            next_hash = HASH(ld32(src + s_next));
            if (ld32(src + b->s) == ld32(src + candidate)) break;
        }
        emit_literal(b, b->s);
        for (;;) {
            size_t base = b->s;
            b->s += 4;
            size_t cand = candidate + 4;
            while (b->s < b->n && src[b->s] == src[cand]) { b->s++; cand++; }
            emit_copy(b, base - candidate, b->s - base);
            b->next_emit = b->s;
            if (b->s >= s_limit) { done(b); return; }
            uint64_t x = ld64(src + b->s - 1);
            table[HASH((uint32_t)x)] = (uint16_t)(b->s - 1);
            uint32_t cur_hash = HASH((uint32_t)(x >> 8));
            candidate = table[cur_hash];
            table[cur_hash] = (uint16_t)b->s;
            if ((uint32_t)(x >> 8) != ld32(src + candidate)) {
                next_hash = HASH((uint32_t)(x >> 16));
                b->s++;
                break;
            }
        }
    }
#undef HASH
}

/* snap::raw::Encoder::compress */
size_t oracle_snappy_raw(const uint8_t *in, size_t n, uint8_t *out)
{
    size_t d = 0;
    uint64_t v = n;
    while (v >= 0x80) { out[d++] = (uint8_t)(v | 0x80); v >>= 7; }
    out[d++] = (uint8_t)v;
    uint16_t table[MAX_TABLE];
    while (n) {
        size_t bl = n > MAX_BLOCK ? MAX_BLOCK : n;
        blk_t b = {in, bl, out, d, 0, 0};
        if (bl < MIN_NON_LITERAL) {
            emit_literal(&b, bl);
        } else {
            unsigned shift = 32 - 8; size_t tsz = 256;
            while (tsz < MAX_TABLE && tsz < bl) { shift--; tsz *= 2; }
            memset(table, 0, tsz * sizeof(uint16_t));
            compress_block(&b, table, shift);
        }
        d = b.d; in += bl; n -= bl;
    }
    return d;
}

/* snap::read::FrameEncoder over the whole input: stream identifier (only when
 * there is at least one byte), then one chunk per <= 65536 source bytes. */
size_t oracle_snappy_frame(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap)
{
    static const uint8_t ident[10] = {0xff, 0x06, 0x00, 0x00, 's', 'N', 'a', 'P', 'p', 'Y'};
    size_t d = 0;
    uint8_t tmp[32 + MAX_BLOCK + MAX_BLOCK / 6];
    if (n == 0) return 0;
    if (out_cap < 10) return 0;
    memcpy(out, ident, 10); d = 10;
    while (n) {
        size_t bl = n > MAX_BLOCK ? MAX_BLOCK : n;
        uint32_t crc = oracle_crc32c_masked(in, bl);
        size_t cl = oracle_snappy_raw(in, bl, tmp);
        int stored = cl >= bl - bl / 8;
        size_t body = stored ? bl : cl, chunk_len = 4 + body;
        if (d + 8 + body > out_cap) return 0;
        out[d++] = stored ? 0x01 : 0x00;
        out[d++] = (uint8_t)chunk_len; out[d++] = (uint8_t)(chunk_len >> 8); out[d++] = (uint8_t)(chunk_len >> 16);
        out[d++] = (uint8_t)crc; out[d++] = (uint8_t)(crc >> 8); out[d++] = (uint8_t)(crc >> 16); out[d++] = (uint8_t)(crc >> 24);
        memcpy(out + d, stored ? in : tmp, body); d += body;
        in += bl; n -= bl;
    }
    return d;
}
