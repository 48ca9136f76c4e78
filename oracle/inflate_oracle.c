/* inflate_oracle.c — CPU restatement of the block DECODE path.  TEST INFRASTRUCTURE ONLY.
 *
 * What a ParDecompress worker does per block (/root/reference/src/par/decompress.rs:163-187):
 * footer values (src/lib.rs:440-447), decode_block = raw DEFLATE inflate into exactly ISIZE
 * bytes (src/deflate.rs:384-405, 532-553), CRC-32 of the result against the footer
 * (:173-181); and the reader loop around it (:190-207: check_header / get_block_size,
 * src/deflate.rs:407-422, 555-570).  The inflate itself lives in un-vendored libdeflate
 * (Cargo.lock:414-430); restated here from RFC 1951 as a plain bit-by-bit canonical
 * decoder (deliberately a different construction from the CUDA kernel's lookup tables).
 * Pinned by tests/test_oracle.py against zlib-produced streams and stock `gzip`.
 */
#include <string.h>

#include "oracle.h"

typedef struct { const uint8_t *in; size_t n, pos; uint32_t bitbuf; int bitcnt; int err; } bitsrc_t;

static uint32_t getbits(bitsrc_t *s, int need)
{
    uint32_t val = s->bitbuf;
    while (s->bitcnt < need) {
        if (s->pos >= s->n) { s->err = 1; return 0; }
        val |= (uint32_t)s->in[s->pos++] << s->bitcnt;
        s->bitcnt += 8;
    }
    s->bitbuf = need == 32 ? 0 : val >> need;
    s->bitcnt -= need;
    return need == 32 ? val : val & ((1u << need) - 1);
}

typedef struct { uint16_t count[16]; uint16_t symbol[288]; } huff_t;

static int build(huff_t *h, const uint8_t *lens, int n)
{
    uint16_t offs[16];
    memset(h->count, 0, sizeof h->count);
    for (int s = 0; s < n; s++) h->count[lens[s]]++;
    if (h->count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= h->count[l]; if (left < 0) return left; }
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h->count[l];
    for (int s = 0; s < n; s++) if (lens[s]) h->symbol[offs[lens[s]]++] = (uint16_t)s;
    return left;
}

static int decode(bitsrc_t *s, const huff_t *h)
{
    int code = 0, first = 0, index = 0;
    for (int l = 1; l <= 15; l++) {
        code |= (int)getbits(s, 1);
        if (s->err) return -1;
        int count = h->count[l];
        if (code - count < first) return h->symbol[index + (code - first)];
        index += count; first += count; first <<= 1; code <<= 1;
    }
    return -1;
}

static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* raw DEFLATE -> out (at most out_cap bytes).  Returns bytes produced, or -1 corrupt data,
 * -2 output does not fit, -3 input exhausted. */
long oracle_inflate(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap)
{
    bitsrc_t s = {in, n, 0, 0, 0, 0};
    size_t op = 0;
    int last;
    do {
        last = (int)getbits(&s, 1);
        int type = (int)getbits(&s, 2);
        if (s.err) return -3;
        if (type == 0) {
            s.bitbuf = 0; s.bitcnt = 0;
            if (s.pos + 4 > n) return -3;
            unsigned len = in[s.pos] | (in[s.pos + 1] << 8), nlen = in[s.pos + 2] | (in[s.pos + 3] << 8);
            s.pos += 4;
            if ((len ^ nlen) != 0xFFFF) return -1;
            if (s.pos + len > n) return -3;
            if (op + len > out_cap) return -2;
            memcpy(out + op, in + s.pos, len);
            op += len; s.pos += len;
        } else if (type == 1 || type == 2) {
            huff_t lh, dh;
            uint8_t lens[320];
            if (type == 1) {
                int i = 0;
                for (; i < 144; i++) lens[i] = 8;
                for (; i < 256; i++) lens[i] = 9;
                for (; i < 280; i++) lens[i] = 7;
                for (; i < 288; i++) lens[i] = 8;
                build(&lh, lens, 288);
                for (i = 0; i < 30; i++) lens[i] = 5;
                build(&dh, lens, 30);
            } else {
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                int nlen = (int)getbits(&s, 5) + 257, ndist = (int)getbits(&s, 5) + 1, ncode = (int)getbits(&s, 4) + 4;
                if (s.err) return -3;
                if (nlen > 286 || ndist > 30) return -1;
                int i;
                for (i = 0; i < ncode; i++) lens[order[i]] = (uint8_t)getbits(&s, 3);
                for (; i < 19; i++) lens[order[i]] = 0;
                if (s.err) return -3;
                if (build(&lh, lens, 19) < 0) return -1;
                i = 0;
                while (i < nlen + ndist) {
                    int sym = decode(&s, &lh);
                    if (sym < 0) return s.err ? -3 : -1;
                    if (sym < 16) lens[i++] = (uint8_t)sym;
                    else {
                        int len = 0, rep;
                        if (sym == 16) { if (i == 0) return -1; len = lens[i - 1]; rep = 3 + (int)getbits(&s, 2); }
                        else if (sym == 17) rep = 3 + (int)getbits(&s, 3);
                        else rep = 11 + (int)getbits(&s, 7);
                        if (s.err) return -3;
                        if (i + rep > nlen + ndist) return -1;
                        while (rep--) lens[i++] = (uint8_t)len;
                    }
                }
                if (lens[256] == 0) return -1;
                if (build(&lh, lens, nlen) < 0) return -1;
                if (build(&dh, lens + nlen, ndist) < 0) return -1;
            }
            for (;;) {
                int sym = decode(&s, &lh);
                if (sym < 0) return s.err ? -3 : -1;
                if (sym < 256) { if (op >= out_cap) return -2; out[op++] = (uint8_t)sym; }
                else if (sym == 256) break;
                else {
                    sym -= 257;
                    if (sym >= 29) return -1;
                    unsigned len = LBASE[sym] + getbits(&s, LEXT[sym]);
                    int ds = decode(&s, &dh);
                    if (ds < 0) return s.err ? -3 : -1;
                    if (ds >= 30) return -1;
                    unsigned dist = DBASE[ds] + getbits(&s, DEXT[ds]);
                    if (s.err) return -3;
                    if (dist > op) return -1;
                    if (op + len > out_cap) return -2;
                    for (unsigned k = 0; k < len; k++) { out[op] = out[op - dist]; op++; }
                }
            }
        } else return -1;
    } while (!last);
    return (long)op;
}

static size_t hdr_size(int format) { return format == ORACLE_FMT_BGZF ? 18 : format == ORACLE_FMT_MGZIP ? 20 : 0; }

/* check_header + get_block_size (deflate.rs:407-422, 555-570): member size or negative status
 * (-11 InvalidHeader, -6 short input, -9 not a block format). */
long oracle_block_size(int format, const uint8_t *h, size_t avail)
{
    size_t hs = hdr_size(format);
    if (!hs) return -9;
    if (avail < hs) return -6;
    if ((h[3] & 4) != 4) return -11;
    if (format == ORACLE_FMT_BGZF) {
        if (h[12] != 'B' || h[13] != 'C') return -11;
        return (long)(h[16] | (h[17] << 8)) + 1;
    }
    if (h[12] != 'I' || h[13] != 'G') return -11;
    return (long)((uint32_t)h[16] | ((uint32_t)h[17] << 8) | ((uint32_t)h[18] << 16) | ((uint32_t)h[19] << 24));
}

/* ParDecompress over an in-memory input.  Status codes follow include/gzpb.h:
 * 0 ok, -6 truncated member, -11 bad header, -12 CRC mismatch (found/expected reported),
 * -13 corrupt DEFLATE data, -14 bad block size, -4 output capacity. */
int oracle_decode_stream(int format, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len,
                         uint32_t *found, uint32_t *expected)
{
    size_t hs = hdr_size(format), pos = 0, op = 0;
    if (!hs) return -9;
    while (n - pos >= hs) {
        long size = oracle_block_size(format, in + pos, n - pos);
        if (size < 0) { *out_len = op; return (int)size; }
        if ((size_t)size < hs + 8) { *out_len = op; return -14; }
        if ((size_t)size > n - pos) { *out_len = op; return -6; }
        const uint8_t *f = in + pos + size - 8;
        uint32_t crc = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
        uint32_t isize = (uint32_t)f[4] | ((uint32_t)f[5] << 8) | ((uint32_t)f[6] << 16) | ((uint32_t)f[7] << 24);
        if (op + isize > out_cap) { *out_len = op; return -4; }
        if (isize) {
            memset(out + op, 0, isize);                                  /* vec![0; orig_size] */
            long r = oracle_inflate(in + pos + hs, (size_t)size - hs - 8, out + op, isize);
            if (r < 0) { *out_len = op; return -13; }
        }
        uint32_t c = oracle_crc32(0, out + op, isize);
        if (c != crc) { if (found) *found = c; if (expected) *expected = crc; *out_len = op; return -12; }
        op += isize; pos += (size_t)size;
    }
    *out_len = op;
    return 0;
}
