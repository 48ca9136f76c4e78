/*
 * oracle/format_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Restates gzp's FormatSpec::encode / header / footer for each format:
 *   Bgzf       /root/reference/src/deflate.rs:613-634, src/bgzf.rs:204-237, 274-303, 24-38
 *   Mgzip      /root/reference/src/deflate.rs:463-481, src/mgzip.rs:187-218, 246-275
 *   Gzip       /root/reference/src/deflate.rs:86-142
 *   Zlib       /root/reference/src/deflate.rs:193-251
 *   RawDeflate /root/reference/src/deflate.rs:304-331
 *   Snap       /root/reference/src/snap.rs:61-82
 * Container bytes are pinned by the reference's own constants (BGZF_EOF,
 * header recipes).  The DEFLATE payload comes from deflate_oracle.c.  For
 * Gzip/Zlib/RawDeflate the reference calls zlib-ng (not vendored); north_star
 * only requires those streams to be decodable by stock gzip, so the payload is
 * this repo's own engine primed with the 32 KiB dictionary and terminated by a
 * sync-flush marker, exactly as `set_dictionary` + `FlushCompress::Sync` /
 * `Finish` arrange it.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define E_BLOCKSIZE (-3)
#define E_COMPRESS (-4)
#define E_LEVEL (-5)
#define E_INVAL (-9)

static size_t extra_amount(size_t n) { size_t e = (size_t)((double)n * 0.1); return e > 128 ? e : 128; }

size_t oracle_encode_capacity(int format, size_t n)
{
    switch (format) {
    case ORACLE_FMT_BGZF: return 18 + n + extra_amount(n) + 8 + 28;
    case ORACLE_FMT_MGZIP: return 20 + n + extra_amount(n) + 8;
    case ORACLE_FMT_SNAP: return 10 + ((n + 65535) / 65536) * 8 + oracle_snappy_max_compress_len(n) + 64;
    default: return n + extra_amount(n);
    }
}

static int xfl(int level) { return level >= 9 ? 2 : level <= 1 ? 4 : 0; }

static void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put32(uint8_t *p, uint32_t v) { put16(p, v); put16(p + 2, v >> 16); }

static const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                                     0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

long oracle_encode_block(int format, int level, const uint8_t *in, size_t n, const uint8_t *dict,
                         size_t dict_len, int is_last, uint8_t *out, size_t out_cap)
{
    if (format == ORACLE_FMT_SNAP) {
        size_t r = oracle_snappy_frame(in, n, out, out_cap);
        return (long)r;
    }
    if (!oracle_level_supported(level)) return E_LEVEL;
    if (format == ORACLE_FMT_BGZF || format == ORACLE_FMT_MGZIP) {
        size_t hs = format == ORACLE_FMT_BGZF ? 18 : 20;
        size_t avail = n + extra_amount(n) + 8; /* buffer[HEADER..] of the zero-filled Vec */
        if (out_cap < hs + avail) return E_INVAL;
        size_t w = oracle_deflate(in, n, level, out + hs, avail);
        if (w == 0) return E_COMPRESS;
        /* bgzf.rs:218-223 tests `>= 65536` on the payload, but BSIZE = payload + 26 - 1 is a u16 (bgzf.rs:299): payloads of
         * 65511..65535 bytes wrap it and the reference (release build) hands back a corrupt member as Ok.  Deliberate
         * deviation, shared with the CUDA path: a member that does not fit 65536 bytes is BlockSizeExceeded. */
        if (format == ORACLE_FMT_BGZF && w + 26 > 65536) return E_BLOCKSIZE;
        uint8_t *h = out;
        h[0] = 31; h[1] = 139; h[2] = 8; h[3] = 4; put32(h + 4, 0); h[8] = (uint8_t)xfl(level); h[9] = 255;
        if (format == ORACLE_FMT_BGZF) {
            put16(h + 10, 6); h[12] = 'B'; h[13] = 'C'; put16(h + 14, 2);
            put16(h + 16, (uint32_t)((uint16_t)w + 26 - 1));
        } else {
            put16(h + 10, 8); h[12] = 'I'; h[13] = 'G'; put16(h + 14, 4);
            put32(h + 16, (uint32_t)w + 28);
        }
        put32(out + hs + w, oracle_crc32(0, in, n));
        put32(out + hs + w + 4, (uint32_t)n);
        size_t tot = hs + w + 8;
        if (format == ORACLE_FMT_BGZF && is_last) {
            if (out_cap < tot + 28) return E_INVAL;
            memcpy(out + tot, BGZF_EOF, 28); tot += 28;
        }
        return (long)tot;
    }
    if (format == ORACLE_FMT_GZIP || format == ORACLE_FMT_ZLIB || format == ORACLE_FMT_RAWDEFLATE) {
        int flush = (format == ORACLE_FMT_RAWDEFLATE) ? 1 : (is_last ? 0 : 1);
        size_t w;
        if (dict_len) {
            /* the engine wants dictionary and data contiguous */
            uint8_t *scratch = malloc(dict_len + n + 16);
            if (!scratch) return E_INVAL;
            memcpy(scratch, dict, dict_len); memcpy(scratch + dict_len, in, n);
            w = oracle_deflate_ex(scratch, dict_len, n, level, flush, out, out_cap, NULL);
            free(scratch);
        } else {
            w = oracle_deflate_ex(in, 0, n, level, flush, out, out_cap, NULL);
        }
        if (w == 0) return E_COMPRESS;
        return (long)w;
    }
    return E_INVAL;
}

size_t oracle_header(int format, int level, uint8_t *out)
{
    if (format == ORACLE_FMT_GZIP) {
        out[0] = 31; out[1] = 139; out[2] = 8; out[3] = 0; put32(out + 4, 0); out[8] = (uint8_t)xfl(level); out[9] = 255;
        return 10;
    }
    if (format == ORACLE_FMT_ZLIB) {
        uint32_t cv = level >= 9 ? 3u << 6 : level == 1 ? 0 : level >= 6 ? 1u << 6 : 2u << 6;
        uint32_t head = (0x78u << 8) + cv;
        head += 31 - (head % 31);
        out[0] = (uint8_t)(head >> 8); out[1] = (uint8_t)head;
        return 2;
    }
    return 0;
}

size_t oracle_footer(int format, uint32_t sum, uint32_t amount, uint8_t *out)
{
    if (format == ORACLE_FMT_GZIP) { put32(out, sum); put32(out + 4, amount); return 8; }
    if (format == ORACLE_FMT_ZLIB) { out[0] = (uint8_t)(sum >> 24); out[1] = (uint8_t)(sum >> 16); out[2] = (uint8_t)(sum >> 8); out[3] = (uint8_t)sum; return 4; }
    return 0;
}
