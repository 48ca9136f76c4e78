/*
 * oracle/check_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Checksums behind gzp's `Check` trait (/root/reference/src/check.rs:16-35):
 *   Crc32  (check.rs:133-164, flate2::Crc -> crc32fast 1.5.0): CRC-32/ISO-HDLC,
 *          reflected poly 0xEDB88320, plus `combine(crc_a, crc_b, len_b)`.
 *   LibDeflateCrc (check.rs:39-82): same polynomial.
 *   Adler32 (check.rs:86-129, zlib-ng adler32 / adler32_combine).
 * CRC-32C (Castagnoli, reflected 0x82F63B78) + the Snappy mask used by
 * snap 1.1.1's frame format (/root/reference/src/snap.rs:70-72).
 * Pinned by known-answer vectors: crc32("123456789")=cbf43926,
 * crc32c("123456789")=e3069283 (masked c78ab0e5), shakespeare.txt crc32
 * 6ae3de4c / adler32 af31c707, and python zlib.
 */
#include "oracle.h"

static uint32_t T[8][256], TC[8][256];
static int ready;

static void init(void)
{
    if (ready) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i, d = i;
        for (int k = 0; k < 8; k++) { c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1))); d = (d >> 1) ^ (0x82F63B78u & (0u - (d & 1))); }
        T[0][i] = c; TC[0][i] = d;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int s = 1; s < 8; s++) {
            T[s][i] = (T[s - 1][i] >> 8) ^ T[0][T[s - 1][i] & 0xFF];
            TC[s][i] = (TC[s - 1][i] >> 8) ^ TC[0][TC[s - 1][i] & 0xFF];
        }
    ready = 1;
}

static uint32_t crc_generic(uint32_t (*tab)[256], uint32_t crc, const uint8_t *p, size_t n)
{
    crc = ~crc;
    while (n >= 8) {
        uint32_t a = (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24;
        a ^= crc;
        crc = tab[7][a & 0xFF] ^ tab[6][(a >> 8) & 0xFF] ^ tab[5][(a >> 16) & 0xFF] ^ tab[4][a >> 24] ^
              tab[3][p[4]] ^ tab[2][p[5]] ^ tab[1][p[6]] ^ tab[0][p[7]];
        p += 8; n -= 8;
    }
    while (n--) crc = (crc >> 8) ^ tab[0][(crc ^ *p++) & 0xFF];
    return ~crc;
}

uint32_t oracle_crc32(uint32_t crc, const uint8_t *p, size_t n) { init(); return crc_generic(T, crc, p, n); }
uint32_t oracle_crc32c(uint32_t crc, const uint8_t *p, size_t n) { init(); return crc_generic(TC, crc, p, n); }

uint32_t oracle_crc32c_masked(const uint8_t *p, size_t n)
{
    uint32_t c = oracle_crc32c(0, p, n);
    return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

/* crc(A||B) from crc(A), crc(B), len(B): multiply crc(A) by x^(8*len(B)) mod P
 * using repeated squaring of the "times x" operator (the zlib construction that
 * crc32fast::Hasher::combine also implements). */
static uint32_t gf2_times(const uint32_t *mat, uint32_t vec)
{
    uint32_t s = 0;
    while (vec) { if (vec & 1) s ^= *mat; vec >>= 1; mat++; }
    return s;
}
static void gf2_square(uint32_t *sq, const uint32_t *mat) { for (int n = 0; n < 32; n++) sq[n] = gf2_times(mat, mat[n]); }

uint32_t oracle_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2)
{
    uint32_t even[32], odd[32];
    if (len2 == 0) return crc1;
    odd[0] = 0xEDB88320u;
    uint32_t row = 1;
    for (int n = 1; n < 32; n++) { odd[n] = row; row <<= 1; }
    gf2_square(even, odd);
    gf2_square(odd, even);
    do {
        gf2_square(even, odd);
        if (len2 & 1) crc1 = gf2_times(even, crc1);
        len2 >>= 1;
        if (len2 == 0) break;
        gf2_square(odd, even);
        if (len2 & 1) crc1 = gf2_times(odd, crc1);
        len2 >>= 1;
    } while (len2 != 0);
    return crc1 ^ crc2;
}

#define ADLER_BASE 65521u
uint32_t oracle_adler32(uint32_t adler, const uint8_t *p, size_t n)
{
    uint32_t a = adler & 0xFFFF, b = adler >> 16;
    while (n) {
        size_t k = n < 5552 ? n : 5552;
        n -= k;
        while (k--) { a += *p++; b += a; }
        a %= ADLER_BASE; b %= ADLER_BASE;
    }
    return (b << 16) | a;
}

uint32_t oracle_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
    uint32_t rem = (uint32_t)(len2 % ADLER_BASE);
    uint32_t sum1 = a1 & 0xFFFF;
    uint32_t sum2 = (rem * sum1) % ADLER_BASE;
    sum1 += (a2 & 0xFFFF) + ADLER_BASE - 1;
    sum2 += (a1 >> 16) + (a2 >> 16) + ADLER_BASE - rem;
    if (sum1 >= ADLER_BASE) sum1 -= ADLER_BASE;
    if (sum1 >= ADLER_BASE) sum1 -= ADLER_BASE;
    if (sum2 >= (ADLER_BASE << 1)) sum2 -= (ADLER_BASE << 1);
    if (sum2 >= ADLER_BASE) sum2 -= ADLER_BASE;
    return sum1 | (sum2 << 16);
}
