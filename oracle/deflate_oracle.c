/*
 * oracle/deflate_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the DEFLATE encoder that gzp's block formats call through
 * `libdeflater::Compressor::deflate_compress` (reference call sites:
 * /root/reference/src/bgzf.rs:214-216, /root/reference/src/mgzip.rs:201-203).
 * The arithmetic lives in libdeflate v1.24 (libdeflate-sys 1.24.0,
 * /root/reference/Cargo.lock:414-430), which is NOT vendored under
 * /root/reference and is not on this machine.  This file restates libdeflate's
 * published algorithm (hash-chain match finder, greedy / lazy / lazy2 parsers,
 * block splitting, length-limited canonical Huffman, cheapest-of
 * dynamic/static/stored block emission) from its documented behaviour.
 *
 * PARITY UNPINNED against real libdeflate bytes: the reference's own tests only
 * pin round trips, so this oracle is anchored on (i) stock
 * decoders (zlib / gzip), (ii) the container golden vectors in the
 * reference (BGZF_EOF, header recipes) and (iii) the libdeflate outputs that
 * follow exactly from its documented rules (empty input, pass-through of
 * inputs <= 55 - 4*level bytes, level 0: tests/golden/libdeflate_exact_vectors.json).
 * DESIGN.md §2 lists what is pinned and what is not.  The CUDA path is held
 * bit-exact to THIS file.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this code.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* ---- DEFLATE constants (RFC 1951) ------------------------------------- */
#define MIN_MATCH 3
#define MAX_MATCH 258
#define NUM_LITLEN 288
#define NUM_OFFSET 32
#define NUM_PRECODE 19
#define END_OF_BLOCK 256
#define FIRST_LEN_SYM 257
#define MAX_LITLEN_CW 14 /* libdeflate limits litlen codewords to 14 bits */
#define MAX_OFFSET_CW 15
#define MAX_PRE_CW 7
#define WINDOW 32768

#define SOFT_MAX_BLOCK_LENGTH 300000
#define MIN_BLOCK_LENGTH 5000
#define SEQ_STORE_LENGTH 50000
#define NUM_OBS_TYPES 10
#define OBS_PER_CHECK 512

static const uint16_t len_base[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t len_extra[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t off_base[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t off_extra[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const uint8_t pre_extra[19] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,3,7};
static const uint8_t pre_perm[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};

static uint8_t len_slot_tab[MAX_MATCH + 1];
static uint8_t off_slot_tab[WINDOW + 1];
static int tabs_ready;

static void init_tabs(void)
{
    if (tabs_ready) return;
    for (int s = 0; s < 29; s++) {
        int hi = (s == 28) ? 258 : len_base[s] + (1 << len_extra[s]) - 1;
        if (s == 27) hi = 257;
        for (int l = len_base[s]; l <= hi; l++) len_slot_tab[l] = (uint8_t)s;
    }
    for (int s = 0; s < 30; s++) {
        int hi = off_base[s] + (1 << off_extra[s]) - 1;
        for (int o = off_base[s]; o <= hi && o <= WINDOW; o++) off_slot_tab[o] = (uint8_t)s;
    }
    tabs_ready = 1;
}

/* ---- Huffman code construction --------------------------------------- */
/* Restates libdeflate's deflate_make_huffman_code(): symbols sorted ascending
 * by (freq, symbol); in-place Moffat-Katajainen tree; length limiting by
 * adjusting the per-length counts; canonical codewords, bit-reversed. */
#define SYM_BITS 10
#define SYM_MASK ((1u << SYM_BITS) - 1)
#define FREQ_MASK (~SYM_MASK)

static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

static uint32_t bitrev(uint32_t cw, unsigned len)
{
    uint32_t r = 0;
    for (unsigned i = 0; i < len; i++) r |= ((cw >> i) & 1u) << (len - 1 - i);
    return r;
}

void oracle_make_huffman_code(unsigned num_syms, unsigned max_len, const uint32_t *freqs,
                              uint8_t *lens, uint32_t *codewords)
{
    uint32_t A[NUM_LITLEN];
    unsigned n = 0;
    for (unsigned s = 0; s < num_syms; s++) {
        lens[s] = 0;
        if (freqs[s]) A[n++] = s | (freqs[s] << SYM_BITS);
    }
    qsort(A, n, sizeof(A[0]), cmp_u32); /* total order (freq, sym) */

    if (n < 2) {
        /* fewer than two used symbols: two 1-bit codewords, symbol 0 and the
         * used symbol (or symbol 1) */
        unsigned sym = n ? (A[0] & SYM_MASK) : 0;
        unsigned nz = sym ? sym : 1;
        for (unsigned s = 0; s < num_syms; s++) codewords[s] = 0;
        codewords[0] = 0; lens[0] = 1;
        codewords[nz] = 1; lens[nz] = 1;
        return;
    }

    /* build tree (parents stored in the high bits) */
    {
        const unsigned last = n - 1;
        unsigned i = 0, b = 0, e = 0;
        do {
            uint32_t nf;
            if (i + 1 <= last && (b == e || (A[i + 1] & FREQ_MASK) <= (A[b] & FREQ_MASK))) {
                nf = (A[i] & FREQ_MASK) + (A[i + 1] & FREQ_MASK);
                i += 2;
            } else if (b + 2 <= e && (i > last || (A[b + 1] & FREQ_MASK) < (A[i] & FREQ_MASK))) {
                nf = (A[b] & FREQ_MASK) + (A[b + 1] & FREQ_MASK);
                A[b] = (e << SYM_BITS) | (A[b] & SYM_MASK);
                A[b + 1] = (e << SYM_BITS) | (A[b + 1] & SYM_MASK);
                b += 2;
            } else {
                nf = (A[i] & FREQ_MASK) + (A[b] & FREQ_MASK);
                A[b] = (e << SYM_BITS) | (A[b] & SYM_MASK);
                i++; b++;
            }
            A[e] = nf | (A[e] & SYM_MASK);
        } while (++e < last);
    }

    unsigned len_counts[16];
    {
        int root = (int)n - 2;
        for (unsigned l = 0; l <= max_len; l++) len_counts[l] = 0;
        len_counts[1] = 2;
        A[root] &= SYM_MASK;
        for (int node = root - 1; node >= 0; node--) {
            unsigned parent = A[node] >> SYM_BITS;
            unsigned depth = (A[parent] >> SYM_BITS) + 1;
            A[node] = (A[node] & SYM_MASK) | (depth << SYM_BITS);
            if (depth >= max_len) {
                depth = max_len;
                do { depth--; } while (len_counts[depth] == 0);
            }
            len_counts[depth]--;
            len_counts[depth + 1] += 2;
        }
    }

    {
        unsigned i = 0;
        for (unsigned len = max_len; len >= 1; len--) {
            unsigned c = len_counts[len];
            while (c--) lens[A[i++] & SYM_MASK] = (uint8_t)len;
        }
        uint32_t next_cw[17];
        next_cw[0] = 0; next_cw[1] = 0;
        for (unsigned len = 2; len <= max_len; len++)
            next_cw[len] = (next_cw[len - 1] + len_counts[len - 1]) << 1;
        for (unsigned s = 0; s < num_syms; s++)
            codewords[s] = lens[s] ? bitrev(next_cw[lens[s]]++, lens[s]) : 0;
    }
}

/* ---- bit writer -------------------------------------------------------- */
typedef struct {
    uint8_t *out; size_t cap; size_t pos;
    uint64_t bitbuf; unsigned bitcount; int overflow;
} bitw_t;

/* bitcount stays < 32 between calls; whole 32-bit words are spilled eagerly */
static inline void bw_add(bitw_t *w, uint32_t bits, unsigned n)
{
    w->bitbuf |= (uint64_t)bits << w->bitcount;
    w->bitcount += n;
    if (w->bitcount >= 32) {
        if (w->pos + 4 <= w->cap) { uint32_t v = (uint32_t)w->bitbuf; memcpy(w->out + w->pos, &v, 4); }
        else w->overflow = 1;
        w->pos += 4; w->bitbuf >>= 32; w->bitcount -= 32;
    }
}

static inline void bw_align(bitw_t *w) { bw_add(w, 0, (0u - w->bitcount) & 7); }

/* write out the remaining whole/partial bytes */
static void bw_flush(bitw_t *w)
{
    while (w->bitcount > 0) {
        if (w->pos < w->cap) w->out[w->pos] = (uint8_t)w->bitbuf; else w->overflow = 1;
        w->pos++;
        w->bitbuf >>= 8; w->bitcount = w->bitcount >= 8 ? w->bitcount - 8 : 0;
    }
}

/* ---- compressor state -------------------------------------------------- */
typedef struct {
    int level;
    unsigned max_depth, nice;
    int mode; /* 0 greedy, 1 lazy, 2 lazy2 */
    /* match finder: absolute positions, -1 = empty */
    int32_t *head3, *head4, *next;
    /* current deflate block */
    uint32_t fl[NUM_LITLEN], fo[NUM_OFFSET];
    uint32_t *tokens; size_t ntok, nmatch; /* token: lit | (1u<<31)|len<<16|off */
    /* split stats */
    uint32_t obs[NUM_OBS_TYPES], new_obs[NUM_OBS_TYPES], num_obs, num_new_obs;
    /* static code */
    uint8_t sl_len[NUM_LITLEN]; uint32_t sl_cw[NUM_LITLEN];
    uint8_t so_len[NUM_OFFSET]; uint32_t so_cw[NUM_OFFSET];
    /* trace */
    oracle_trace_t *trace;
} comp_t;

static inline uint32_t ld32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t lz_hash(uint32_t seq, unsigned bits) { return (uint32_t)(seq * 0x1E35A7BDu) >> (32 - bits); }

static inline unsigned lz_extend(const uint8_t *a, const uint8_t *b, unsigned len, unsigned max_len)
{
    while (len + 8 <= max_len) {
        uint64_t x, y; memcpy(&x, a + len, 8); memcpy(&y, b + len, 8);
        if (x != y) return len + (unsigned)(__builtin_ctzll(x ^ y) >> 3);
        len += 8;
    }
    while (len < max_len && a[len] == b[len]) len++;
    return len;
}

static unsigned bsr32(uint32_t v) { unsigned r = 0; while (v >>= 1) r++; return r; }

/* hc_matchfinder_longest_match restated with absolute positions.  `nh` holds the
 * hashes of the position about to be searched (computed one call earlier; both
 * start at 0 for the first position, as in libdeflate). */
/* analysis aid (tests/search_stats.py): searches the parser actually asked for, and chain hops inside them */
static __thread uint64_t g_stat_searches, g_stat_hops;
void oracle_search_stats(uint64_t *searches, uint64_t *hops, int reset)
{
    if (searches) *searches = g_stat_searches;
    if (hops) *hops = g_stat_hops;
    if (reset) { g_stat_searches = 0; g_stat_hops = 0; }
}

static unsigned longest_match(comp_t *c, const uint8_t *in, size_t n, size_t p, unsigned best_len,
                              unsigned max_len, unsigned nice_len, unsigned depth, uint32_t nh[2],
                              unsigned *off_ret)
{
    size_t best_q = p;
    const unsigned depth0 = depth;
    (void)n;
    if (max_len < 5) goto out;
    g_stat_searches++;
    {
        uint32_t h3 = nh[0], h4 = nh[1];
        int32_t c3 = c->head3[h3], c4 = c->head4[h4];
        c->head3[h3] = (int32_t)p;
        c->head4[h4] = (int32_t)p;
        c->next[p] = c4;
        uint32_t nseq = ld32(in + p + 1);
        nh[0] = lz_hash(nseq & 0xFFFFFF, 15);
        nh[1] = lz_hash(nseq, 16);
#define INWIN(q) ((q) >= 0 && (int64_t)p - (q) < WINDOW)
        uint32_t seq4 = ld32(in + p);
        if (best_len < 4) {
            if (!INWIN(c3)) goto out;
            if (best_len < 3) {
                if ((ld32(in + c3) & 0xFFFFFF) == (seq4 & 0xFFFFFF)) { best_len = 3; best_q = (size_t)c3; }
            }
            if (!INWIN(c4)) goto out;
            for (;;) {
                if (ld32(in + c4) == seq4) break;
                c4 = c->next[c4];
                if (!INWIN(c4) || !--depth) goto out;
            }
            best_q = (size_t)c4;
            best_len = lz_extend(in + p, in + c4, 4, max_len);
            if (best_len >= nice_len) goto out;
            c4 = c->next[c4];
            if (!INWIN(c4) || !--depth) goto out;
        } else {
            if (!INWIN(c4) || best_len >= nice_len) goto out;
        }
        for (;;) {
            for (;;) {
                if (ld32(in + c4 + best_len - 3) == ld32(in + p + best_len - 3) && ld32(in + c4) == seq4) break;
                c4 = c->next[c4];
                if (!INWIN(c4) || !--depth) goto out;
            }
            unsigned l = lz_extend(in + p, in + c4, 4, max_len);
            if (l > best_len) {
                best_len = l; best_q = (size_t)c4;
                if (best_len >= nice_len) goto out;
            }
            c4 = c->next[c4];
            if (!INWIN(c4) || !--depth) goto out;
        }
    }
out:
    g_stat_hops += depth0 - depth;
    *off_ret = (unsigned)(p - best_q);
    return best_len;
}

static void skip_bytes(comp_t *c, const uint8_t *in, size_t n, size_t p, unsigned count, uint32_t nh[2])
{
    if ((size_t)count + 5 > n - p) return;
    uint32_t h3 = nh[0], h4 = nh[1];
    do {
        c->head3[h3] = (int32_t)p;
        c->next[p] = c->head4[h4];
        c->head4[h4] = (int32_t)p;
        p++;
        uint32_t nseq = ld32(in + p);
        h3 = lz_hash(nseq & 0xFFFFFF, 15);
        h4 = lz_hash(nseq, 16);
    } while (--count);
    nh[0] = h3; nh[1] = h4;
}

static unsigned choose_min_match_len(unsigned num_used, unsigned depth)
{
    static const uint8_t min_lens[80] = {
        9,9,9,9,9,9,8,8,7,7,6,6,6,6,6,6,
        5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
        5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
        5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,
        4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4};
    if (num_used >= 80) return 3;
    unsigned m = min_lens[num_used];
    if (depth < 16) {
        unsigned cap = depth < 5 ? 4 : depth < 10 ? 5 : 7;
        if (m > cap) m = cap;
    }
    return m;
}

static unsigned calc_min_match_len(const uint8_t *d, size_t len, unsigned depth)
{
    uint8_t used[256] = {0};
    unsigned nused = 0;
    if (len < 512) return MIN_MATCH;
    if (len > 4096) len = 4096;
    for (size_t i = 0; i < len; i++) used[d[i]] = 1;
    for (int i = 0; i < 256; i++) nused += used[i];
    return choose_min_match_len(nused, depth);
}

static unsigned recalc_min_match_len(const uint32_t *fl, unsigned depth)
{
    uint32_t total = 0, cutoff; unsigned nused = 0;
    for (int i = 0; i < 256; i++) total += fl[i];
    cutoff = total >> 10;
    for (int i = 0; i < 256; i++) if (fl[i] > cutoff) nused++;
    return choose_min_match_len(nused, depth);
}

static void begin_block(comp_t *c)
{
    memset(c->fl, 0, sizeof c->fl); memset(c->fo, 0, sizeof c->fo);
    c->ntok = 0; c->nmatch = 0;
    memset(c->obs, 0, sizeof c->obs); memset(c->new_obs, 0, sizeof c->new_obs);
    c->num_obs = c->num_new_obs = 0;
}

static void choose_literal(comp_t *c, uint8_t lit)
{
    c->fl[lit]++;
    c->new_obs[((lit >> 5) & 6) | (lit & 1)]++; c->num_new_obs++;
    c->tokens[c->ntok++] = lit;
}

static void choose_match(comp_t *c, unsigned len, unsigned off)
{
    c->fl[FIRST_LEN_SYM + len_slot_tab[len]]++;
    c->fo[off_slot_tab[off]]++;
    c->new_obs[8 + (len >= 9)]++; c->num_new_obs++;
    c->tokens[c->ntok++] = 0x80000000u | (len << 16) | off;
    c->nmatch++;
}

static int should_end_block(comp_t *c, size_t block_begin, size_t p, size_t n)
{
    if (!(c->num_new_obs >= OBS_PER_CHECK && p - block_begin >= MIN_BLOCK_LENGTH && n - p >= MIN_BLOCK_LENGTH))
        return 0;
    uint32_t block_length = (uint32_t)(p - block_begin);
    if (c->num_obs > 0) {
        uint32_t total_delta = 0;
        for (int i = 0; i < NUM_OBS_TYPES; i++) {
            uint32_t expected = c->obs[i] * c->num_new_obs;
            uint32_t actual = c->new_obs[i] * c->num_obs;
            total_delta += actual > expected ? actual - expected : expected - actual;
        }
        uint32_t num_items = c->num_obs + c->num_new_obs;
        uint32_t cutoff = c->num_new_obs * 200 / 512 * c->num_obs;
        if (block_length < 10000 && num_items < 8192)
            cutoff += (uint32_t)((uint64_t)cutoff * (8192 - num_items) / 8192);
        if (total_delta + (block_length / 4096) * c->num_obs >= cutoff) return 1;
    }
    for (int i = 0; i < NUM_OBS_TYPES; i++) { c->obs[i] += c->new_obs[i]; c->new_obs[i] = 0; }
    c->num_obs += c->num_new_obs; c->num_new_obs = 0;
    return 0;
}

/* RLE of the concatenated litlen+offset codeword lengths into precode items */
static unsigned compute_precode_items(const uint8_t *lens, unsigned num_lens, uint32_t *pfreq, unsigned *items)
{
    unsigned *ip = items, run_start = 0;
    memset(pfreq, 0, NUM_PRECODE * sizeof(uint32_t));
    do {
        uint8_t len = lens[run_start];
        unsigned run_end = run_start;
        do { run_end++; } while (run_end != num_lens && len == lens[run_end]);
        if (len == 0) {
            while (run_end - run_start >= 11) {
                unsigned eb = run_end - run_start - 11; if (eb > 0x7F) eb = 0x7F;
                pfreq[18]++; *ip++ = 18 | (eb << 5); run_start += 11 + eb;
            }
            if (run_end - run_start >= 3) {
                unsigned eb = run_end - run_start - 3; if (eb > 7) eb = 7;
                pfreq[17]++; *ip++ = 17 | (eb << 5); run_start += 3 + eb;
            }
        } else {
            if (run_end - run_start >= 4) {
                pfreq[len]++; *ip++ = len; run_start++;
                do {
                    unsigned eb = run_end - run_start - 3; if (eb > 3) eb = 3;
                    pfreq[16]++; *ip++ = 16 | (eb << 5); run_start += 3 + eb;
                } while (run_end - run_start >= 3);
            }
        }
        while (run_start != run_end) { pfreq[len]++; *ip++ = len; run_start++; }
    } while (run_start != num_lens);
    return (unsigned)(ip - items);
}

static void finish_block(comp_t *c, bitw_t *w, const uint8_t *in, size_t block_begin, size_t block_len, int is_final)
{
    uint8_t ll[NUM_LITLEN + NUM_OFFSET]; uint32_t lcw[NUM_LITLEN];
    uint8_t *ol = ll + NUM_LITLEN; uint32_t ocw[NUM_OFFSET];
    uint8_t olens_tmp[NUM_OFFSET];

    c->fl[END_OF_BLOCK]++;
    oracle_make_huffman_code(NUM_LITLEN, MAX_LITLEN_CW, c->fl, ll, lcw);
    oracle_make_huffman_code(NUM_OFFSET, MAX_OFFSET_CW, c->fo, olens_tmp, ocw);
    memcpy(ol, olens_tmp, NUM_OFFSET);

    unsigned nlit = NUM_LITLEN, noff = NUM_OFFSET;
    while (nlit > 257 && ll[nlit - 1] == 0) nlit--;
    while (noff > 1 && ol[noff - 1] == 0) noff--;
    uint8_t cat[NUM_LITLEN + NUM_OFFSET];
    memcpy(cat, ll, nlit); memcpy(cat + nlit, ol, noff);
    uint32_t pfreq[NUM_PRECODE]; unsigned items[NUM_LITLEN + NUM_OFFSET];
    unsigned nitems = compute_precode_items(cat, nlit + noff, pfreq, items);
    uint8_t plen[NUM_PRECODE]; uint32_t pcw[NUM_PRECODE];
    oracle_make_huffman_code(NUM_PRECODE, MAX_PRE_CW, pfreq, plen, pcw);
    unsigned nexpl = NUM_PRECODE;
    while (nexpl > 4 && plen[pre_perm[nexpl - 1]] == 0) nexpl--;

    uint32_t dyn = 3, stat = 3, unc = 3;
    dyn += 5 + 5 + 4 + 3 * nexpl;
    for (unsigned s = 0; s < NUM_PRECODE; s++) dyn += pfreq[s] * (pre_extra[s] + plen[s]);
    for (unsigned s = 0; s < 144; s++) { dyn += c->fl[s] * ll[s]; stat += c->fl[s] * 8; }
    for (unsigned s = 144; s < 256; s++) { dyn += c->fl[s] * ll[s]; stat += c->fl[s] * 9; }
    dyn += ll[END_OF_BLOCK]; stat += 7;
    for (unsigned s = 0; s < 29; s++) {
        dyn += c->fl[FIRST_LEN_SYM + s] * (len_extra[s] + ll[FIRST_LEN_SYM + s]);
        stat += c->fl[FIRST_LEN_SYM + s] * (len_extra[s] + c->sl_len[FIRST_LEN_SYM + s]);
    }
    for (unsigned s = 0; s < 30; s++) {
        dyn += c->fo[s] * (off_extra[s] + ol[s]);
        stat += c->fo[s] * (off_extra[s] + 5);
    }
    unc += ((0u - (w->bitcount + 3)) & 7) + 32 + 40 * (uint32_t)((block_len + 65534) / 65535 - 1) + 8 * (uint32_t)block_len;

    uint32_t best = dyn < stat ? dyn : stat; if (unc < best) best = unc;
    int btype = (best == unc) ? 0 : (best == stat) ? 1 : 2;

    if (c->trace && c->trace->nblocks < ORACLE_TRACE_MAX_BLOCKS) {
        oracle_block_info_t *bi = &c->trace->blocks[c->trace->nblocks++];
        bi->begin = (uint32_t)block_begin; bi->length = (uint32_t)block_len; bi->btype = btype;
        bi->ntokens = (uint32_t)c->ntok; bi->cost_dyn = dyn; bi->cost_static = stat; bi->cost_stored = unc;
        bi->tok_offset = (uint32_t)c->trace->ntokens;
        if (c->trace->tokens) {
            for (size_t i = 0; i < c->ntok && c->trace->ntokens < c->trace->tokens_cap; i++)
                c->trace->tokens[c->trace->ntokens++] = c->tokens[i];
        }
        memcpy(bi->litlen_lens, ll, NUM_LITLEN); memcpy(bi->offset_lens, ol, NUM_OFFSET);
    }

    if (btype == 0) {
        size_t pos = block_begin, end = block_begin + block_len;
        do {
            int bfinal = 0; size_t len = 65535;
            if (end - pos <= 65535) { bfinal = is_final; len = end - pos; }
            bw_add(w, (uint32_t)bfinal, 1); bw_add(w, 0, 2);
            bw_align(w);
            bw_add(w, (uint32_t)len, 16); bw_add(w, (uint32_t)(~len & 0xFFFF), 16);
            for (size_t i = 0; i < len; i++) bw_add(w, in[pos + i], 8);
            pos += len;
        } while (pos != end);
        return;
    }
    const uint8_t *L, *O; const uint32_t *LC, *OC;
    bw_add(w, (uint32_t)is_final, 1);
    if (btype == 1) {
        bw_add(w, 1, 2);
        L = c->sl_len; LC = c->sl_cw; O = c->so_len; OC = c->so_cw;
    } else {
        bw_add(w, 2, 2);
        bw_add(w, nlit - 257, 5); bw_add(w, noff - 1, 5); bw_add(w, nexpl - 4, 4);
        for (unsigned i = 0; i < nexpl; i++) bw_add(w, plen[pre_perm[i]], 3);
        for (unsigned i = 0; i < nitems; i++) {
            unsigned sym = items[i] & 0x1F;
            bw_add(w, pcw[sym], plen[sym]);
            bw_add(w, items[i] >> 5, pre_extra[sym]);
        }
        L = ll; LC = lcw; O = ol; OC = ocw;
    }
    for (size_t i = 0; i < c->ntok; i++) {
        uint32_t t = c->tokens[i];
        if (!(t & 0x80000000u)) { bw_add(w, LC[t], L[t]); continue; }
        unsigned len = (t >> 16) & 0x1FF, off = t & 0xFFFF;
        unsigned ls = len_slot_tab[len], os = off_slot_tab[off];
        bw_add(w, LC[FIRST_LEN_SYM + ls], L[FIRST_LEN_SYM + ls]);
        bw_add(w, len - len_base[ls], len_extra[ls]);
        bw_add(w, OC[os], O[os]);
        bw_add(w, off - off_base[os], off_extra[os]);
    }
    bw_add(w, LC[END_OF_BLOCK], L[END_OF_BLOCK]);
}

static void init_static(comp_t *c)
{
    uint32_t f[NUM_LITLEN];
    /* libdeflate builds the static codes from synthetic frequencies; the result is
     * the RFC 1951 fixed code. */
    for (int i = 0; i < 144; i++) f[i] = 1 << (9 - 8);
    for (int i = 144; i < 256; i++) f[i] = 1 << (9 - 9);
    for (int i = 256; i < 280; i++) f[i] = 1 << (9 - 7);
    for (int i = 280; i < 288; i++) f[i] = 1 << (9 - 8);
    oracle_make_huffman_code(NUM_LITLEN, 15, f, c->sl_len, c->sl_cw);
    for (int i = 0; i < NUM_OFFSET; i++) f[i] = 1;
    oracle_make_huffman_code(NUM_OFFSET, 15, f, c->so_len, c->so_cw);
}

/* The parsers.  `start` > 0 means in[0..start) is a preset dictionary: it is
 * inserted into the match finder and never emitted (used by the Gzip/RawDeflate
 * formats; not a libdeflate feature — see DESIGN.md). */
static void compress_hc(comp_t *c, const uint8_t *in, size_t start, size_t n, bitw_t *w, int final_block)
{
    uint32_t nh[2] = {0, 0};
    size_t p = start;
    unsigned max_len = MAX_MATCH, nice_len = c->nice < MAX_MATCH ? c->nice : MAX_MATCH;

    for (size_t i = 0; i < 32768; i++) c->head3[i] = -1;
    for (size_t i = 0; i < 65536; i++) c->head4[i] = -1;
    if (start) {
        /* dictionary priming: every dictionary position is inserted */
        nh[0] = lz_hash(ld32(in) & 0xFFFFFF, 15); nh[1] = lz_hash(ld32(in), 16);
        skip_bytes(c, in, n, 0, (unsigned)start, nh);
        if ((size_t)start + 5 > n) { nh[0] = nh[1] = 0; }
    }

#define ADJUST(rem) do { if ((rem) < MAX_MATCH) { max_len = (unsigned)(rem); if (nice_len > max_len) nice_len = max_len; } } while (0)
    do {
        const size_t block_begin = p;
        const size_t max_block_end = (n - p < SOFT_MAX_BLOCK_LENGTH + MIN_BLOCK_LENGTH) ? n : p + SOFT_MAX_BLOCK_LENGTH;
        size_t next_recalc = p + (n - p < 10000 ? n - p : 10000);
        unsigned min_len;
        begin_block(c);
        min_len = calc_min_match_len(in + p, max_block_end - p, c->max_depth);
        do {
            unsigned cur_len, cur_off, next_len, next_off;
            if (c->mode == 0) {
                ADJUST(n - p);
                cur_len = longest_match(c, in, n, p, min_len - 1, max_len, nice_len, c->max_depth, nh, &cur_off);
                if (cur_len >= min_len && (cur_len > MIN_MATCH || cur_off <= 4096)) {
                    choose_match(c, cur_len, cur_off);
                    skip_bytes(c, in, n, p + 1, cur_len - 1, nh);
                    p += cur_len;
                } else {
                    choose_literal(c, in[p++]);
                }
                continue;
            }
            if (p >= next_recalc) {
                min_len = recalc_min_match_len(c->fl, c->max_depth);
                size_t a = n - next_recalc, b = p - block_begin;
                next_recalc += a < b ? a : b;
            }
            ADJUST(n - p);
            cur_len = longest_match(c, in, n, p, min_len - 1, max_len, nice_len, c->max_depth, nh, &cur_off);
            if (cur_len < min_len || (cur_len == MIN_MATCH && cur_off > 8192)) {
                choose_literal(c, in[p++]);
                continue;
            }
            p++;
        have_cur_match:
            if (cur_len >= nice_len) {
                choose_match(c, cur_len, cur_off);
                skip_bytes(c, in, n, p, cur_len - 1, nh);
                p += cur_len - 1;
                continue;
            }
            ADJUST(n - p);
            next_len = longest_match(c, in, n, p, cur_len - 1, max_len, nice_len, c->max_depth >> 1, nh, &next_off);
            p++;
            if (next_len >= cur_len && 4 * (int)(next_len - cur_len) + ((int)bsr32(cur_off) - (int)bsr32(next_off)) > 2) {
                choose_literal(c, in[p - 2]);
                cur_len = next_len; cur_off = next_off;
                goto have_cur_match;
            }
            if (c->mode == 2) {
                ADJUST(n - p);
                next_len = longest_match(c, in, n, p, cur_len - 1, max_len, nice_len, c->max_depth >> 2, nh, &next_off);
                p++;
                if (next_len >= cur_len && 4 * (int)(next_len - cur_len) + ((int)bsr32(cur_off) - (int)bsr32(next_off)) > 6) {
                    choose_literal(c, in[p - 3]);
                    choose_literal(c, in[p - 2]);
                    cur_len = next_len; cur_off = next_off;
                    goto have_cur_match;
                }
                choose_match(c, cur_len, cur_off);
                if (cur_len > 3) {
                    skip_bytes(c, in, n, p, cur_len - 3, nh);
                    p += cur_len - 3;
                }
            } else {
                choose_match(c, cur_len, cur_off);
                skip_bytes(c, in, n, p, cur_len - 2, nh);
                p += cur_len - 2;
            }
        } while (p < max_block_end && c->nmatch < SEQ_STORE_LENGTH && !should_end_block(c, block_begin, p, n));
        finish_block(c, w, in, block_begin, p - block_begin, final_block && p == n);
    } while (p != n && !w->overflow);
}


/* ---- level 1: deflate_compress_fastest() + ht_matchfinder restated ---------------------------
 * Hash table of 2-entry buckets keyed by a 15-bit hash of 4 bytes (HT_MATCHFINDER_HASH_ORDER 15,
 * BUCKET_SIZE 2, MIN_MATCH_LEN 4, REQUIRED_NBYTES 5); every position is inserted (searched ones by
 * longest_match, skipped ones by skip_bytes), position 0 under hash 0 because next_hash starts at 0.
 * Greedy parse, nice_match_length 32, no min_len heuristics, no split statistics; a block ends at
 * FAST_SOFT_MAX_BLOCK_LENGTH (65535) bytes or FAST_SEQ_STORE_LENGTH (8192) matches. */
#define FAST_SOFT_MAX_BLOCK_LENGTH 65535
#define FAST_SEQ_STORE_LENGTH 8192
#define HT_ORDER 15

static unsigned ht_longest_match(int32_t (*tab)[2], const uint8_t *in, size_t p, unsigned max_len, unsigned nice_len,
                                 uint32_t *next_hash, unsigned *off_ret)
{
    unsigned best_len = 0;
    size_t best_q = p;
    uint32_t hash = *next_hash;
    *next_hash = lz_hash(ld32(in + p + 1), HT_ORDER);
    uint32_t seq = ld32(in + p);
    int32_t cur = tab[hash][0];
    tab[hash][0] = (int32_t)p;
    if (!INWIN(cur)) goto out;
    {
        int32_t to_insert = cur;
        int32_t second = tab[hash][1];
        tab[hash][1] = to_insert;
        if (ld32(in + cur) == seq) {
            best_len = lz_extend(in + p, in + cur, 4, max_len);
            best_q = (size_t)cur;
            if (!INWIN(second) || best_len >= nice_len) goto out;
            if (ld32(in + second) == seq && ld32(in + second + best_len - 3) == ld32(in + p + best_len - 3)) {
                unsigned len = lz_extend(in + p, in + second, 4, max_len);
                if (len > best_len) { best_len = len; best_q = (size_t)second; }
            }
        } else {
            if (!INWIN(second)) goto out;
            if (ld32(in + second) == seq) { best_len = lz_extend(in + p, in + second, 4, max_len); best_q = (size_t)second; }
        }
    }
out:
    *off_ret = (unsigned)(p - best_q);
    return best_len;
}

static void ht_skip_bytes(int32_t (*tab)[2], const uint8_t *in, size_t n, size_t p, unsigned count, uint32_t *next_hash)
{
    if ((size_t)count + 5 > n - p) return;
    uint32_t hash = *next_hash;
    do {
        tab[hash][1] = tab[hash][0];
        tab[hash][0] = (int32_t)p;
        p++;
        hash = lz_hash(ld32(in + p), HT_ORDER);
    } while (--count);
    *next_hash = hash;
}

static void compress_fastest(comp_t *c, const uint8_t *in, size_t start, size_t n, bitw_t *w, int final_block)
{
    int32_t (*tab)[2] = (int32_t (*)[2])c->head4;     /* 32768 buckets x 2 = the 65536-entry scratch */
    uint32_t next_hash = 0;
    size_t p = start;
    unsigned max_len = MAX_MATCH, nice_len = c->nice < MAX_MATCH ? c->nice : MAX_MATCH;
    for (size_t i = 0; i < 32768; i++) tab[i][0] = tab[i][1] = -1;
    if (start) {
        next_hash = lz_hash(ld32(in), HT_ORDER);
        ht_skip_bytes(tab, in, n, 0, (unsigned)start, &next_hash);
        if ((size_t)start + 5 > n) next_hash = 0;
    }
    do {
        const size_t block_begin = p;
        const size_t max_block_end = (n - p < FAST_SOFT_MAX_BLOCK_LENGTH + MIN_BLOCK_LENGTH) ? n : p + FAST_SOFT_MAX_BLOCK_LENGTH;
        begin_block(c);
        do {
            size_t remaining = n - p;
            unsigned length, offset;
            if (remaining < MAX_MATCH) {
                max_len = (unsigned)remaining;
                if (max_len < 5) {
                    do { choose_literal(c, in[p++]); } while (--max_len);
                    break;
                }
                if (nice_len > max_len) nice_len = max_len;
            }
            length = ht_longest_match(tab, in, p, max_len, nice_len, &next_hash, &offset);
            if (length) {
                choose_match(c, length, offset);
                ht_skip_bytes(tab, in, n, p + 1, length - 1, &next_hash);
                p += length;
            } else {
                choose_literal(c, in[p++]);
            }
        } while (p < max_block_end && c->nmatch < FAST_SEQ_STORE_LENGTH);
        finish_block(c, w, in, block_begin, p - block_begin, final_block && p == n);
    } while (p != n && !w->overflow);
}

static void compress_none(const uint8_t *in, size_t n, bitw_t *w, int final_block)
{
    size_t pos = 0;
    if (n == 0) { bw_add(w, (uint32_t)final_block, 8); bw_add(w, 0, 16); bw_add(w, 0xFFFF, 16); return; }
    do {
        int bfinal = 0; size_t len = 65535;
        if (n - pos <= 65535) { bfinal = final_block; len = n - pos; }
        bw_add(w, (uint32_t)bfinal, 8);
        bw_add(w, (uint32_t)len, 16); bw_add(w, (uint32_t)(~len & 0xFFFF), 16);
        for (size_t i = 0; i < len; i++) bw_add(w, in[pos + i], 8);
        pos += len;
    } while (pos != n);
}

static int level_params(int level, unsigned *depth, unsigned *nice, int *mode)
{
    switch (level) {
    case 1: *mode = 3; *depth = 2; *nice = 32; return 0;    /* deflate_compress_fastest (ht_matchfinder) */
    case 2: *mode = 0; *depth = 6; *nice = 10; return 0;
    case 3: *mode = 0; *depth = 12; *nice = 14; return 0;
    case 4: *mode = 0; *depth = 16; *nice = 30; return 0;
    case 5: *mode = 1; *depth = 16; *nice = 30; return 0;
    case 6: *mode = 1; *depth = 35; *nice = 65; return 0;
    case 7: *mode = 1; *depth = 100; *nice = 130; return 0;
    case 8: *mode = 2; *depth = 300; *nice = 258; return 0;
    case 9: *mode = 2; *depth = 600; *nice = 258; return 0;
    default: return -1;
    }
}

int oracle_level_supported(int level) { unsigned d, nn; int m; return level == 0 || level_params(level, &d, &nn, &m) == 0; }

/* per-thread scratch (see oracle_deflate_ex); released by oracle_thread_cleanup() */
static __thread int32_t *t_head3, *t_head4, *t_next;
static __thread uint32_t *t_tokens;
static __thread size_t t_cap;

void oracle_thread_cleanup(void)
{
    free(t_head3); free(t_head4); free(t_next); free(t_tokens);
    t_head3 = t_head4 = t_next = NULL; t_tokens = NULL; t_cap = 0;
}

/* General entry: in[0..dict_len) is a preset dictionary, in[dict_len..dict_len+n)
 * the data.  flush: 0 = finish (BFINAL on the last block), 1 = sync flush (no
 * BFINAL; append an empty stored block so the segment ends byte-aligned). */
size_t oracle_deflate_ex(const uint8_t *in, size_t dict_len, size_t n, int level, int flush,
                         uint8_t *out, size_t out_cap, oracle_trace_t *trace)
{
    bitw_t w = {out, out_cap, 0, 0, 0, 0};
    comp_t c; memset(&c, 0, sizeof c);
    init_tabs();
    c.level = level; c.trace = trace;
    if (trace) { trace->nblocks = 0; trace->ntokens = 0; }
    int final_block = (flush == 0);
    size_t passthrough = (level == 0) ? (size_t)-1 : (size_t)(55 - level * 4);
    if (level != 0 && level_params(level, &c.max_depth, &c.nice, &c.mode) != 0) return 0;
    if (n <= passthrough && !(flush == 1 && n == 0)) {
        compress_none(in + dict_len, n, &w, final_block);
    } else if (n > 0) {
        size_t tot = dict_len + n;
        /* per-thread scratch, grown on demand and reused across calls (like a
         * libdeflate_compressor that lives as long as its worker thread) */
        if (!t_head3) { t_head3 = malloc(32768 * sizeof(int32_t)); t_head4 = malloc(65536 * sizeof(int32_t)); }
        if (t_cap < tot + 1) {
            free(t_next); free(t_tokens);
            t_cap = tot + 1 + tot / 4;
            t_next = malloc(t_cap * sizeof(int32_t)); t_tokens = malloc(t_cap * sizeof(uint32_t));
        }
        c.head3 = t_head3; c.head4 = t_head4; c.next = t_next; c.tokens = t_tokens;
        init_static(&c);
        if (c.mode == 3) compress_fastest(&c, in, dict_len, tot, &w, final_block);
        else compress_hc(&c, in, dict_len, tot, &w, final_block);
    }
    if (flush == 1) {
        /* Z_SYNC_FLUSH marker: empty stored block, byte aligned: 00 00 ff ff */
        bw_add(&w, 0, 3);
        bw_align(&w);
        bw_add(&w, 0, 16); bw_add(&w, 0xFFFF, 16);
    }
    bw_flush(&w);
    if (w.overflow) return 0;
    return w.pos;
}

/* libdeflate_deflate_compress() equivalent (no dictionary, finish). */
size_t oracle_deflate(const uint8_t *in, size_t n, int level, uint8_t *out, size_t out_cap)
{
    return oracle_deflate_ex(in, 0, n, level, 0, out, out_cap, NULL);
}
