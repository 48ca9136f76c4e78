/*
 * oracle/oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * C interface of the CPU oracle (see README.md in this directory).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 */
#ifndef GZP_ORACLE_H
#define GZP_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_TRACE_MAX_BLOCKS 64

typedef struct {
    uint32_t begin, length, ntokens, tok_offset;
    int32_t btype; /* 0 stored, 1 static, 2 dynamic */
    uint32_t cost_dyn, cost_static, cost_stored;
    uint8_t litlen_lens[288];
    uint8_t offset_lens[32];
} oracle_block_info_t;

typedef struct {
    uint32_t nblocks;
    oracle_block_info_t blocks[ORACLE_TRACE_MAX_BLOCKS];
    uint32_t *tokens; /* optional, caller-allocated */
    size_t tokens_cap, ntokens;
} oracle_trace_t;

/* formats, numbered like include/gzpb.h */
enum { ORACLE_FMT_GZIP = 0, ORACLE_FMT_ZLIB = 1, ORACLE_FMT_RAWDEFLATE = 2, ORACLE_FMT_MGZIP = 3,
       ORACLE_FMT_BGZF = 4, ORACLE_FMT_SNAP = 5 };

/* deflate_oracle.c */
int oracle_level_supported(int level);
void oracle_thread_cleanup(void);
size_t oracle_deflate(const uint8_t *in, size_t n, int level, uint8_t *out, size_t out_cap);
void oracle_search_stats(uint64_t *searches, uint64_t *hops, int reset);
size_t oracle_deflate_ex(const uint8_t *in, size_t dict_len, size_t n, int level, int flush,
                         uint8_t *out, size_t out_cap, oracle_trace_t *trace);
void oracle_make_huffman_code(unsigned num_syms, unsigned max_len, const uint32_t *freqs,
                              uint8_t *lens, uint32_t *codewords);

/* check_oracle.c */
uint32_t oracle_crc32(uint32_t crc, const uint8_t *p, size_t n);
uint32_t oracle_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);
uint32_t oracle_crc32c(uint32_t crc, const uint8_t *p, size_t n);
uint32_t oracle_crc32c_masked(const uint8_t *p, size_t n);
uint32_t oracle_adler32(uint32_t adler, const uint8_t *p, size_t n);
uint32_t oracle_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2);

/* snappy_oracle.c */
size_t oracle_snappy_max_compress_len(size_t n);
size_t oracle_snappy_raw(const uint8_t *in, size_t n, uint8_t *out);
size_t oracle_snappy_frame(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap);

/* format_oracle.c — FormatSpec::encode / header / footer per format.
 * Returns the encoded size, or a negative GZPB_E* code (same numbering as
 * include/gzpb.h). */
size_t oracle_encode_capacity(int format, size_t n);
long oracle_encode_block(int format, int level, const uint8_t *in, size_t n, const uint8_t *dict,
                         size_t dict_len, int is_last, uint8_t *out, size_t out_cap);
size_t oracle_header(int format, int level, uint8_t *out);
size_t oracle_footer(int format, uint32_t sum, uint32_t amount, uint8_t *out);

/* par_oracle.c — ParCompress thread topology on the CPU (the reference's CPU
 * path, used by bench.py --impl reference and the cpu_baseline leg). */
double oracle_par_compress(int format, int level, size_t buffer_size, int num_threads, const uint8_t *in,
                           size_t n, uint8_t *out, size_t out_cap, size_t *out_len);

/* inflate_oracle.c — the block DECODE path (ParDecompress worker + reader loop). */
long oracle_inflate(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap);
long oracle_block_size(int format, const uint8_t *hdr, size_t avail);
int oracle_decode_stream(int format, const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *out_len,
                         uint32_t *found, uint32_t *expected);

#ifdef __cplusplus
}
#endif
#endif
