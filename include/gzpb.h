/*
 * gzpb.h — C ABI of the B200-native per-block encode path for gzp.
 *
 * This is the drop-in boundary: the entry points a Rust `FormatSpec` impl (or the
 * worker loop of gzp's ParCompress) would bind over FFI instead of calling
 * libdeflate / zlib-ng / snap on a CPU thread.  Plain C types only; no CUDA or
 * torch types cross the boundary (streams and device pointers travel as void*).
 * Every entry point cites the reference interface it replaces
 * (paths relative to the reference repo sstadick/gzp v2.0.1).
 *
 * Threading: a gzpb_ctx is owned by ONE host thread at a time (Send, not Sync),
 * like the per-worker `Compressor` of src/par/compress.rs:278.
 * Errors: negative int codes, 1:1 with GzpError variants (src/lib.rs:114-163);
 * nothing unwinds across the ABI; no partial output on error.
 */
#ifndef GZPB_H
#define GZPB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Format types (src/deflate.rs:61,169,279,357,506; src/snap.rs:35). */
enum gzpb_format {
    GZPB_GZIP = 0,
    GZPB_ZLIB = 1,
    GZPB_RAWDEFLATE = 2,
    GZPB_MGZIP = 3,
    GZPB_BGZF = 4,
    GZPB_SNAP = 5
};

/* Error codes <-> GzpError (src/lib.rs:114-163). */
enum gzpb_status {
    GZPB_OK = 0,
    GZPB_EBUFFERSIZE = -1,  /* BufferSize(got, min): buffer_size < 32 KiB (src/par/compress.rs:68-74) */
    GZPB_ENUMTHREADS = -2,  /* NumThreads(0) (src/par/compress.rs:84-90) */
    GZPB_EBLOCKSIZE = -3,   /* BlockSizeExceeded(n, 65536) (src/bgzf.rs:218-223) */
    GZPB_ECOMPRESS = -4,    /* LibDeflaterCompress: output did not fit the capacity (src/bgzf.rs:214-216) */
    GZPB_ELEVEL = -5,       /* LibDeflaterCompressionLvl: level not supported by this engine */
    GZPB_EIO = -6,          /* Io */
    GZPB_ECHANNEL = -7,     /* ChannelSend / ChannelReceive: stream already finished / torn down */
    GZPB_ECUDA = -8,        /* CUDA runtime / launch failure (no reference analogue) */
    GZPB_EINVAL = -9,       /* bad argument (Unknown) */
    GZPB_ENOMEM = -10,
    GZPB_EHEADER = -11,     /* InvalidHeader: "Extra field flag not set" / "Bad SID" (src/deflate.rs:407-417, 555-565) */
    GZPB_ECHECK = -12,      /* InvalidCheck{found, expected} (src/par/decompress.rs:176-181) */
    GZPB_EDECOMPRESS = -13, /* LibDelfaterDecompress / DecompressError: corrupt DEFLATE data (src/deflate.rs:393, 541) */
    GZPB_EBLOCK = -14,      /* InvalidBlockSize: a member shorter than its own header + footer */
    GZPB_EAGAIN = -15       /* not an error: every lane is in flight (gzpb_submit) / the ticket is not finished yet
                               (gzpb_poll without wait) — the bounded channel of src/par/compress.rs:111-112 is full */
};

/* Constants of the reference (src/lib.rs:105,108; src/bgzf.rs:20,22). */
#define GZPB_BUFSIZE 131072u
#define GZPB_DICT_SIZE 32768u
#define GZPB_BGZF_BLOCK_SIZE 65280u
#define GZPB_MAX_BGZF_BLOCK_SIZE 65536u
/* device layout of one unit (one gzp block) for gzpb_encode_device */
#define GZPB_IN_STRIDE 65600u
#define GZPB_MAX_UNIT_BYTES 65536u

typedef struct gzpb_ctx gzpb_ctx;

/* One message of the compressor channel: `Message{buffer, dictionary, is_last}`
 * (src/lib.rs:282-312). */
typedef struct {
    const void *ptr;
    size_t len;
    const void *dict; /* last 32 KiB of the previous block, or NULL (src/par/compress.rs:419-423) */
    size_t dict_len;
    int is_last;
} gzpb_block_in;

/* One completed oneshot: `(F::C, Vec<u8>)` (src/lib.rs:112). */
typedef struct {
    void *dst;        /* caller-owned, capacity >= gzpb_encode_capacity(format, len) */
    size_t cap;
    size_t out_len;
    uint32_t check_sum;    /* Check::sum of this block (CRC-32 for Gzip, Adler-32 for Zlib, 0 otherwise) */
    uint32_t check_amount; /* Check::amount */
    int status;
} gzpb_block_out;

/* FormatSpec::create_compressor (src/lib.rs:343-346): one context per GPU.
 * max_block_bytes = the ParCompress buffer_size; max_blocks_in_flight bounds the
 * device batch (the analogue of the 2*num_threads channel bound,
 * src/par/compress.rs:111-112).  Fails with GZPB_ECUDA when no sm_100 device /
 * library is usable — there is no CPU fallback. */
int gzpb_create(gzpb_ctx **ctx, int device, int format, int level, size_t max_block_bytes,
                size_t max_blocks_in_flight);
void gzpb_destroy(gzpb_ctx *ctx);

/* FormatSpec::encode + Check::update for n blocks at once
 * (src/par/compress.rs:281-289; src/deflate.rs:86-110,304-323,463-472,613-626;
 * src/snap.rs:61-74).  Host pointers in, host pointers out; synchronous. */
int gzpb_encode_batch(gzpb_ctx *ctx, size_t n, const gzpb_block_in *in, gzpb_block_out *out);

/* The same, asynchronous — the device analogue of `tx_compressor.send(msg)` + the oneshot the writer
 * thread later receives (src/par/compress.rs:283-294, 305-306).  gzpb_submit hands n <= max_blocks_in_flight
 * blocks to the next free lane and returns a ticket at once (GZPB_EAGAIN while all lanes are in flight);
 * gzpb_poll(ticket, wait) completes tickets in submission order up to and including `ticket`: GZPB_OK once
 * its `out[]` entries are filled, GZPB_EAGAIN when `wait` is 0 and the device is not done.  `in`, `out`
 * and the buffers they point to belong to the library until the ticket completes.  Blocks that lie in
 * pinned memory (gzpb_host_alloc) go to the device by DMA from where they are; pageable ones are staged. */
typedef uint64_t gzpb_ticket;
int gzpb_submit(gzpb_ctx *ctx, size_t n, const gzpb_block_in *in, gzpb_block_out *out, gzpb_ticket *ticket);
int gzpb_poll(gzpb_ctx *ctx, gzpb_ticket ticket, int wait);

/* ParCompress end to end for an in-memory input: header + write(in) + finish()
 * + footer, i.e. chunking with the reference's strict '>' hold-back and final
 * flush (src/par/compress.rs:413-463, 332-362), ordered output (:303-313).
 * `in`/`out` may be pageable or pinned (gzpb_host_alloc) host memory. */
int gzpb_encode_stream(gzpb_ctx *ctx, const void *in, size_t in_len, size_t buffer_size, void *out,
                       size_t out_cap, size_t *out_len);

/* The same stream over several GPUs of one box: device batches of max_blocks_in_flight consecutive
 * blocks are dealt round-robin to the contexts (one per GPU, all created with the same format, level and sizes);
 * blocks are independent and a batch's first dictionary comes from the host input, so no data moves between GPUs.
 * The ordered writer thread of src/par/compress.rs:303-313 is an offset chain: with pinned `in` / `out`
 * (gzpb_host_alloc) every batch's scan waits for the previous batch's scan — on whichever GPU it ran — and its
 * blocks land at their final stream position in `out` by DMA.  Byte-identical to gzpb_encode_stream on one GPU.
 * One calling thread drives all contexts. */
int gzpb_encode_stream_multi(gzpb_ctx *const *ctxs, size_t nctx, const void *in, size_t in_len, size_t buffer_size,
                             void *out, size_t out_cap, size_t *out_len);

/* Incremental writer = `ParCompress<F, W>` as a C object (src/par/compress.rs:221-233): the caller
 * `write`s bytes, the writer cuts blocks with the reference's semantics (strict '>' hold-back :415,
 * 32 KiB dictionary carry :419-423, `flush` = flush_last(false) :466-468 incl. the empty block,
 * `finish` = flush_last(true) + footer :377-388) and hands the encoded blocks to `sink` in stream
 * order (the ordered writer loop :303-313).  `sink` plays `W: Write`: nonzero return = io error
 * (GZPB_EIO is then returned by the next call, like the reference surfaces a BrokenPipe).
 * The caller's bytes are copied once into pinned slabs; up to 3 device batches of `blocks_in_flight`
 * blocks per GPU stay in flight while the caller keeps writing (back-pressure = the bounded channels
 * of :111-112); `sink` is called once per finished batch, in order, from the calling thread.
 * buffer_size 0 = the format's default; blocks_in_flight 0 = 1184 (8 units per SM on 148 SMs: a small memory
 * footprint).  Throughput peaks when a batch fills the GPU with one wave of k_emit units — 148 SMs x 32 = 4736
 * blocks (bench.py's BGZF batch; Mgzip 131072-byte blocks: 3256) — at about 2.1 MB of device scratch per 64 KiB
 * block and batch in flight. */
typedef int (*gzpb_sink_fn)(void *user, const void *data, size_t len);
typedef struct gzpb_writer gzpb_writer;
int gzpb_writer_create(gzpb_writer **w, int device, int format, int level, size_t buffer_size,
                       size_t blocks_in_flight, gzpb_sink_fn sink, void *user);
/* The same writer over several GPUs of one box: device batches of `blocks_in_flight` consecutive blocks
 * are dealt round-robin to the devices (SURVEY.md §8e — blocks are independent, a batch's first
 * dictionary comes from host memory, no device-to-device traffic); one ordered sink. */
int gzpb_writer_create_multi(gzpb_writer **w, const int *devices, size_t ndevices, int format, int level,
                             size_t buffer_size, size_t blocks_in_flight, gzpb_sink_fn sink, void *user);
int gzpb_writer_write(gzpb_writer *w, const void *buf, size_t len);
/* The one host copy of `write` (src/par/compress.rs:414) is what bounds a single caller: one core copies about as
 * fast as ONE B200 compresses.  With nthreads > 1, writes of 8 MiB and more are copied into the pinned slab by the
 * caller and nthreads - 1 helper threads side by side (default 1 = the reference's behaviour). */
int gzpb_writer_set_copy_threads(gzpb_writer *w, int nthreads);
/* Zero-copy form of write for callers that can produce their bytes in place (`Read::read(&mut buf)`, read(2)
 * from a file): gzpb_writer_reserve returns the writer's fill position inside its pinned slab and how many
 * contiguous bytes fit there (never 0); gzpb_writer_commit(n) declares the first n of them written and cuts
 * blocks exactly like write.  Removes the single-caller memcpy of src/par/compress.rs:414 from the path. */
int gzpb_writer_reserve(gzpb_writer *w, void **ptr, size_t *room);
int gzpb_writer_commit(gzpb_writer *w, size_t n);
int gzpb_writer_flush(gzpb_writer *w);
int gzpb_writer_finish(gzpb_writer *w);
void gzpb_writer_destroy(gzpb_writer *w);
/* counters of a writer: bytes written by the caller, bytes handed to the sink, device batches, sink calls */
int gzpb_writer_stats(gzpb_writer *w, uint64_t *bytes_in, uint64_t *bytes_out, uint64_t *batches, uint64_t *sink_calls);

/* File to file (the ingest / egress step either side of the path): read(2) fills the
 * writer's pinned slabs directly, the ordered output is written from pinned memory.  What
 * `ZBuilder::<F, _>::new().from_writer(File::create(out))` + `io::copy(&mut File::open(in), &mut z)` + `finish`
 * does in the reference (README.md:60-75), on `ndevices` GPUs. */
int gzpb_compress_file(const int *devices, size_t ndevices, int format, int level, size_t buffer_size,
                       size_t blocks_in_flight, const char *in_path, const char *out_path, uint64_t *bytes_in,
                       uint64_t *bytes_out);

/* Device-resident form of the same path, asynchronous on `cuda_stream`:
 * d_in holds nunits slots of GZPB_IN_STRIDE bytes, d_len/d_flags one u32 per unit
 * (flags bit0 = is_last, bit1 = sync flush); on completion d_packed holds the
 * encoded blocks in order, d_offsets[i] their byte offsets (d_offsets[nunits] =
 * total), d_status[i] the per-block status. */
int gzpb_encode_device(gzpb_ctx *ctx, const void *d_in, const uint32_t *d_len, const uint32_t *d_flags,
                       size_t nunits, void *d_packed, uint64_t *d_offsets, int32_t *d_status,
                       void *cuda_stream);

/* The same for every format: slots of gzpb_unit_stride(ctx) bytes, each [dictionary | data]; d_len[i] = dictionary +
 * data bytes, d_dict[i] = dictionary bytes (NULL when the format takes none).  Snap: d_offsets and the packed stream
 * have one entry per 64 KiB chunk, ceil(max_block_bytes / 65536) per unit; d_flags / d_status may be NULL. */
int gzpb_encode_device_ex(gzpb_ctx *ctx, const void *d_in, const uint32_t *d_len, const uint32_t *d_dict,
                          const uint32_t *d_flags, size_t nunits, void *d_packed, uint64_t *d_offsets,
                          int32_t *d_status, void *cuda_stream);
size_t gzpb_unit_stride(gzpb_ctx *ctx);

/* Capacity contract of the reference's output Vec (src/bgzf.rs:211-212,
 * src/mgzip.rs:194-195, src/deflate.rs:50-53). */
size_t gzpb_encode_capacity(int format, size_t len);
/* FormatSpec::header / footer (src/deflate.rs:113-142, 221-251, 325-331, 474-480, 628-634). */
size_t gzpb_header(int format, int level, void *buf16);
size_t gzpb_footer(int format, uint32_t check_sum, uint32_t check_amount, void *buf16);
/* Check::combine (src/check.rs:162, 121-128). */
uint32_t gzpb_crc32_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b);
uint32_t gzpb_adler32_combine(uint32_t adler_a, uint32_t adler_b, uint64_t len_b);
/* FormatSpec::DEFAULT_BUFSIZE / needs_dict (src/lib.rs:330, src/deflate.rs:79-82,583). */
size_t gzpb_default_bufsize(int format);
int gzpb_needs_dict(int format);
int gzpb_level_supported(int format, int level);

/* Pinned host memory for zero-copy staging (replaces the `Bytes` buffers the
 * reference moves through its channels, src/par/compress.rs:416). */
void *gzpb_host_alloc(size_t bytes);
void gzpb_host_free(void *p);

/* Per-kernel device time of the calls since the last reset, measured with CUDA
 * events on the launching stream (for bench.py's roofline block).
 * names: "chain","match","emit","gather","crc","snap" */
int gzpb_set_profiling(gzpb_ctx *ctx, int on);
int gzpb_kernel_ms(gzpb_ctx *ctx, const char *name, double *total_ms, uint64_t *launches);
uint64_t gzpb_launch_count(gzpb_ctx *ctx);
/* which kernel set this context runs: "split+link+match" (the DEFLATE family) or "snap" */
const char *gzpb_ctx_variant(gzpb_ctx *ctx);

/* ---- decoder: the ParDecompress path -------------------------------
 * `trait BlockFormatSpec` (src/lib.rs:411-448) is implemented by Mgzip and Bgzf
 * (src/deflate.rs:359-423, 508-571); ParDecompress::run (src/par/decompress.rs:132-220) reads
 * one member at a time (check_header, get_block_size), and a worker inflates it into ISIZE
 * bytes and verifies the CRC-32 of the footer. */
typedef struct gzpb_decoder gzpb_decoder;

/* One member to decode (also the device layout consumed by gzpb_decode_device): raw DEFLATE
 * payload at comp + in_off, `out_len` (ISIZE) bytes to produce at out + out_off, footer CRC-32. */
typedef struct {
    uint64_t in_off;
    uint64_t out_off;
    uint32_t in_len;
    uint32_t out_len;
    uint32_t crc;
    uint32_t pad;
} gzpb_block_desc;

/* BlockFormatSpec::HEADER_SIZE (src/deflate.rs:370, 519). */
size_t gzpb_block_header_size(int format);
/* BlockFormatSpec::check_header + get_block_size (src/deflate.rs:407-422, 555-570): total size of
 * the member starting at `hdr`, or a negative status code: GZPB_EHEADER, or GZPB_EIO when fewer than
 * HEADER_SIZE bytes are available. */
long gzpb_block_size(int format, const void *hdr, size_t avail);
/* The reader loop of ParDecompress::run (src/par/decompress.rs:190-207) over an in-memory input:
 * one descriptor per member (get_footer_values, src/lib.rs:440-447).  `descs` may be NULL to count.
 * A short trailing header is EOF; a truncated member returns GZPB_EIO with `consumed` at its start. */
int gzpb_scan_blocks(int format, const void *in, size_t in_len, gzpb_block_desc *descs, size_t max_descs,
                     size_t *nblocks, size_t *consumed, uint64_t *total_out);
/* BlockFormatSpec::create_decompressor (src/deflate.rs:372-381, 521-530): one decoder per GPU. */
int gzpb_decoder_create(gzpb_decoder **d, int device, int format, size_t max_blocks_in_flight);
void gzpb_decoder_destroy(gzpb_decoder *d);
/* ParDecompress end to end for an in-memory input of concatenated members: decoded bytes in stream
 * order.  With `consumed` non-NULL an incomplete trailing member is left unconsumed (incremental
 * readers call again with more bytes); with NULL it is GZPB_EIO like the reference's read_exact. */
int gzpb_decode_stream(gzpb_decoder *d, const void *in, size_t in_len, void *out, size_t out_cap, size_t *out_len,
                       size_t *consumed);
/* Device-resident form, asynchronous on `cuda_stream`: d_status[i] = 0 ok, 1 bad data,
 * 2 output overrun, 3 input overrun, 4 CRC mismatch; d_crc_found[i] = CRC-32 of the decoded block.
 * d_comp must be readable for 8 bytes past the last member's payload (the bit reader loads aligned 32-bit word
 * pairs; it never starts a load at or beyond a member's end, whatever the data says). */
int gzpb_decode_device(gzpb_decoder *d, const void *d_comp, const gzpb_block_desc *d_desc, size_t nblocks, void *d_out,
                       int32_t *d_status, uint32_t *d_crc_found, void *cuda_stream);
/* Incremental reader = `ParDecompress<F>` as a C object (src/par/decompress.rs:113-352): `source` plays
 * `R: Read` (returns bytes read, 0 at EOF, negative on error); gzpb_reader_read is `Read::read` (:238-287):
 * decoded bytes in stream order, 0 at the end of the stream, a negative status on error (errors are sticky;
 * GZPB_ECHECK details through gzpb_reader_last_check).  Compressed bytes are pulled `chunk_bytes` at a time
 * (0 = 64 MiB) into pinned memory and decoded member-parallel on the GPU. */
typedef long (*gzpb_source_fn)(void *user, void *buf, size_t cap);
typedef struct gzpb_reader gzpb_reader;
int gzpb_reader_create(gzpb_reader **r, int device, int format, size_t blocks_in_flight, size_t chunk_bytes,
                       gzpb_source_fn source, void *user);
long gzpb_reader_read(gzpb_reader *r, void *buf, size_t len);
int gzpb_reader_last_check(gzpb_reader *r, uint32_t *found, uint32_t *expected);
int gzpb_reader_finish(gzpb_reader *r);
void gzpb_reader_destroy(gzpb_reader *r);
/* InvalidCheck{found, expected} of the last GZPB_ECHECK (and the index of the failing block). */
int gzpb_decoder_last_check(gzpb_decoder *d, uint32_t *found, uint32_t *expected, uint64_t *block_index);
int gzpb_decoder_set_profiling(gzpb_decoder *d, int on);
int gzpb_decoder_kernel_ms(gzpb_decoder *d, double *total_ms, uint64_t *launches);
uint64_t gzpb_decoder_launch_count(gzpb_decoder *d);

/* BGZF block index (.gzi, the layout htslib's `bgzip -i` writes: u64 count, then {u64 compressed offset,
 * u64 uncompressed offset} per data block after the first) and virtual offsets (block_offset << 16 |
 * offset within the block) — SURVEY.md §8(f) rank 2; a TODO of the reference (README.md:161).
 * `out` may be NULL to size the index. */
int gzpb_bgzf_index(const void *bgzf, size_t len, void *out, size_t out_cap, size_t *out_len);
/* The same index straight from a Bgzf writer: it knows every block's compressed size when the block's batch
 * retires, so the .gzi is a by-product of writing (covers the blocks handed to the sink so far; call after
 * gzpb_writer_finish for the whole stream).  `out` may be NULL to size the index. */
int gzpb_writer_bgzf_index(gzpb_writer *w, void *out, size_t out_cap, size_t *out_len);
uint64_t gzpb_bgzf_virtual_offset(uint64_t block_offset, uint32_t within_block);

const char *gzpb_strerror(int code);
const char *gzpb_version(void);

#ifdef __cplusplus
}
#endif
#endif
