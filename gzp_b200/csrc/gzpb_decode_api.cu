// gzpb_decode_api.cu — C ABI of the block-parallel DECODE path (include/gzpb.h,
// "decoder" section): the reader loop of ParDecompress::run restated on the host
// (/root/reference/src/par/decompress.rs:190-207: read HEADER_SIZE bytes,
// check_header, get_block_size, read the remainder) and the worker body
// (:163-187) on the GPU (inflate_kernels.cu).  Blocks are decoded straight to their
// final offsets, which the host knows from the ISIZE fields before any block is
// decoded, so ordering (the FIFO of oneshot receivers, :203-204) needs no work.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/gzpb.h"
#include "gzpb_common.cuh"
#include "inflate_kernels.cuh"

using namespace gzpb;

static_assert(sizeof(gzpb_block_desc) == sizeof(InflateDesc), "public and device descriptor layouts must match");

#define DCK(x)                                                            \
    do {                                                                  \
        cudaError_t e_ = (x);                                             \
        if (e_ != cudaSuccess) {                                          \
            fprintf(stderr, "[gzpb] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return GZPB_ECUDA;                                            \
        }                                                                 \
    } while (0)

namespace {
constexpr int kDLanes = 3;
struct DLane {
    cudaStream_t st = nullptr;
    cudaEvent_t ev_done = nullptr;
    uint8_t *d_comp = nullptr, *d_out = nullptr;
    InflateDesc *d_desc = nullptr, *h_desc = nullptr;
    int32_t *d_status = nullptr, *h_status = nullptr;
    uint32_t *d_crc = nullptr, *h_crc = nullptr;
    size_t comp_cap = 0, out_cap = 0;
    bool busy = false;
};
}  // namespace

struct gzpb_decoder {
    int device = 0, format = 0;
    size_t max_blocks = 0;
    DLane lanes[kDLanes];
    KernelTimer timer;
    bool profiling = false;
    uint64_t launches = 0;
    uint32_t check_found = 0, check_expected = 0;
    uint64_t bad_block = 0;
};

static size_t header_size(int format) { return format == GZPB_BGZF ? 18 : format == GZPB_MGZIP ? 20 : 0; }

extern "C" size_t gzpb_block_header_size(int format) { return header_size(format); }

// BlockFormatSpec::check_header + get_block_size (deflate.rs:407-422, 555-570)
extern "C" long gzpb_block_size(int format, const void *hdr_v, size_t avail)
{
    const uint8_t *h = (const uint8_t *)hdr_v;
    const size_t hs = header_size(format);
    if (!hs || !h) return GZPB_EINVAL;
    if (avail < hs) return GZPB_EIO;
    if ((h[3] & 4) != 4) return GZPB_EHEADER;                         // "Extra field flag not set"
    if (format == GZPB_BGZF) {
        if (h[12] != 'B' || h[13] != 'C') return GZPB_EHEADER;        // "Bad SID"
        return (long)((uint32_t)h[16] | ((uint32_t)h[17] << 8)) + 1;
    }
    if (h[12] != 'I' || h[13] != 'G') return GZPB_EHEADER;
    return (long)((uint32_t)h[16] | ((uint32_t)h[17] << 8) | ((uint32_t)h[18] << 16) | ((uint32_t)h[19] << 24));
}

// The reader loop: walk the members of `in`, one descriptor per block.
extern "C" int gzpb_scan_blocks(int format, const void *in_v, size_t in_len, gzpb_block_desc *descs, size_t max_descs,
                                size_t *nblocks, size_t *consumed, uint64_t *total_out)
{
    const uint8_t *in = (const uint8_t *)in_v;
    const size_t hs = header_size(format);
    if (!hs || (in_len && !in)) return GZPB_EINVAL;
    size_t pos = 0, n = 0;
    uint64_t opos = 0;
    int rc = GZPB_OK;
    while (in_len - pos >= hs) {                                       // a short trailing header = EOF (decompress.rs:193, 205-206)
        const long size = gzpb_block_size(format, in + pos, in_len - pos);
        if (size < 0) { rc = (int)size; break; }
        if ((size_t)size < hs + 8) { rc = GZPB_EBLOCK; break; }
        if ((size_t)size > in_len - pos) { rc = GZPB_EIO; break; }     // read_exact of the remainder fails (UnexpectedEof)
        const uint8_t *f = in + pos + size - 8;                        // get_footer_values (lib.rs:440-447)
        const uint32_t crc = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
        const uint32_t isize = (uint32_t)f[4] | ((uint32_t)f[5] << 8) | ((uint32_t)f[6] << 16) | ((uint32_t)f[7] << 24);
        // ISIZE sizes the output buffers before anything is decoded: DEFLATE expands at most 1032:1, so a footer that
        // claims more than its payload can hold is corrupt data — not a multi-GiB allocation request
        if ((uint64_t)isize > 1032ull * (uint64_t)((size_t)size - hs - 8) + 64) { rc = GZPB_EDECOMPRESS; break; }
        if (descs) {
            if (n >= max_descs) break;
            gzpb_block_desc &d = descs[n];
            d.in_off = pos + hs; d.in_len = (uint32_t)(size - hs - 8); d.out_off = opos; d.out_len = isize; d.crc = crc; d.pad = 0;
        }
        n++; opos += isize; pos += (size_t)size;
    }
    if (nblocks) *nblocks = n;
    if (consumed) *consumed = pos;
    if (total_out) *total_out = opos;
    return rc;
}

static void dlane_free(DLane &L)
{
    cudaFree(L.d_comp); cudaFree(L.d_out); cudaFree(L.d_desc); cudaFree(L.d_status); cudaFree(L.d_crc);
    cudaFreeHost(L.h_desc); cudaFreeHost(L.h_status); cudaFreeHost(L.h_crc);
    if (L.ev_done) cudaEventDestroy(L.ev_done);
    if (L.st) cudaStreamDestroy(L.st);
    L = DLane();
}

static int dlane_reserve(DLane &L, size_t comp_bytes, size_t out_bytes)
{
    if (comp_bytes + 64 > L.comp_cap) {
        cudaFree(L.d_comp); L.d_comp = nullptr;
        L.comp_cap = comp_bytes + comp_bytes / 4 + 4096;
        DCK(cudaMalloc((void **)&L.d_comp, L.comp_cap));
    }
    if (out_bytes + 64 > L.out_cap) {
        cudaFree(L.d_out); L.d_out = nullptr;
        L.out_cap = out_bytes + out_bytes / 4 + 4096;
        DCK(cudaMalloc((void **)&L.d_out, L.out_cap));
    }
    return GZPB_OK;
}

extern "C" int gzpb_decoder_create(gzpb_decoder **out, int device, int format, size_t max_blocks_in_flight)
{
    if (!out) return GZPB_EINVAL;
    *out = nullptr;
    if (format != GZPB_BGZF && format != GZPB_MGZIP) return GZPB_EINVAL;   // BlockFormatSpec is implemented for these two (deflate.rs:359, 508)
    if (max_blocks_in_flight == 0) max_blocks_in_flight = 5328;   // 148 SMs x 36 resident decode warps
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return GZPB_ECUDA;
    DCK(cudaSetDevice(device));
    cudaDeviceProp prop;
    DCK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return GZPB_ECUDA;                                // sm_100a only; no CPU fallback
    gzpb_decoder *d = new gzpb_decoder();
    d->device = device; d->format = format; d->max_blocks = max_blocks_in_flight;
    upload_inflate_constants();
    for (int i = 0; i < kDLanes; i++) {
        DLane &L = d->lanes[i];
        const size_t U = d->max_blocks;
        bool ok = cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming) == cudaSuccess &&
                  cudaMalloc((void **)&L.d_desc, U * sizeof(InflateDesc)) == cudaSuccess &&
                  cudaMalloc((void **)&L.d_status, U * sizeof(int32_t)) == cudaSuccess &&
                  cudaMalloc((void **)&L.d_crc, U * sizeof(uint32_t)) == cudaSuccess &&
                  cudaHostAlloc((void **)&L.h_desc, U * sizeof(InflateDesc), cudaHostAllocPortable) == cudaSuccess &&
                  cudaHostAlloc((void **)&L.h_status, U * sizeof(int32_t), cudaHostAllocPortable) == cudaSuccess &&
                  cudaHostAlloc((void **)&L.h_crc, U * sizeof(uint32_t), cudaHostAllocPortable) == cudaSuccess;
        if (!ok) { gzpb_decoder_destroy(d); return GZPB_ECUDA; }
    }
    DCK(cudaDeviceSynchronize());
    *out = d;
    return GZPB_OK;
}

extern "C" void gzpb_decoder_destroy(gzpb_decoder *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kDLanes; i++) dlane_free(d->lanes[i]);
    d->timer.collect();
    for (auto e : d->timer.pool) cudaEventDestroy(e);
    delete d;
}

extern "C" int gzpb_decoder_last_check(gzpb_decoder *d, uint32_t *found, uint32_t *expected, uint64_t *block_index)
{
    if (!d) return GZPB_EINVAL;
    if (found) *found = d->check_found;
    if (expected) *expected = d->check_expected;
    if (block_index) *block_index = d->bad_block;
    return GZPB_OK;
}

extern "C" int gzpb_decoder_set_profiling(gzpb_decoder *d, int on)
{
    if (!d) return GZPB_EINVAL;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    d->timer.reset();
    d->profiling = on != 0;
    return GZPB_OK;
}

extern "C" int gzpb_decoder_kernel_ms(gzpb_decoder *d, double *total_ms, uint64_t *launches)
{
    if (!d) return GZPB_EINVAL;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();
    d->timer.collect();
    if (total_ms) *total_ms = d->timer.total_ms[KT_INFLATE];
    if (launches) *launches = d->timer.launches[KT_INFLATE];
    return GZPB_OK;
}

extern "C" uint64_t gzpb_decoder_launch_count(gzpb_decoder *d) { return d ? d->launches : 0; }

// Device-resident form: descriptors, compressed bytes and output already in HBM.
extern "C" int gzpb_decode_device(gzpb_decoder *d, const void *d_comp, const gzpb_block_desc *d_desc, size_t nblocks, void *d_out,
                                  int32_t *d_status, uint32_t *d_crc_found, void *cuda_stream)
{
    if (!d || (nblocks && (!d_comp || !d_desc || !d_out || !d_status || !d_crc_found))) return GZPB_EINVAL;
    DCK(cudaSetDevice(d->device));
    InflateBatch b;
    b.nblocks = (uint32_t)nblocks; b.comp = (const uint8_t *)d_comp; b.desc = (const InflateDesc *)d_desc; b.out = (uint8_t *)d_out;
    b.status = d_status; b.crc_found = d_crc_found; b.timer = d->profiling ? &d->timer : nullptr;
    DCK(launch_inflate(b, (cudaStream_t)cuda_stream));
    d->launches += 1;
    return GZPB_OK;
}

// ParDecompress over an in-memory input (header walk on the host, blocks on the GPU).
extern "C" int gzpb_decode_stream(gzpb_decoder *d, const void *in_v, size_t in_len, void *out_v, size_t out_cap, size_t *out_len,
                                  size_t *consumed)
{
    if (!d || !out_len || (in_len && !in_v)) return GZPB_EINVAL;
    *out_len = 0;
    if (consumed) *consumed = 0;
    DCK(cudaSetDevice(d->device));
    const uint8_t *in = (const uint8_t *)in_v;
    uint8_t *out = (uint8_t *)out_v;

    size_t nblocks = 0, used = 0;
    uint64_t total = 0;
    int rc = gzpb_scan_blocks(d->format, in, in_len, nullptr, 0, &nblocks, &used, &total);
    // an incomplete trailing block is left for the next call when the caller tracks `consumed`
    if (rc == GZPB_EIO && consumed) rc = GZPB_OK;
    const int scan_rc = rc;                                            // a header error surfaces after the blocks before it
    if (total > out_cap) return GZPB_ECOMPRESS;
    if (total && !out) return GZPB_EINVAL;
    std::vector<gzpb_block_desc> descs(nblocks ? nblocks : 1);
    {
        size_t n2 = 0;
        int rc2 = gzpb_scan_blocks(d->format, in, used, descs.data(), nblocks, &n2, nullptr, nullptr);
        if (rc2 != GZPB_OK || n2 != nblocks) return GZPB_EINVAL;
    }

    const size_t hs = header_size(d->format);
    struct Pending { size_t first, count; int lane; size_t out0, out_bytes; };
    std::vector<Pending> pend;
    rc = GZPB_OK;
    auto retire = [&](const Pending &p) -> int {
        DLane &L = d->lanes[p.lane];
        if (cudaEventSynchronize(L.ev_done) != cudaSuccess) return GZPB_ECUDA;
        if (d->profiling) d->timer.collect();
        L.busy = false;
        for (size_t i = 0; i < p.count; i++) {
            const int32_t s = L.h_status[i];
            if (s == 0) continue;
            d->bad_block = p.first + i;
            if (s == 4) { d->check_found = L.h_crc[i]; d->check_expected = descs[p.first + i].crc; return GZPB_ECHECK; }
            return GZPB_EDECOMPRESS;
        }
        return GZPB_OK;
    };

    size_t done = 0;
    int li = 0;
    while (done < nblocks && rc == GZPB_OK) {
        // batch = consecutive blocks, bounded by count and by ~256 MiB of output
        size_t cnt = 0, obytes = 0;
        while (done + cnt < nblocks && cnt < d->max_blocks && (cnt == 0 || obytes + descs[done + cnt].out_len <= (256u << 20))) {
            obytes += descs[done + cnt].out_len; cnt++;
        }
        DLane &L = d->lanes[li];
        if (L.busy) {
            rc = retire(pend.front()); pend.erase(pend.begin());
            if (rc != GZPB_OK) break;
        }
        const size_t c0 = (size_t)descs[done].in_off - hs;                                     // first header of the batch
        const size_t c1 = (size_t)descs[done + cnt - 1].in_off + descs[done + cnt - 1].in_len + 8;   // end of the last footer
        const size_t o0 = (size_t)descs[done].out_off;
        rc = dlane_reserve(L, c1 - c0, obytes);
        if (rc != GZPB_OK) break;
        for (size_t i = 0; i < cnt; i++) {
            const gzpb_block_desc &s = descs[done + i];
            InflateDesc &t = L.h_desc[i];
            t.in_off = s.in_off - c0; t.in_len = s.in_len; t.out_off = s.out_off - o0; t.out_len = s.out_len; t.crc = s.crc; t.pad = 0;
        }
        DCK(cudaMemcpyAsync(L.d_comp, in + c0, c1 - c0, cudaMemcpyHostToDevice, L.st));
        DCK(cudaMemcpyAsync(L.d_desc, L.h_desc, cnt * sizeof(InflateDesc), cudaMemcpyHostToDevice, L.st));
        InflateBatch b;
        b.nblocks = (uint32_t)cnt; b.comp = L.d_comp; b.desc = L.d_desc; b.out = L.d_out; b.status = L.d_status; b.crc_found = L.d_crc;
        b.timer = d->profiling ? &d->timer : nullptr;
        DCK(launch_inflate(b, L.st));
        d->launches += 1;
        if (obytes) DCK(cudaMemcpyAsync(out + o0, L.d_out, obytes, cudaMemcpyDeviceToHost, L.st));
        DCK(cudaMemcpyAsync(L.h_status, L.d_status, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, L.st));
        DCK(cudaMemcpyAsync(L.h_crc, L.d_crc, cnt * sizeof(uint32_t), cudaMemcpyDeviceToHost, L.st));
        DCK(cudaEventRecord(L.ev_done, L.st));
        L.busy = true;
        pend.push_back(Pending{done, cnt, li, o0, obytes});
        done += cnt;
        li = (li + 1) % kDLanes;
    }
    for (auto &p : pend) {
        int r = retire(p);
        if (rc == GZPB_OK) rc = r;
    }
    if (rc != GZPB_OK) { for (int i = 0; i < kDLanes; i++) d->lanes[i].busy = false; cudaDeviceSynchronize(); return rc; }
    *out_len = (size_t)total;
    if (consumed) *consumed = used;
    return scan_rc;
}

// ---- incremental reader: ParDecompress<F> as a C object ------------------------------------
// The reader loop of ParDecompress::run (par/decompress.rs:190-207) pulls bytes from `source` (= `R: Read`) into a
// pinned buffer, whole members go to the GPU in one gzpb_decode_stream call (straight from / to pinned memory), an
// incomplete trailing member stays for the next round; `read` hands the decoded bytes out in stream order
// (:238-287).  A short trailing header is EOF (:193, 205-206), a truncated member is an Io error (:197).
struct gzpb_reader {
    gzpb_decoder *dec = nullptr;
    int format = 0;
    gzpb_source_fn source = nullptr;
    void *user = nullptr;
    uint8_t *in = nullptr, *out = nullptr;       // pinned
    size_t in_cap = 0, in_len = 0, out_cap = 0, out_len = 0, out_pos = 0, chunk = 0;
    bool eof = false;
    int error = GZPB_OK;
    int pending_error = GZPB_OK;                 // a bad member behind good ones: the good ones are handed out first
    uint64_t bytes_in = 0, bytes_out = 0;
};

static int reader_grow(uint8_t **buf, size_t *cap, size_t keep, size_t need)
{
    if (need <= *cap) return GZPB_OK;
    size_t ncap = std::max(need, *cap + *cap / 2);
    uint8_t *p = nullptr;
    if (cudaHostAlloc((void **)&p, ncap, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return GZPB_ENOMEM; }
    if (keep) memcpy(p, *buf, keep);
    cudaFreeHost(*buf);
    *buf = p; *cap = ncap;
    return GZPB_OK;
}

extern "C" int gzpb_reader_create(gzpb_reader **out, int device, int format, size_t blocks_in_flight, size_t chunk_bytes,
                                  gzpb_source_fn source, void *user)
{
    if (!out || !source) return GZPB_EINVAL;
    *out = nullptr;
    if (!header_size(format)) return GZPB_EINVAL;                      // ParDecompress needs a BlockFormatSpec (Mgzip, Bgzf)
    gzpb_reader *r = new gzpb_reader();
    int rc = gzpb_decoder_create(&r->dec, device, format, blocks_in_flight ? blocks_in_flight : 5328);
    if (rc != GZPB_OK) { delete r; return rc; }
    r->format = format; r->source = source; r->user = user;
    r->chunk = chunk_bytes ? std::max(chunk_bytes, (size_t)65536) : ((size_t)64 << 20);
    *out = r;
    return GZPB_OK;
}

// one round of the reader loop: pull up to `chunk` bytes, decode the whole members that are there
static int reader_fill(gzpb_reader *r)
{
    if (r->pending_error) return r->pending_error;
    int rc = reader_grow(&r->in, &r->in_cap, r->in_len, r->in_len + r->chunk);
    if (rc != GZPB_OK) return rc;
    while (!r->eof && r->in_len < r->in_cap) {                          // a source may return short reads
        const long k = r->source(r->user, r->in + r->in_len, r->in_cap - r->in_len);
        if (k < 0) return GZPB_EIO;
        if (k == 0) { r->eof = true; break; }
        r->in_len += (size_t)k; r->bytes_in += (uint64_t)k;
        if (r->in_len >= r->chunk) break;
    }
    size_t consumed = 0, nb = 0;
    uint64_t total = 0;
    rc = gzpb_scan_blocks(r->format, r->in, r->in_len, nullptr, 0, &nb, &consumed, &total);
    if (rc == GZPB_EIO && !r->eof) rc = GZPB_OK;                        // incomplete trailing member: wait for more bytes
    if (rc == GZPB_OK && r->eof && r->in_len - consumed >= header_size(r->format)) rc = GZPB_EIO;
    if (rc != GZPB_OK) {
        // like the reference's reader thread (decompress.rs:190-207), which has already sent the members in front of a
        // bad one to the workers: they are decoded and delivered, the error surfaces on the read after them
        if (nb == 0) return rc;
        r->pending_error = rc;
    }
    if (nb == 0) {
        if (r->eof) r->in_len = 0;                                      // a short trailing header is EOF
        return GZPB_OK;                                                 // (a member larger than the chunk: the buffer grows next round)
    }
    rc = reader_grow(&r->out, &r->out_cap, 0, (size_t)total + 64);
    if (rc != GZPB_OK) return rc;
    size_t olen = 0, used = 0;
    rc = gzpb_decode_stream(r->dec, r->in, consumed, r->out, r->out_cap, &olen, &used);
    if (rc != GZPB_OK) return rc;
    r->out_len = olen; r->out_pos = 0;
    memmove(r->in, r->in + used, r->in_len - used);
    r->in_len -= used;
    return GZPB_OK;
}

extern "C" long gzpb_reader_read(gzpb_reader *r, void *buf, size_t len)
{
    if (!r || (len && !buf)) return GZPB_EINVAL;
    if (r->error) return r->error;
    size_t done = 0;
    while (done < len) {
        if (r->out_pos == r->out_len) {
            if (r->eof && r->in_len == 0) break;
            r->out_pos = r->out_len = 0;
            const int rc = reader_fill(r);
            if (rc != GZPB_OK) { r->error = rc; return done ? (long)done : rc; }
            if (r->out_len == 0 && r->eof && r->in_len == 0) break;
            continue;
        }
        const size_t k = std::min(len - done, r->out_len - r->out_pos);
        memcpy((uint8_t *)buf + done, r->out + r->out_pos, k);
        r->out_pos += k; done += k; r->bytes_out += k;
    }
    return (long)done;
}

extern "C" int gzpb_reader_last_check(gzpb_reader *r, uint32_t *found, uint32_t *expected)
{
    if (!r) return GZPB_EINVAL;
    return gzpb_decoder_last_check(r->dec, found, expected, nullptr);
}

extern "C" int gzpb_reader_finish(gzpb_reader *r) { return r ? r->error : GZPB_EINVAL; }   // "close things in such a way as to get errors" (:222-236)

extern "C" void gzpb_reader_destroy(gzpb_reader *r)
{
    if (!r) return;
    if (r->dec) gzpb_decoder_destroy(r->dec);
    cudaFreeHost(r->in); cudaFreeHost(r->out);
    delete r;
}

// ---- BGZF block index (.gzi) and virtual offsets (a TODO of the reference,
// /root/reference/README.md:161).  Layout as written by htslib's `bgzip -i`: u64 LE entry count, then per
// entry {u64 LE compressed offset, u64 LE uncompressed offset} of every data block after the first.
extern "C" int gzpb_bgzf_index(const void *bgzf_v, size_t len, void *out_v, size_t out_cap, size_t *out_len)
{
    const uint8_t *in = (const uint8_t *)bgzf_v;
    uint8_t *out = (uint8_t *)out_v;
    if (!out_len || (len && !in)) return GZPB_EINVAL;
    size_t pos = 0;
    uint64_t upos = 0, n = 0;
    bool first = true;
    auto put64 = [&](size_t at, uint64_t v) { for (int i = 0; i < 8; i++) out[at + i] = (uint8_t)(v >> (8 * i)); };
    while (len - pos >= 18) {
        const long size = gzpb_block_size(GZPB_BGZF, in + pos, len - pos);
        if (size < 0) return (int)size;
        if ((size_t)size < 26) return GZPB_EBLOCK;
        if ((size_t)size > len - pos) return GZPB_EIO;
        const uint8_t *f = in + pos + size - 4;
        const uint32_t isize = (uint32_t)f[0] | ((uint32_t)f[1] << 8) | ((uint32_t)f[2] << 16) | ((uint32_t)f[3] << 24);
        if (isize) {                                   // the EOF marker and empty flush blocks carry no data
            if (!first) {
                const size_t at = 8 + 16 * (size_t)n;
                if (out) { if (at + 16 > out_cap) return GZPB_ECOMPRESS; put64(at, pos); put64(at + 8, upos); }
                n++;
            }
            first = false;
        }
        upos += isize; pos += (size_t)size;
    }
    if (out) { if (out_cap < 8) return GZPB_ECOMPRESS; put64(0, n); }
    *out_len = 8 + 16 * (size_t)n;
    return GZPB_OK;
}

extern "C" uint64_t gzpb_bgzf_virtual_offset(uint64_t block_offset, uint32_t within_block)
{
    return (block_offset << 16) | (within_block & 0xFFFFu);
}
