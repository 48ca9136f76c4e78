// gzpb_api.cu — C-ABI implementation (include/gzpb.h): device context, batch
// staging (pinned memory, H2D/D2H overlap across lanes), per-format unit
// preparation, and the host-side FormatSpec helpers (header / footer / combine).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/gzpb.h"
#include "deflate_kernels.cuh"
#include "snappy_kernels.cuh"
#include "gzpb_common.cuh"

using namespace gzpb;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            snprintf(g_last_cuda_error, sizeof g_last_cuda_error, "%s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return GZPB_ECUDA;                                                             \
        }                                                                                  \
    } while (0)

static thread_local char g_last_cuda_error[256];

namespace {

constexpr int kLanes = 3;

struct Lane {
    cudaStream_t st = nullptr;
    cudaEvent_t ev_scan = nullptr, ev_done = nullptr;
    // device
    uint8_t *d_in = nullptr;
    uint32_t *d_len = nullptr, *d_dict = nullptr, *d_flags = nullptr, *d_crc = nullptr, *d_sum_part = nullptr, *d_tokens = nullptr, *d_out_len = nullptr;
    uint16_t *d_next4 = nullptr, *d_prev3 = nullptr, *d_order = nullptr;
    uint8_t *d_clen = nullptr;
    uint64_t *d_mtab = nullptr, *d_offsets = nullptr;
    uint32_t *d_mtab2 = nullptr, *d_lists = nullptr, *d_list_start = nullptr;
    uint8_t *d_out = nullptr, *d_packed = nullptr;
    int32_t *d_status = nullptr, *d_overflow = nullptr;
    uint32_t *d_comb = nullptr, *h_comb = nullptr;   // batch-combined check {sum, len lo, len hi}
    uint64_t *h_end = nullptr;                       // mapped pinned: stream offset behind this lane's batch (the offset chain of a multi-device stream)
    // pinned host
    uint8_t *h_in = nullptr, *h_packed = nullptr;
    uint32_t *h_len = nullptr, *h_dict = nullptr, *h_flags = nullptr, *h_crc = nullptr;
    uint64_t *h_offsets = nullptr;
    int32_t *h_status = nullptr, *h_overflow = nullptr;
    size_t nunits = 0;
    bool busy = false;
};

}  // namespace

struct gzpb_ctx {
    int device = 0, format = 0, level = 0;
    size_t max_block_bytes = 0, max_units = 0;
    uint32_t dict_cap = 0;                         // 32 KiB for the dictionary formats, else 0
    uint32_t in_stride = 0, m_stride = 0, tok_stride = 0, out_stride = 0, spu = 1, seg = 0;
    int check_kind = -1;
    uint32_t cpu = 1;                              // gather entries per unit (Snap: 64 KiB chunks per block)
    Lane lanes[kLanes];
    int gather_host_ctas = 64;                       // thread blocks of a k_gather that writes host memory (GZPB_GATHER_CTAS; measured 37/74/148/296/592/one per unit: e2e 9.09/9.04/9.01 | 8.68/8.55/8.45 GiB/s)
    uint64_t *d_base0 = nullptr, *h_base0 = nullptr; // first stream offset of a gzpb_encode_stream call: device copy / mapped pinned copy
    bool scratch_only = false;
    KernelTimer timer;
    bool profiling = false;
    uint64_t launches = 0;
    // gzpb_submit / gzpb_poll: batches in flight, oldest first (tickets complete in submission order)
    struct Ticket { uint64_t id; const gzpb_block_in *in; gzpb_block_out *out; size_t count; int lane; };
    std::deque<Ticket> tickets;
    uint64_t next_ticket = 1;
    int next_lane = 0;
};

static bool is_deflate_format(int f) { return f == GZPB_GZIP || f == GZPB_ZLIB || f == GZPB_RAWDEFLATE || f == GZPB_MGZIP || f == GZPB_BGZF; }

static size_t extra_amount(size_t n) { size_t e = (size_t)((double)n * 0.1); return e > 128 ? e : 128; }

extern "C" size_t gzpb_encode_capacity(int format, size_t n)
{
    switch (format) {
    case GZPB_BGZF: return 18 + n + extra_amount(n) + 8 + 28;
    case GZPB_MGZIP: return 20 + n + extra_amount(n) + 8;
    case GZPB_SNAP: return 10 + ((n + 65535) / 65536) * 8 + (32 + n + n / 6) + 64;
    default: return n + extra_amount(n);
    }
}

extern "C" size_t gzpb_default_bufsize(int format) { return format == GZPB_BGZF ? GZPB_BGZF_BLOCK_SIZE : GZPB_BUFSIZE; }
extern "C" int gzpb_needs_dict(int format) { return format == GZPB_GZIP || format == GZPB_ZLIB || format == GZPB_RAWDEFLATE; }

extern "C" int gzpb_level_supported(int format, int level)
{
    if (format == GZPB_SNAP) return 1;
    LevelParams lp;
    return level_params(level, &lp) ? 1 : 0;
}

static int xfl(int level) { return level >= 9 ? 2 : level <= 1 ? 4 : 0; }

extern "C" size_t gzpb_header(int format, int level, void *buf)
{
    uint8_t *o = (uint8_t *)buf;
    if (format == GZPB_GZIP) {
        const uint8_t h[10] = {31, 139, 8, 0, 0, 0, 0, 0, (uint8_t)xfl(level), 255};
        memcpy(o, h, 10);
        return 10;
    }
    if (format == GZPB_ZLIB) {
        uint32_t cv = level >= 9 ? 3u << 6 : level == 1 ? 0 : level >= 6 ? 1u << 6 : 2u << 6;
        uint32_t head = (0x78u << 8) + cv;
        head += 31 - (head % 31);
        o[0] = (uint8_t)(head >> 8); o[1] = (uint8_t)head;
        return 2;
    }
    return 0;
}

extern "C" size_t gzpb_footer(int format, uint32_t sum, uint32_t amount, void *buf)
{
    uint8_t *o = (uint8_t *)buf;
    if (format == GZPB_GZIP) {
        for (int i = 0; i < 4; i++) { o[i] = (uint8_t)(sum >> (8 * i)); o[4 + i] = (uint8_t)(amount >> (8 * i)); }
        return 8;
    }
    if (format == GZPB_ZLIB) {
        for (int i = 0; i < 4; i++) o[i] = (uint8_t)(sum >> (24 - 8 * i));
        return 4;
    }
    return 0;
}

extern "C" uint32_t gzpb_crc32_combine(uint32_t a, uint32_t b, uint64_t len_b)
{
    if (len_b == 0) return a;
    return gf2_mulmod(a, gf2_xpow8(len_b, kCrcPoly), kCrcPoly) ^ b;
}

extern "C" uint32_t gzpb_adler32_combine(uint32_t a1, uint32_t a2, uint64_t len2)
{
    const uint32_t BASE = 65521u;
    uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t sum1 = a1 & 0xFFFF;
    uint32_t sum2 = (rem * sum1) % BASE;
    sum1 += (a2 & 0xFFFF) + BASE - 1;
    sum2 += (a1 >> 16) + (a2 >> 16) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return sum1 | (sum2 << 16);
}

extern "C" const char *gzpb_strerror(int code)
{
    switch (code) {
    case GZPB_OK: return "ok";
    case GZPB_EBUFFERSIZE: return "Invalid buffer size, must be >= 32768";
    case GZPB_ENUMTHREADS: return "Invalid number of threads selected";
    case GZPB_EBLOCKSIZE: return "Compressed block size exceeds max allowed (65536), try increasing compression";
    case GZPB_ECOMPRESS: return "compression error: insufficient space in the output buffer";
    case GZPB_ELEVEL: return "compression level not supported by the B200 engine";
    case GZPB_EIO: return "io error";
    case GZPB_ECHANNEL: return "failed to send over channel (stream finished)";
    case GZPB_ECUDA: return g_last_cuda_error[0] ? g_last_cuda_error : "CUDA error";
    case GZPB_EINVAL: return "invalid argument";
    case GZPB_ENOMEM: return "out of memory";
    case GZPB_EHEADER: return "Invalid block header";
    case GZPB_ECHECK: return "Invalid checksum";
    case GZPB_EDECOMPRESS: return "decompression error: corrupt DEFLATE data";
    case GZPB_EBLOCK: return "Invalid block size";
    default: return "unknown";
    }
}

extern "C" const char *gzpb_version(void) { return "gzp-b200 0.1 (sm_100a)"; }

extern "C" void *gzpb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void gzpb_host_free(void *p) { if (p) cudaFreeHost(p); }

template <typename T>
static cudaError_t dmalloc(T **p, size_t count) { return cudaMalloc((void **)p, count * sizeof(T)); }
template <typename T>
static cudaError_t hmalloc(T **p, size_t count) { return cudaHostAlloc((void **)p, count * sizeof(T), cudaHostAllocPortable | cudaHostAllocMapped); }

static int lane_alloc(gzpb_ctx *c, Lane &L, bool with_io)
{
    const size_t U = c->max_units;
    CK(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&L.ev_scan, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
    if (c->format == GZPB_SNAP) {
        const size_t E = U * c->cpu;
        CK(dmalloc(&L.d_out, E * c->out_stride + 256));
        CK(dmalloc(&L.d_out_len, E * 2));
        CK(dmalloc(&L.d_overflow, 1));
        CK(cudaMemset(L.d_overflow, 0, sizeof(int32_t)));
        CK(dmalloc(&L.d_in, U * c->in_stride + 256));
        CK(dmalloc(&L.d_len, U));
        CK(dmalloc(&L.d_offsets, E + 1));
        CK(hmalloc(&L.h_len, U));
        CK(hmalloc(&L.h_dict, U));
        CK(hmalloc(&L.h_flags, U));
        CK(hmalloc(&L.h_crc, U));
        CK(hmalloc(&L.h_offsets, E + 1));
        CK(hmalloc(&L.h_status, U));
        CK(hmalloc(&L.h_overflow, 1));
        CK(hmalloc(&L.h_end, 2));
        memset(L.h_status, 0, U * sizeof(int32_t));
        memset(L.h_crc, 0, U * sizeof(uint32_t));
        return GZPB_OK;
    }
    CK(dmalloc(&L.d_next4, U * (size_t)c->m_stride + 64));      // per unit position (k_link carries its heads across sub-units)
    CK(dmalloc(&L.d_prev3, U * (size_t)c->m_stride + 64));
    CK(dmalloc(&L.d_order, U * c->spu * kMaxUnitBytes));
    CK(dmalloc(&L.d_lists, U * c->spu * 2 * kMaxUnitBytes));
    CK(dmalloc(&L.d_list_start, U * c->spu * 64));       // kLsStride
    CK(dmalloc(&L.d_clen, U * (size_t)c->m_stride + 64));
    CK(dmalloc(&L.d_mtab, U * c->m_stride));
    if (c->level >= 8) CK(dmalloc(&L.d_mtab2, U * c->m_stride));
    CK(dmalloc(&L.d_crc, U));
    CK(dmalloc(&L.d_sum_part, U * c->spu));
    CK(dmalloc(&L.d_tokens, U * c->tok_stride));
    CK(dmalloc(&L.d_out, U * c->out_stride + 256));
    CK(dmalloc(&L.d_out_len, U * 2));
    CK(dmalloc(&L.d_overflow, 1));
    CK(cudaMemset(L.d_overflow, 0, sizeof(int32_t)));
    CK(dmalloc(&L.d_dict, U));
    CK(cudaMemset(L.d_dict, 0, U * sizeof(uint32_t)));
    CK(dmalloc(&L.d_comb, 4));
    CK(hmalloc(&L.h_comb, 4));
    if (with_io) {
        CK(dmalloc(&L.d_in, U * c->in_stride + 256));
        CK(dmalloc(&L.d_len, U));
        CK(dmalloc(&L.d_flags, U));
        CK(dmalloc(&L.d_offsets, U + 1));
        CK(dmalloc(&L.d_status, U));
        CK(hmalloc(&L.h_len, U));
        CK(hmalloc(&L.h_dict, U));
        CK(hmalloc(&L.h_flags, U));
        CK(hmalloc(&L.h_crc, U));
        CK(hmalloc(&L.h_offsets, U + 1));
        CK(hmalloc(&L.h_status, U));
        CK(hmalloc(&L.h_overflow, 1));
        CK(hmalloc(&L.h_end, 2));
    }
    return GZPB_OK;
}

static void lane_free(Lane &L)
{
    cudaFree(L.d_in); cudaFree(L.d_len); cudaFree(L.d_dict); cudaFree(L.d_flags); cudaFree(L.d_crc); cudaFree(L.d_sum_part); cudaFree(L.d_tokens); cudaFree(L.d_out_len);
    cudaFree(L.d_next4); cudaFree(L.d_prev3); cudaFree(L.d_order); cudaFree(L.d_clen); cudaFree(L.d_mtab2); cudaFree(L.d_lists); cudaFree(L.d_list_start); cudaFree(L.d_mtab); cudaFree(L.d_offsets); cudaFree(L.d_out); cudaFree(L.d_packed);
    cudaFree(L.d_status); cudaFree(L.d_overflow); cudaFree(L.d_comb); cudaFreeHost(L.h_comb);
    cudaFreeHost(L.h_in); cudaFreeHost(L.h_packed); cudaFreeHost(L.h_len); cudaFreeHost(L.h_dict); cudaFreeHost(L.h_flags); cudaFreeHost(L.h_crc);
    cudaFreeHost(L.h_offsets); cudaFreeHost(L.h_status); cudaFreeHost(L.h_overflow); cudaFreeHost(L.h_end);
    if (L.ev_scan) cudaEventDestroy(L.ev_scan);
    if (L.ev_done) cudaEventDestroy(L.ev_done);
    if (L.st) cudaStreamDestroy(L.st);
    L = Lane();
}

extern "C" int gzpb_create(gzpb_ctx **out, int device, int format, int level, size_t max_block_bytes,
                           size_t max_blocks_in_flight)
{
    if (!out) return GZPB_EINVAL;
    *out = nullptr;
    if (format < GZPB_GZIP || format > GZPB_SNAP) return GZPB_EINVAL;
    if (!gzpb_level_supported(format, level)) return GZPB_ELEVEL;
    if (max_block_bytes == 0) max_block_bytes = gzpb_default_bufsize(format);
    if (max_blocks_in_flight == 0) max_blocks_in_flight = 1024;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { snprintf(g_last_cuda_error, sizeof g_last_cuda_error, "no CUDA device %d", device); return GZPB_ECUDA; }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        snprintf(g_last_cuda_error, sizeof g_last_cuda_error, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return GZPB_ECUDA;
    }
    size_t dict = gzpb_needs_dict(format) ? GZPB_DICT_SIZE : 0;
    if (max_block_bytes > (1u << 22)) return GZPB_EBUFFERSIZE;
    gzpb_ctx *c = new gzpb_ctx();
    c->device = device; c->format = format; c->level = level;
    c->max_block_bytes = max_block_bytes; c->max_units = max_blocks_in_flight;
    { const char *e = getenv("GZPB_GATHER_CTAS"); if (e && atoi(e) > 0) c->gather_host_ctas = atoi(e); }
    {
        const size_t U = dict + max_block_bytes;
        c->dict_cap = (uint32_t)dict;
        if (U <= (size_t)kMaxUnitBytes) { c->spu = 1; c->seg = (uint32_t)U; c->in_stride = kInStride; }
        else {
            c->seg = 32256;
            c->spu = (uint32_t)((max_block_bytes + c->seg - 1) / c->seg);
            c->in_stride = (uint32_t)((U + 80 + 127) & ~(size_t)127);
        }
        c->m_stride = (uint32_t)(((U + 15) & ~(size_t)15) + 32);
        c->tok_stride = (uint32_t)(U + 64);
        c->out_stride = (uint32_t)((kOutPayloadOff + gzpb_encode_capacity(format, max_block_bytes) + 64 + 127) & ~(size_t)127);
        c->check_kind = (format == GZPB_ZLIB) ? 1 : (format == GZPB_RAWDEFLATE ? -1 : 0);
        if (format == GZPB_SNAP) {
            c->cpu = (uint32_t)((max_block_bytes + 65535) / 65536);
            if (c->cpu == 0) c->cpu = 1;
            c->in_stride = (uint32_t)((max_block_bytes + 80 + 127) & ~(size_t)127);
            c->out_stride = 76544;   // 32 + max_compress_len(65536) rounded up: one slot per 64 KiB chunk
            c->spu = 1;
        }
    }
    upload_snappy_constants();
    upload_deflate_constants();
    for (int i = 0; i < kLanes; i++) {
        int r = lane_alloc(c, c->lanes[i], true);
        if (r != GZPB_OK) { gzpb_destroy(c); return r; }
    }
    if (cudaMalloc((void **)&c->d_base0, sizeof(uint64_t)) != cudaSuccess || hmalloc(&c->h_base0, 2) != cudaSuccess) { gzpb_destroy(c); return GZPB_ENOMEM; }
    CK(cudaDeviceSynchronize());
    *out = c;
    return GZPB_OK;
}

extern "C" void gzpb_destroy(gzpb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kLanes; i++) lane_free(c->lanes[i]);
    cudaFree(c->d_base0); cudaFreeHost(c->h_base0);
    c->timer.collect();
    for (auto e : c->timer.pool) cudaEventDestroy(e);
    delete c;
}

extern "C" int gzpb_set_profiling(gzpb_ctx *c, int on)
{
    if (!c) return GZPB_EINVAL;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->timer.reset();
    c->profiling = on != 0;
    return GZPB_OK;
}

extern "C" int gzpb_kernel_ms(gzpb_ctx *c, const char *name, double *total_ms, uint64_t *launches)
{
    static const char *names[KT_COUNT] = {"chain", "match", "emit", "gather", "crc", "snap", "inflate"};
    if (!c || !name) return GZPB_EINVAL;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->timer.collect();
    for (int i = 0; i < KT_COUNT; i++)
        if (!strcmp(name, names[i])) {
            if (total_ms) *total_ms = c->timer.total_ms[i];
            if (launches) *launches = c->timer.launches[i];
            return GZPB_OK;
        }
    return GZPB_EINVAL;
}

extern "C" uint64_t gzpb_launch_count(gzpb_ctx *c) { return c ? c->launches : 0; }

extern "C" const char *gzpb_ctx_variant(gzpb_ctx *c)
{
    if (!c) return "";
    return c->format == GZPB_SNAP ? "snap" : "split+link+match";
}

/* debugging aid (not part of the reference-facing surface): SM-cycle totals per kernel phase */
extern "C" int gzpb_debug_phase_cycles(gzpb_ctx *c, uint64_t *out32, int reset)
{
    if (!c || !out32) return GZPB_EINVAL;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    read_phase_counters((unsigned long long *)out32, reset != 0);
    return GZPB_OK;
}

static uint32_t unit_flags_for(int format, int is_last)
{
    switch (format) {
    case GZPB_BGZF: return is_last ? 1u : 0u;
    case GZPB_MGZIP: return 0u;
    case GZPB_RAWDEFLATE: return 2u;                      // always Z_SYNC_FLUSH (deflate.rs:319-320)
    default: return is_last ? 0u : 2u;                    // Gzip / Zlib: Sync unless last (deflate.rs:96-100)
    }
}

static void fill_batch(gzpb_ctx *c, Lane &L, DeflateBatch &b, size_t n)
{
    b.nunits = (uint32_t)n; b.level = c->level; b.format = c->format;
    b.in = L.d_in; b.unit_len = L.d_len; b.unit_dict = L.d_dict; b.unit_flags = L.d_flags;
    b.in_stride = c->in_stride; b.m_stride = c->m_stride; b.tok_stride = c->tok_stride; b.out_stride = c->out_stride;
    b.spu = c->spu; b.seg = c->seg; b.check_kind = c->check_kind;
    b.next4 = L.d_next4; b.prev3 = L.d_prev3; b.clen = L.d_clen; b.order = L.d_order; b.mtab = L.d_mtab; b.mtab2 = L.d_mtab2; b.lists = L.d_lists; b.list_start = L.d_list_start; b.crc = L.d_crc; b.sum_part = getenv("GZPB_SEPARATE_CHECK") ? nullptr : L.d_sum_part; b.tokens = L.d_tokens;
    b.out = L.d_out; b.out_len = L.d_out_len; b.status = L.d_status; b.offsets = L.d_offsets;
    b.packed = nullptr; b.packed_cap = 0; b.packed_on_host = 0; b.base_ptr = nullptr; b.end_mirror = nullptr; b.overflow = L.d_overflow;
    b.timer = c->profiling ? &c->timer : nullptr;
    b.launch_counter = &c->launches;
}

static bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

extern "C" size_t gzpb_unit_stride(gzpb_ctx *c) { return c ? c->in_stride : 0; }

// Device-resident form for every format.  d_in: nunits slots of gzpb_unit_stride(ctx) bytes, each [dictionary | data];
// d_len[i] = dictionary + data bytes, d_dict[i] = dictionary bytes (NULL = none), d_flags as gzpb_encode_device.
// Snap: d_offsets / the packed stream have one entry per 64 KiB chunk (ceil(max_block_bytes / 65536) per unit).
extern "C" int gzpb_encode_device_ex(gzpb_ctx *c, const void *d_in, const uint32_t *d_len, const uint32_t *d_dict,
                                     const uint32_t *d_flags, size_t nunits, void *d_packed, uint64_t *d_offsets,
                                     int32_t *d_status, void *cuda_stream)
{
    if (!c || !d_in || !d_len || !d_packed || !d_offsets) return GZPB_EINVAL;
    if (c->format != GZPB_SNAP && (!d_flags || !d_status)) return GZPB_EINVAL;
    if (c->dict_cap == 0 && d_dict) return GZPB_EINVAL;
    CK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Lane &L = c->lanes[0];
    CK(cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), st));
    for (size_t done = 0; done < nunits; done += c->max_units) {
        size_t n = std::min(c->max_units, nunits - done);
        DeflateBatch b;
        fill_batch(c, L, b, n);
        const uint8_t *in = (const uint8_t *)d_in + done * (size_t)c->in_stride;
        if (c->format == GZPB_SNAP) {
            SnapBatch sb;
            sb.nunits = (uint32_t)n; sb.cpu = c->cpu; sb.in = in; sb.unit_len = d_len + done; sb.in_stride = c->in_stride;
            sb.out = L.d_out; sb.out_stride = c->out_stride; sb.out_len = L.d_out_len; sb.timer = c->profiling ? &c->timer : nullptr;
            CK(launch_snap(sb, st));
            b.nunits = (uint32_t)(n * c->cpu);
            b.offsets = d_offsets + done * c->cpu; b.base_ptr = d_offsets + done * c->cpu;
            c->launches += 3;   // k_snap, k_scan, k_gather
        } else {
            b.in = in; b.unit_len = d_len + done; b.unit_flags = d_flags + done;
            if (d_dict) b.unit_dict = d_dict + done;
            b.status = d_status + done; b.offsets = d_offsets + done; b.base_ptr = d_offsets + done;
            CK(launch_deflate_pipeline(b, st));
            c->launches += 2;   // k_scan, k_gather (the pipeline's kernels count themselves: DeflateBatch::launch_counter)
        }
        b.packed = (uint8_t *)d_packed; b.packed_cap = ~0ull;
        CK(launch_pack(b, st));
    }
    return GZPB_OK;
}

extern "C" int gzpb_encode_device(gzpb_ctx *c, const void *d_in, const uint32_t *d_len, const uint32_t *d_flags,
                                  size_t nunits, void *d_packed, uint64_t *d_offsets, int32_t *d_status,
                                  void *cuda_stream)
{
    if (!c || !is_deflate_format(c->format) || c->dict_cap) return GZPB_EINVAL;   // formats without dictionary
    return gzpb_encode_device_ex(c, d_in, d_len, nullptr, d_flags, nunits, d_packed, d_offsets, d_status, cuda_stream);
}

// ---- host-buffer paths ---------------------------------------------------------
namespace {
struct UnitRef { const uint8_t *ptr; size_t len; const uint8_t *dict; size_t dict_len; int is_last; };
}

// Launch one batch on a lane: H2D + kernels.  `src_pinned`: every unit (and dictionary) lies in pinned host
// memory, so the units go H2D straight from where they are — maximal runs of equally sized units at a
// constant pitch (consecutive blocks of one buffer) as ONE strided DMA, everything else unit by unit.
// Pageable sources are packed into the lane's pinned staging buffer first.
static int lane_launch(gzpb_ctx *c, Lane &L, const UnitRef *units, size_t n, bool src_pinned)
{
    const size_t IS = c->in_stride;
    for (size_t i = 0; i < n; i++) {
        L.h_len[i] = (uint32_t)(units[i].dict_len + units[i].len);
        L.h_dict[i] = (uint32_t)units[i].dict_len;
        L.h_flags[i] = unit_flags_for(c->format, units[i].is_last);
    }
    if (src_pinned) {
        auto adjacent = [](const UnitRef &u) { return u.dict_len == 0 || u.dict + u.dict_len == u.ptr; };
        size_t i = 0;
        while (i < n) {
            const UnitRef &u = units[i];
            const bool adj = adjacent(u);
            size_t j = i + 1;
            size_t P = 0;
            if (adj && u.len > 0 && j < n && (uintptr_t)units[j].ptr > (uintptr_t)u.ptr) {
                P = (size_t)((uintptr_t)units[j].ptr - (uintptr_t)u.ptr);
                while (j < n && units[j].len == u.len && units[j].dict_len == u.dict_len && adjacent(units[j]) &&
                       (uintptr_t)units[j].ptr - (uintptr_t)units[j - 1].ptr == P)
                    j++;
            }
            if (j - i >= 2 && P >= u.len && P >= u.dict_len) {
                // rows of a 2-D copy may not overlap: dictionary columns and data columns go separately
                const size_t rows = j - i;
                if (u.dict_len) CK(cudaMemcpy2DAsync(L.d_in + i * IS, IS, u.ptr - u.dict_len, P, u.dict_len, rows, cudaMemcpyHostToDevice, L.st));
                CK(cudaMemcpy2DAsync(L.d_in + i * IS + u.dict_len, IS, u.ptr, P, u.len, rows, cudaMemcpyHostToDevice, L.st));
            } else {
                j = i + 1;
                if (adj) {
                    if (u.dict_len + u.len) CK(cudaMemcpyAsync(L.d_in + i * IS, u.ptr - u.dict_len, u.dict_len + u.len, cudaMemcpyHostToDevice, L.st));
                } else {
                    CK(cudaMemcpyAsync(L.d_in + i * IS, u.dict, u.dict_len, cudaMemcpyHostToDevice, L.st));
                    if (u.len) CK(cudaMemcpyAsync(L.d_in + i * IS + u.dict_len, u.ptr, u.len, cudaMemcpyHostToDevice, L.st));
                }
            }
            i = j;
        }
    } else {
        if (!L.h_in) CK(cudaHostAlloc((void **)&L.h_in, c->max_units * IS, cudaHostAllocPortable));
        size_t used = 0;
        for (size_t i = 0; i < n; i++) {
            uint8_t *dst = L.h_in + i * IS;
            if (units[i].dict_len) memcpy(dst, units[i].dict, units[i].dict_len);
            if (units[i].len) memcpy(dst + units[i].dict_len, units[i].ptr, units[i].len);
            used = (i + 1) * IS;
        }
        if (used) CK(cudaMemcpyAsync(L.d_in, L.h_in, used, cudaMemcpyHostToDevice, L.st));
    }
    CK(cudaMemcpyAsync(L.d_len, L.h_len, n * sizeof(uint32_t), cudaMemcpyHostToDevice, L.st));
    if (c->format == GZPB_SNAP) {
        SnapBatch sb;
        sb.nunits = (uint32_t)n; sb.cpu = c->cpu; sb.in = L.d_in; sb.unit_len = L.d_len; sb.in_stride = c->in_stride;
        sb.out = L.d_out; sb.out_stride = c->out_stride; sb.out_len = L.d_out_len; sb.timer = c->profiling ? &c->timer : nullptr;
        CK(launch_snap(sb, L.st));
        c->launches += 1;
        L.nunits = n; L.busy = true;
        return GZPB_OK;
    }
    CK(cudaMemcpyAsync(L.d_dict, L.h_dict, n * sizeof(uint32_t), cudaMemcpyHostToDevice, L.st));
    CK(cudaMemcpyAsync(L.d_flags, L.h_flags, n * sizeof(uint32_t), cudaMemcpyHostToDevice, L.st));
    DeflateBatch b;
    fill_batch(c, L, b, n);
    CK(launch_deflate_pipeline(b, L.st));
    L.nunits = n; L.busy = true;
    return GZPB_OK;
}

// The lane's own pinned output buffer (ticket / writer / pageable-output paths).  Allocated on first use: the
// zero-copy stream path gathers straight into the caller's pinned buffer and never needs it.
static int lane_host_packed(gzpb_ctx *c, Lane &L)
{
    if (!L.h_packed) CK(hmalloc(&L.h_packed, c->max_units * c->cpu * (size_t)c->out_stride));
    return GZPB_OK;
}

// scan + gather into `packed` (device-visible), chained after `prev` lane's scan (`prev` may belong to another
// device: the event wait is cross-device and `base_ptr` then points into mapped pinned host memory, where the
// previous lane's k_scan mirrored its end offset through `end_mirror`)
static int lane_pack(gzpb_ctx *c, Lane &L, uint8_t *packed, uint64_t cap, const uint64_t *base_ptr, Lane *prev, uint64_t *end_mirror = nullptr)
{
    DeflateBatch b;
    fill_batch(c, L, b, L.nunits);
    const size_t entries = L.nunits * c->cpu;
    b.nunits = (uint32_t)entries;
    b.packed = packed; b.packed_cap = cap; b.base_ptr = base_ptr; b.end_mirror = end_mirror;
    b.packed_on_host = c->gather_host_ctas;          // lane_pack always gathers into pinned host memory (zero-copy D2H)
    if (prev) CK(cudaStreamWaitEvent(L.st, prev->ev_scan, 0));
    CK(launch_pack(b, L.st));
    CK(cudaEventRecord(L.ev_scan, L.st));
    c->launches += 2;
    CK(cudaMemcpyAsync(L.h_offsets, L.d_offsets, (entries + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, L.st));
    if (c->format != GZPB_SNAP) {
        CK(cudaMemcpyAsync(L.h_status, L.d_status, L.nunits * sizeof(int32_t), cudaMemcpyDeviceToHost, L.st));
        CK(cudaMemcpyAsync(L.h_crc, L.d_crc, L.nunits * sizeof(uint32_t), cudaMemcpyDeviceToHost, L.st));
        if (c->format == GZPB_GZIP || c->format == GZPB_ZLIB) {
            // Check::combine over the batch on the device: the host folds one value per batch
            CK(launch_check_combine(L.d_crc, L.d_len, L.d_dict, (uint32_t)L.nunits, c->check_kind, L.d_comb, L.st));
            CK(cudaMemcpyAsync(L.h_comb, L.d_comb, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, L.st));
            c->launches += 1;
        }
    }
    CK(cudaMemcpyAsync(L.h_overflow, L.d_overflow, sizeof(int32_t), cudaMemcpyDeviceToHost, L.st));
    CK(cudaEventRecord(L.ev_done, L.st));
    return GZPB_OK;
}

static int lane_wait(gzpb_ctx *c, Lane &L)
{
    CK(cudaEventSynchronize(L.ev_done));
    if (c->profiling) c->timer.collect();
    L.busy = false;
    return GZPB_OK;
}

// completes the oldest ticket: wait for its lane, scatter the compacted blocks into the caller's buffers
static int ticket_retire(gzpb_ctx *c)
{
    const gzpb_ctx::Ticket t = c->tickets.front();
    c->tickets.pop_front();
    Lane &L = c->lanes[t.lane];
    int r = lane_wait(c, L);
    if (r == GZPB_OK && *L.h_overflow) r = GZPB_ECOMPRESS;
    if (r != GZPB_OK) {
        // the batch is lost as a whole: the lane is free again (after the device went idle) and every block says why
        cudaStreamSynchronize(L.st);
        L.busy = false;
        for (size_t i = 0; i < t.count; i++) { t.out[i].status = r; t.out[i].out_len = 0; t.out[i].check_sum = 0; t.out[i].check_amount = 0; }
        return r;
    }
    for (size_t i = 0; i < t.count; i++) {
        gzpb_block_out &o = t.out[i];
        o.status = L.h_status[i];
        const size_t o0 = (size_t)L.h_offsets[i * c->cpu], o1 = (size_t)L.h_offsets[(i + 1) * c->cpu];
        const size_t len = o1 - o0;
        o.out_len = 0; o.check_sum = 0; o.check_amount = 0;
        if (o.status == GZPB_OK) {
            if (len > o.cap) { o.status = GZPB_ECOMPRESS; continue; }
            memcpy(o.dst, L.h_packed + o0, len);
            o.out_len = len;
            if (c->format == GZPB_GZIP || c->format == GZPB_ZLIB) { o.check_sum = L.h_crc[i]; o.check_amount = (uint32_t)t.in[i].len; }
        }
    }
    return GZPB_OK;
}

extern "C" int gzpb_submit(gzpb_ctx *c, size_t n, const gzpb_block_in *in, gzpb_block_out *out, gzpb_ticket *ticket)
{
    if (!c || !ticket || n == 0 || !in || !out) return GZPB_EINVAL;
    if (n > c->max_units) return GZPB_EINVAL;
    CK(cudaSetDevice(c->device));
    const bool dict_fmt = gzpb_needs_dict(c->format);
    for (size_t i = 0; i < n; i++) {
        size_t dl = (dict_fmt && in[i].dict) ? in[i].dict_len : 0;
        if (dl > c->dict_cap || in[i].len > c->max_block_bytes) return GZPB_EBUFFERSIZE;
        if (out[i].cap < gzpb_encode_capacity(c->format, in[i].len)) return GZPB_EINVAL;
    }
    Lane &L = c->lanes[c->next_lane];
    if (L.busy) return GZPB_EAGAIN;                         // all lanes in flight: the bounded channel is full
    std::vector<UnitRef> units(n);
    bool pinned = true;
    for (size_t i = 0; i < n; i++) {
        const gzpb_block_in &b = in[i];
        size_t dl = (dict_fmt && b.dict) ? b.dict_len : 0;
        units[i] = UnitRef{(const uint8_t *)b.ptr, b.len, (const uint8_t *)b.dict, dl, b.is_last};
        if (pinned && ((b.len && !is_pinned(b.ptr)) || (dl && !is_pinned(b.dict)))) pinned = false;
    }
    int rc = lane_host_packed(c, L);
    if (rc != GZPB_OK) return rc;
    rc = lane_launch(c, L, units.data(), n, pinned);
    if (rc != GZPB_OK) return rc;
    rc = lane_pack(c, L, L.h_packed, (uint64_t)c->max_units * c->cpu * c->out_stride, nullptr, nullptr);
    if (rc != GZPB_OK) return rc;
    *ticket = c->next_ticket++;
    c->tickets.push_back(gzpb_ctx::Ticket{*ticket, in, out, n, c->next_lane});
    c->next_lane = (c->next_lane + 1) % kLanes;
    return GZPB_OK;
}

extern "C" int gzpb_poll(gzpb_ctx *c, gzpb_ticket ticket, int wait)
{
    if (!c || ticket == 0 || ticket >= c->next_ticket) return GZPB_EINVAL;
    CK(cudaSetDevice(c->device));
    while (!c->tickets.empty() && c->tickets.front().id <= ticket) {
        if (!wait) {
            cudaError_t q = cudaEventQuery(c->lanes[c->tickets.front().lane].ev_done);
            if (q == cudaErrorNotReady) return GZPB_EAGAIN;
            CK(q);
        }
        int rc = ticket_retire(c);
        if (rc != GZPB_OK) return rc;
    }
    return GZPB_OK;
}

extern "C" int gzpb_encode_batch(gzpb_ctx *c, size_t n, const gzpb_block_in *in, gzpb_block_out *out)
{
    if (!c || (n && (!in || !out))) return GZPB_EINVAL;
    if (!c->tickets.empty()) return GZPB_EAGAIN;            // finish the asynchronous tickets first
    size_t done = 0;
    gzpb_ticket last = 0;
    int rc = GZPB_OK;
    while (done < n && rc == GZPB_OK) {
        const size_t cnt = std::min(c->max_units, n - done);
        gzpb_ticket t = 0;
        rc = gzpb_submit(c, cnt, in + done, out + done, &t);
        if (rc == GZPB_EAGAIN) { rc = gzpb_poll(c, c->tickets.front().id, 1); continue; }
        if (rc == GZPB_OK) { last = t; done += cnt; }
    }
    if (last) { int r = gzpb_poll(c, last, 1); if (rc == GZPB_OK) rc = r; }
    if (rc != GZPB_OK) {                                    // leave no batch half-finished behind an error
        cudaDeviceSynchronize();
        c->tickets.clear();
        for (int i = 0; i < kLanes; i++) c->lanes[i].busy = false;
    }
    return rc;
}

// ParCompress end to end over an in-memory input, on one or several GPUs.  Device batches of max_units
// consecutive blocks are dealt round-robin: batch k runs on context k % G, lane (k / G) % kLanes (SURVEY §8e —
// blocks are independent; the dictionary of a batch's first block comes from the host input, so there is no
// device-to-device traffic).  With pinned `out` the ordered writer (par/compress.rs:303-313) is the offset
// chain: every batch's k_scan waits for the previous batch's k_scan (an event wait, cross-device when G > 1),
// takes its first stream offset from where that one ended and k_gather writes the blocks at their final
// stream position in the caller's buffer.  One host thread drives all devices: per batch it issues ~25
// asynchronous calls and the retire reads a few status words.
static int encode_stream_impl(gzpb_ctx *const *cs, size_t G, const void *in_v, size_t in_len, size_t buffer_size, void *out_v,
                              size_t out_cap, size_t *out_len)
{
    gzpb_ctx *c0 = cs[0];
    if (buffer_size == 0) buffer_size = c0->max_block_bytes;
    if (buffer_size < GZPB_DICT_SIZE) return GZPB_EBUFFERSIZE;   // par/compress.rs:68-74
    for (size_t g = 0; g < G; g++) {
        gzpb_ctx *c = cs[g];
        if (!c || c->format != c0->format || c->level != c0->level || c->max_units != c0->max_units ||
            c->max_block_bytes != c0->max_block_bytes)
            return GZPB_EINVAL;
        if (buffer_size > c->max_block_bytes) return GZPB_EBUFFERSIZE;
        if (!c->tickets.empty()) return GZPB_EAGAIN;
        for (size_t h = 0; h < g; h++) if (cs[h] == c) return GZPB_EINVAL;
    }
    const int format = c0->format;
    const uint32_t cpu = c0->cpu;
    const uint8_t *in = (const uint8_t *)in_v;
    uint8_t *out = (uint8_t *)out_v;
    const bool dict_fmt = gzpb_needs_dict(format);
    const bool in_pinned = in_len && is_pinned(in);
    const bool out_pinned = is_pinned(out);

    // ParCompress::write + finish: full blocks while MORE than buffer_size bytes remain,
    // then flush_last(true) — always at least one (possibly empty) is_last block.
    size_t nblocks = (in_len > buffer_size) ? (in_len - 1) / buffer_size + 1 : 1;

    if (out_cap < 64) return GZPB_ECOMPRESS;
    size_t pos_out = gzpb_header(format, c0->level, out);
    uint8_t *dev_out = nullptr;
    CK(cudaSetDevice(c0->device));
    if (out_pinned) CK(cudaHostGetDevicePointer((void **)&dev_out, out, 0));

    // first stream offset: a device word for one GPU, a mapped pinned word (visible to every GPU) for several
    const uint64_t base0 = out_pinned ? pos_out : 0;
    const uint64_t *first_base;
    if (G == 1) { CK(cudaMemcpy(c0->d_base0, &base0, sizeof base0, cudaMemcpyHostToDevice)); first_base = c0->d_base0; }
    else { *c0->h_base0 = base0; first_base = c0->h_base0; }

    struct Pending { size_t count; int dev, lane; };
    std::deque<Pending> pend;
    std::vector<UnitRef> units;
    uint32_t run_sum = (format == GZPB_ZLIB) ? 1u : 0u, run_amount = 0;
    int rc = GZPB_OK;
    Lane *prev = nullptr;
    const uint64_t *prev_end = first_base;

    auto retire = [&](const Pending &p) -> int {
        gzpb_ctx *c = cs[p.dev];
        Lane &L = c->lanes[p.lane];
        int r = lane_wait(c, L);
        if (r != GZPB_OK) return r;
        if (*L.h_overflow) return GZPB_ECOMPRESS;
        if (format != GZPB_SNAP)
            for (size_t i = 0; i < p.count; i++)
                if (L.h_status[i] != GZPB_OK) return L.h_status[i];
        if (format == GZPB_GZIP || format == GZPB_ZLIB) {
            // k_check_combine folded the batch; fold the batch into the running check (par/compress.rs:308)
            const uint64_t blen = (uint64_t)L.h_comb[1] | ((uint64_t)L.h_comb[2] << 32);
            if (blen) run_sum = format == GZPB_GZIP ? gzpb_crc32_combine(run_sum, L.h_comb[0], blen)
                                                    : gzpb_adler32_combine(run_sum, L.h_comb[0], blen);
            run_amount += (uint32_t)blen;
        }
        if (out_pinned) {
            pos_out = (size_t)L.h_offsets[p.count * cpu];
        } else {
            size_t len = (size_t)L.h_offsets[p.count * cpu];
            if (pos_out + len > out_cap) return GZPB_ECOMPRESS;
            memcpy(out + pos_out, L.h_packed, len);
            pos_out += len;
        }
        return GZPB_OK;
    };

    size_t done = 0, k = 0;
    while (done < nblocks && rc == GZPB_OK) {
        const int dev = (int)(k % G), li = (int)((k / G) % kLanes);
        gzpb_ctx *c = cs[dev];
        const size_t cnt = std::min(c->max_units, nblocks - done);
        Lane &L = c->lanes[li];
        while (L.busy && rc == GZPB_OK) { rc = retire(pend.front()); pend.pop_front(); }   // in-order drain up to this lane's batch
        if (rc != GZPB_OK) break;
        if (cudaSetDevice(c->device) != cudaSuccess) { rc = GZPB_ECUDA; break; }
        units.resize(cnt);
        for (size_t i = 0; i < cnt; i++) {
            size_t bi = done + i, b0 = bi * buffer_size;
            size_t len = (bi + 1 == nblocks) ? in_len - b0 : buffer_size;
            const uint8_t *d = (dict_fmt && bi > 0) ? in + b0 - GZPB_DICT_SIZE : nullptr;
            units[i] = UnitRef{in + b0, len, d, d ? (size_t)GZPB_DICT_SIZE : 0, bi + 1 == nblocks};
        }
        rc = lane_launch(c, L, units.data(), cnt, in_pinned);
        if (rc != GZPB_OK) break;
        if (out_pinned) rc = lane_pack(c, L, dev_out, out_cap, prev_end, prev, G > 1 ? L.h_end : nullptr);
        else { rc = lane_host_packed(c, L); if (rc == GZPB_OK) rc = lane_pack(c, L, L.h_packed, (uint64_t)c->max_units * cpu * c->out_stride, nullptr, nullptr); }
        if (rc != GZPB_OK) break;
        prev = &L; prev_end = (G > 1) ? L.h_end : L.d_offsets + cnt * cpu;
        pend.push_back(Pending{cnt, dev, li});
        done += cnt; k++;
    }
    while (!pend.empty()) {
        int r = retire(pend.front()); pend.pop_front();
        if (rc == GZPB_OK) rc = r;
    }
    if (rc != GZPB_OK) {                                          // leave no batch half-finished behind an error
        for (size_t g = 0; g < G; g++) {
            cudaSetDevice(cs[g]->device); cudaDeviceSynchronize();
            for (int i = 0; i < kLanes; i++) cs[g]->lanes[i].busy = false;
        }
        return rc;
    }
    uint8_t foot[16];
    size_t fl = gzpb_footer(format, run_sum, run_amount, foot);
    if (pos_out + fl > out_cap) return GZPB_ECOMPRESS;
    memcpy(out + pos_out, foot, fl);
    pos_out += fl;
    *out_len = pos_out;
    return GZPB_OK;
}

extern "C" int gzpb_encode_stream(gzpb_ctx *c, const void *in_v, size_t in_len, size_t buffer_size, void *out_v,
                                  size_t out_cap, size_t *out_len)
{
    if (!c || !out_v || !out_len || (in_len && !in_v)) return GZPB_EINVAL;
    return encode_stream_impl(&c, 1, in_v, in_len, buffer_size, out_v, out_cap, out_len);
}

extern "C" int gzpb_encode_stream_multi(gzpb_ctx *const *ctxs, size_t nctx, const void *in_v, size_t in_len, size_t buffer_size,
                                        void *out_v, size_t out_cap, size_t *out_len)
{
    if (!ctxs || nctx == 0 || nctx > 64 || !out_v || !out_len || (in_len && !in_v)) return GZPB_EINVAL;
    return encode_stream_impl(ctxs, nctx, in_v, in_len, buffer_size, out_v, out_cap, out_len);
}


// ---- incremental writer: ParCompress<F, W> as a C object ---------------------------------
//
// The caller's bytes are copied ONCE, straight into a pinned slab (the reference's single
// `buffer.extend_from_slice`, par/compress.rs:414); blocks are cut in place, so a batch is a run of
// consecutive slices of the slab and goes H2D as one strided DMA.  Batches are dealt round-robin over
// the writer's devices and their lanes and stay in flight while the caller keeps writing;
// they retire strictly in submission order — the ticket FIFO of par/compress.rs:303-313 — each with ONE
// sink call on the lane's compacted pinned output.
// Helper threads for the one host copy of ParCompress::write (par/compress.rs:414).  A single core copies
// ~10 GB/s, one B200 compresses ~9 GiB/s and a box has eight: large writes are cut into 2 MiB pieces that
// the caller and `n - 1` helpers copy side by side into the pinned slab.  Pure memcpy — no CUDA calls.
namespace {
struct CopyPool {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv, cv_done;
    const uint8_t *src = nullptr;
    uint8_t *dst = nullptr;
    size_t len = 0;
    std::atomic<size_t> next{0};
    uint64_t gen = 0;
    int working = 0;
    bool stop = false;
    static constexpr size_t kPiece = 2u << 20;
    explicit CopyPool(int helpers)
    {
        for (int i = 0; i < helpers; i++) th.emplace_back([this] { loop(); });
    }
    ~CopyPool()
    {
        { std::lock_guard<std::mutex> g(m); stop = true; }
        cv.notify_all();
        for (auto &t : th) t.join();
    }
    void pieces()
    {
        for (;;) {
            const size_t o = next.fetch_add(kPiece);
            if (o >= len) return;
            memcpy(dst + o, src + o, std::min(kPiece, len - o));
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> g(m);
        for (;;) {
            cv.wait(g, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen;
            g.unlock();
            pieces();
            g.lock();
            if (--working == 0) cv_done.notify_one();
        }
    }
    void copy(uint8_t *d, const uint8_t *s, size_t n)
    {
        {
            std::lock_guard<std::mutex> g(m);
            dst = d; src = s; len = n; next.store(0);
            working = (int)th.size();
            gen++;
        }
        cv.notify_all();
        pieces();
        std::unique_lock<std::mutex> g(m);
        cv_done.wait(g, [&] { return working == 0; });
    }
};
}  // namespace

struct gzpb_writer {
    std::vector<gzpb_ctx *> ctx;                                 // one per device
    CopyPool *pool = nullptr;                                    // gzpb_writer_set_copy_threads
    int format = 0, level = 0;
    size_t buffer_size = 0, batch_blocks = 0;
    gzpb_sink_fn sink = nullptr;
    void *user = nullptr;
    // slabs: [32 KiB dictionary prefix | (batch_blocks + 1) * buffer_size bytes of stream]
    std::vector<uint8_t *> slabs;
    size_t slab_cap = 0;                                         // stream bytes per slab
    size_t cur = 0;                                              // slab being filled
    size_t fill = 0, cut = 0;                                    // bytes written / bytes already cut into blocks
    std::vector<UnitRef> msgs;                                   // blocks cut from the current slab (FIFO = ticket order)
    bool have_dict = false;                                      // the next block takes the 32 KiB before `cut` as dictionary
    struct Flight { int dev, lane; size_t count; };
    std::deque<Flight> flights;                                  // batches in flight, oldest first
    uint64_t nbatches = 0;
    bool wrote_header = false, finished = false;
    uint32_t sum = 0, amount = 0;
    int error = GZPB_OK;
    uint64_t sink_calls = 0, bytes_in = 0, bytes_out = 0;
    // BGZF only: the .gzi entries, a by-product of knowing every block's compressed size when its batch retires
    std::vector<uint64_t> gzi;                                   // {compressed offset, uncompressed offset} pairs
    uint64_t gzi_upos = 0;
    bool gzi_first = true;
};

static int writer_emit(gzpb_writer *w, const void *p, size_t n)
{
    if (n == 0) return GZPB_OK;
    w->sink_calls++; w->bytes_out += n;
    if (w->sink(w->user, p, n) != 0) { w->error = GZPB_EIO; return GZPB_EIO; }
    return GZPB_OK;
}

static int writer_header(gzpb_writer *w)
{
    if (w->wrote_header) return GZPB_OK;
    uint8_t hb[16];
    size_t hl = gzpb_header(w->format, w->level, hb);
    w->wrote_header = true;
    return writer_emit(w, hb, hl);
}

// the writer loop body (par/compress.rs:305-311) for the oldest batch in flight
static int writer_retire(gzpb_writer *w)
{
    const gzpb_writer::Flight f = w->flights.front();
    w->flights.pop_front();
    gzpb_ctx *c = w->ctx[f.dev];
    Lane &L = c->lanes[f.lane];
    if (cudaSetDevice(c->device) != cudaSuccess) return w->error = GZPB_ECUDA;
    int r = lane_wait(c, L);
    if (r != GZPB_OK) return w->error = r;
    if (*L.h_overflow) return w->error = GZPB_ECOMPRESS;
    if (c->format != GZPB_SNAP)
        for (size_t i = 0; i < f.count; i++)
            if (L.h_status[i] != GZPB_OK) return w->error = L.h_status[i];          // a failed block fails the stream
    if (w->format == GZPB_GZIP || w->format == GZPB_ZLIB) {
        const uint64_t blen = (uint64_t)L.h_comb[1] | ((uint64_t)L.h_comb[2] << 32);
        if (blen) w->sum = w->format == GZPB_GZIP ? gzpb_crc32_combine(w->sum, L.h_comb[0], blen)
                                                  : gzpb_adler32_combine(w->sum, L.h_comb[0], blen);
        w->amount += (uint32_t)blen;
    }
    if (w->format == GZPB_BGZF) {
        // same rule as gzpb_bgzf_index: one entry per data block after the first; empty blocks carry no data
        for (size_t i = 0; i < f.count; i++) {
            const uint32_t isize = L.h_len[i];
            if (!isize) continue;
            if (!w->gzi_first) { w->gzi.push_back(w->bytes_out + L.h_offsets[i]); w->gzi.push_back(w->gzi_upos); }
            w->gzi_first = false;
            w->gzi_upos += isize;
        }
    }
    return writer_emit(w, L.h_packed, (size_t)L.h_offsets[f.count * c->cpu]);
}

static int writer_drain(gzpb_writer *w)
{
    while (!w->flights.empty()) { int r = writer_retire(w); if (r != GZPB_OK) return r; }
    return GZPB_OK;
}

// hand the blocks cut so far to the next device lane and continue in the next slab
static int writer_submit(gzpb_writer *w)
{
    if (w->error) return w->error;
    int r = writer_header(w);
    if (r != GZPB_OK) return r;
    if (w->msgs.empty()) return GZPB_OK;
    const size_t G = w->ctx.size();
    const int dev = (int)(w->nbatches % G), lane = (int)((w->nbatches / G) % kLanes);
    gzpb_ctx *c = w->ctx[dev];
    Lane &L = c->lanes[lane];
    // batches that are already finished leave now (keeps the sink fed); a busy lane is back-pressure
    while (!w->flights.empty()) {
        const gzpb_writer::Flight &f = w->flights.front();
        if (!L.busy && cudaEventQuery(w->ctx[f.dev]->lanes[f.lane].ev_done) != cudaSuccess) break;
        r = writer_retire(w);
        if (r != GZPB_OK) return r;
    }
    cudaGetLastError();
    if (cudaSetDevice(c->device) != cudaSuccess) return w->error = GZPB_ECUDA;
    r = lane_host_packed(c, L);
    if (r == GZPB_OK) r = lane_launch(c, L, w->msgs.data(), w->msgs.size(), true);
    if (r == GZPB_OK) r = lane_pack(c, L, L.h_packed, (uint64_t)c->max_units * c->cpu * c->out_stride, nullptr, nullptr);
    if (r != GZPB_OK) return w->error = r;
    w->flights.push_back(gzpb_writer::Flight{dev, lane, w->msgs.size()});
    w->nbatches++;
    w->msgs.clear();
    // next slab: carry the dictionary (the 32 KiB before `cut`) and the bytes not yet cut
    uint8_t *from = w->slabs[w->cur];
    w->cur = (w->cur + 1) % w->slabs.size();
    uint8_t *to = w->slabs[w->cur];
    const size_t rem = w->fill - w->cut;
    if (w->have_dict) memcpy(to - GZPB_DICT_SIZE, from + w->cut - GZPB_DICT_SIZE, GZPB_DICT_SIZE + rem);
    else if (rem) memcpy(to, from + w->cut, rem);
    w->fill = rem; w->cut = 0;
    return GZPB_OK;
}

// cut [cut, cut + k) as one message (Message{buffer, dictionary, is_last}, lib.rs:282-312)
static int writer_cut(gzpb_writer *w, size_t k, bool is_last, bool next_has_dict)
{
    uint8_t *base = w->slabs[w->cur];
    UnitRef u{base + w->cut, k, nullptr, 0, is_last ? 1 : 0};
    if (w->have_dict) { u.dict = base + w->cut - GZPB_DICT_SIZE; u.dict_len = GZPB_DICT_SIZE; }
    w->msgs.push_back(u);
    w->cut += k;
    w->have_dict = next_has_dict;
    if (w->msgs.size() >= w->batch_blocks) return writer_submit(w);
    return GZPB_OK;
}

extern "C" int gzpb_writer_create_multi(gzpb_writer **out, const int *devices, size_t ndevices, int format, int level,
                                        size_t buffer_size, size_t blocks_in_flight, gzpb_sink_fn sink, void *user)
{
    if (!out || !sink || !devices || ndevices == 0 || ndevices > 64) return GZPB_EINVAL;
    *out = nullptr;
    if (buffer_size == 0) buffer_size = gzpb_default_bufsize(format);
    if (buffer_size < GZPB_DICT_SIZE) return GZPB_EBUFFERSIZE;   // par/compress.rs:68-74
    if (blocks_in_flight == 0) blocks_in_flight = 1184;          // 8 CTAs per SM on 148 SMs per batch
    gzpb_writer *w = new gzpb_writer();
    w->format = format; w->level = level; w->buffer_size = buffer_size; w->batch_blocks = blocks_in_flight;
    w->sink = sink; w->user = user;
    w->sum = (format == GZPB_ZLIB) ? 1u : 0u;
    for (size_t i = 0; i < ndevices; i++) {
        gzpb_ctx *c = nullptr;
        int rc = gzpb_create(&c, devices[i], format, level, buffer_size, blocks_in_flight);
        if (rc != GZPB_OK) { gzpb_writer_destroy(w); return rc; }
        w->ctx.push_back(c);
        for (int l = 0; l < kLanes; l++)                        // the writer's batches always land in the lanes' pinned output
            if ((rc = lane_host_packed(c, c->lanes[l])) != GZPB_OK) { gzpb_writer_destroy(w); return rc; }
    }
    w->slab_cap = (blocks_in_flight + 1) * buffer_size;
    const size_t nslabs = ndevices * kLanes + 1;                 // every batch in flight keeps its slab + the one being filled
    for (size_t i = 0; i < nslabs; i++) {
        uint8_t *p = nullptr;
        if (cudaHostAlloc((void **)&p, GZPB_DICT_SIZE + w->slab_cap, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            gzpb_writer_destroy(w);
            return GZPB_ENOMEM;
        }
        w->slabs.push_back(p + GZPB_DICT_SIZE);
    }
    *out = w;
    return GZPB_OK;
}

extern "C" int gzpb_writer_create(gzpb_writer **out, int device, int format, int level, size_t buffer_size,
                                  size_t blocks_in_flight, gzpb_sink_fn sink, void *user)
{
    return gzpb_writer_create_multi(out, &device, 1, format, level, buffer_size, blocks_in_flight, sink, user);
}

extern "C" int gzpb_writer_set_copy_threads(gzpb_writer *w, int nthreads)
{
    if (!w || nthreads < 1 || nthreads > 256) return GZPB_EINVAL;
    delete w->pool;
    w->pool = nthreads > 1 ? new CopyPool(nthreads - 1) : nullptr;
    return GZPB_OK;
}

extern "C" int gzpb_writer_reserve(gzpb_writer *w, void **ptr, size_t *room)
{
    if (!w || !ptr || !room) return GZPB_EINVAL;
    if (w->finished) return GZPB_ECHANNEL;
    if (w->error) return w->error;
    *ptr = w->slabs[w->cur] + w->fill;
    *room = w->slab_cap - w->fill;                                // never 0: at most buffer_size bytes are held back uncut
    return GZPB_OK;
}

extern "C" int gzpb_writer_commit(gzpb_writer *w, size_t n)
{
    if (!w || n > w->slab_cap - w->fill) return GZPB_EINVAL;
    if (w->finished) return GZPB_ECHANNEL;
    if (w->error) return w->error;
    const bool dict_fmt = gzpb_needs_dict(w->format) != 0;
    w->bytes_in += n;
    w->fill += n;
    while (w->fill - w->cut > w->buffer_size) {                   // strict '>' (par/compress.rs:415)
        int rc = writer_cut(w, w->buffer_size, false, dict_fmt);  // dictionary = last 32 KiB of this block (:419-423)
        if (rc != GZPB_OK) return rc;
    }
    return GZPB_OK;
}

extern "C" int gzpb_writer_write(gzpb_writer *w, const void *data, size_t len)
{
    if (!w || (len && !data)) return GZPB_EINVAL;
    if (w->finished) return GZPB_ECHANNEL;
    if (w->error) return w->error;
    const uint8_t *p = (const uint8_t *)data;
    while (len) {
        const size_t k = std::min(len, w->slab_cap - w->fill);
        if (w->pool && k >= 4 * CopyPool::kPiece) w->pool->copy(w->slabs[w->cur] + w->fill, p, k);
        else memcpy(w->slabs[w->cur] + w->fill, p, k);            // the one host copy (par/compress.rs:414)
        p += k; len -= k;
        int rc = gzpb_writer_commit(w, k);
        if (rc != GZPB_OK) return rc;
    }
    return GZPB_OK;
}

static int writer_flush_last(gzpb_writer *w, bool is_last)
{
    const bool dict_fmt = gzpb_needs_dict(w->format) != 0;
    for (;;) {                                                     // par/compress.rs:332-362
        const size_t k = std::min(w->fill - w->cut, w->buffer_size);
        const bool last = is_last && (w->cut + k == w->fill);
        int rc = writer_cut(w, k, last, k >= GZPB_DICT_SIZE && !last && dict_fmt);   // :343-345
        if (rc != GZPB_OK) return rc;
        if (w->cut == w->fill) break;
    }
    return writer_submit(w);
}

extern "C" int gzpb_writer_flush(gzpb_writer *w)
{
    if (!w) return GZPB_EINVAL;
    if (w->finished) return GZPB_ECHANNEL;
    if (w->error) return w->error;
    int rc = writer_flush_last(w, false);
    if (rc != GZPB_OK) return rc;
    return writer_drain(w);
}

extern "C" int gzpb_writer_finish(gzpb_writer *w)
{
    if (!w) return GZPB_EINVAL;
    if (w->finished) return GZPB_ECHANNEL;
    int rc = w->error;
    if (rc == GZPB_OK) rc = writer_flush_last(w, true);
    if (rc == GZPB_OK) rc = writer_drain(w);
    if (rc == GZPB_OK) {
        uint8_t fb[16];
        size_t fl = gzpb_footer(w->format, w->sum, w->amount, fb);
        rc = writer_emit(w, fb, fl);
    }
    w->finished = true;
    return rc;
}

extern "C" int gzpb_writer_stats(gzpb_writer *w, uint64_t *bytes_in, uint64_t *bytes_out, uint64_t *batches, uint64_t *sink_calls)
{
    if (!w) return GZPB_EINVAL;
    if (bytes_in) *bytes_in = w->bytes_in;
    if (bytes_out) *bytes_out = w->bytes_out;
    if (batches) *batches = w->nbatches;
    if (sink_calls) *sink_calls = w->sink_calls;
    return GZPB_OK;
}

extern "C" int gzpb_writer_bgzf_index(gzpb_writer *w, void *out_v, size_t out_cap, size_t *out_len)
{
    if (!w || !out_len) return GZPB_EINVAL;
    if (w->format != GZPB_BGZF) return GZPB_EINVAL;
    const size_t n = w->gzi.size() / 2, need = 8 + 16 * n;
    *out_len = need;
    if (!out_v) return GZPB_OK;
    if (out_cap < need) return GZPB_ECOMPRESS;
    uint8_t *out = (uint8_t *)out_v;
    auto put64 = [&](size_t at, uint64_t v) { for (int i = 0; i < 8; i++) out[at + i] = (uint8_t)(v >> (8 * i)); };
    put64(0, n);
    for (size_t i = 0; i < 2 * n; i++) put64(8 + 8 * i, w->gzi[i]);
    return GZPB_OK;
}

extern "C" void gzpb_writer_destroy(gzpb_writer *w)
{
    if (!w) return;
    for (gzpb_ctx *c : w->ctx) {                                   // batches still in flight read the slabs
        if (!c) continue;
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        for (int i = 0; i < kLanes; i++) c->lanes[i].busy = false;
    }
    for (uint8_t *p : w->slabs) cudaFreeHost(p - GZPB_DICT_SIZE);
    for (gzpb_ctx *c : w->ctx) gzpb_destroy(c);
    delete w->pool;
    delete w;
}

// ---- file ingest / egress around the writer --------------------------------
// read(2) lands in the writer's pinned slab at its fill position (no intermediate buffer, no second
// host copy), the ordered device output goes to the output file with write(2) from pinned memory.
#ifndef GZPB_EMU_NO_FILES
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

static int file_sink(void *user, const void *data, size_t len)
{
    const int fd = *(const int *)user;
    const uint8_t *p = (const uint8_t *)data;
    while (len) {
        ssize_t k = write(fd, p, len);
        if (k < 0) { if (errno == EINTR) continue; return 1; }
        p += k; len -= (size_t)k;
    }
    return 0;
}

extern "C" int gzpb_compress_file(const int *devices, size_t ndevices, int format, int level, size_t buffer_size,
                                  size_t blocks_in_flight, const char *in_path, const char *out_path,
                                  uint64_t *bytes_in, uint64_t *bytes_out)
{
    if (!in_path || !out_path) return GZPB_EINVAL;
    const int fin = open(in_path, O_RDONLY);
    if (fin < 0) return GZPB_EIO;
    int fout = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fout < 0) { close(fin); return GZPB_EIO; }
#ifdef POSIX_FADV_SEQUENTIAL
    posix_fadvise(fin, 0, 0, POSIX_FADV_SEQUENTIAL);
#endif
    if (blocks_in_flight == 0) {
        // context and slab allocation scale with the batch size: a file that fills only a few batches gets smaller ones
        struct stat sb;
        const size_t bs = buffer_size ? buffer_size : gzpb_default_bufsize(format);
        if (fstat(fin, &sb) == 0 && sb.st_size > 0 && ndevices) {
            const size_t nblocks = (size_t)sb.st_size / bs + 1, per = nblocks / (kLanes * ndevices) + 1;
            blocks_in_flight = std::min((size_t)1184, std::max((size_t)74, per));
        }
    }
    gzpb_writer *w = nullptr;
    int rc = gzpb_writer_create_multi(&w, devices, ndevices, format, level, buffer_size, blocks_in_flight, file_sink, &fout);
    while (rc == GZPB_OK) {
        void *p = nullptr;
        size_t room = 0;
        rc = gzpb_writer_reserve(w, &p, &room);
        if (rc != GZPB_OK) break;
        ssize_t k = read(fin, p, std::min(room, (size_t)16 << 20));
        if (k < 0) { if (errno == EINTR) continue; rc = GZPB_EIO; break; }
        if (k == 0) break;
        rc = gzpb_writer_commit(w, (size_t)k);
    }
    if (w) {
        int r = gzpb_writer_finish(w);
        if (rc == GZPB_OK) rc = r;
        gzpb_writer_stats(w, bytes_in, bytes_out, nullptr, nullptr);
        gzpb_writer_destroy(w);
    }
    close(fin);
    if (close(fout) != 0 && rc == GZPB_OK) rc = GZPB_EIO;
    return rc;
}
#endif
