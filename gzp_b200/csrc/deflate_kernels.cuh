// deflate_kernels.cuh — host-visible launch interface of the DEFLATE pipeline.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace gzpb {

enum KernelId { KT_CHAIN = 0, KT_MATCH, KT_EMIT, KT_GATHER, KT_CRC, KT_SNAP, KT_INFLATE, KT_COUNT };

// CUDA-event timing of individual kernels on their launching stream.
struct KernelTimer {
    struct Rec { int id; cudaEvent_t a, b; };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    double total_ms[KT_COUNT] = {0};
    uint64_t launches[KT_COUNT] = {0};
    cudaEvent_t get() { if (pool.empty()) { cudaEvent_t e; cudaEventCreate(&e); return e; } cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    void start(int id, cudaStream_t st) { Rec r{id, get(), get()}; cudaEventRecord(r.a, st); pending.push_back(r); }
    void stop(cudaStream_t st) { cudaEventRecord(pending.back().b, st); }
    // call after the stream work has completed
    void collect() {
        for (auto &r : pending) {
            float ms = 0; if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { total_ms[r.id] += ms; launches[r.id]++; }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        pending.clear();
    }
    void reset() { collect(); for (int i = 0; i < KT_COUNT; i++) { total_ms[i] = 0; launches[i] = 0; } }
};

// Device arrays of one batch (all fixed-stride per unit, see gzpb_common.cuh).
struct DeflateBatch {
    uint32_t nunits;
    int level;
    int format;               // gzpb_format
    const uint8_t *in;        // nunits * kInStride
    const uint32_t *unit_len; // nunits: dictionary + data bytes of the unit
    const uint32_t *unit_dict; // nunits: dictionary bytes in front of the data
    const uint32_t *unit_flags;  // bit0 = is_last (BGZF EOF), bit1 = sync flush (no BFINAL)
    uint16_t *next4;          // nunits * m_stride: hash4 chain link of every unit position (distance, 0 = none)
    uint16_t *prev3;          // nunits * m_stride: hash3 link
    uint32_t *lists;          // nunits * spu * 2 * 65536 (k_split position lists)
    uint32_t *list_start;     // nunits * spu * 16
    uint64_t *mtab;           // nunits * m_stride
    uint32_t *mtab2;          // nunits * m_stride, lazy2 levels only (depth/4 column), else NULL
    uint8_t *clen;            // nunits * m_stride (chain-length estimates: k_link -> k_match)
    uint16_t *order;          // nunits * spu * 65536 (positions sorted by chain length)
    uint32_t *crc;            // nunits: Check::sum per unit
    uint32_t *sum_part;       // nunits * spu: per-sub-unit partial sums (k_split -> k_emit), NULL = separate k_check pass
    uint32_t *tokens;         // nunits * kTokStride
    uint8_t *out;             // nunits * kOutStride
    uint32_t *out_len;        // nunits * 2  (total bytes, header offset in the slot)
    int32_t *status;          // nunits
    uint64_t *offsets;        // nunits + 1 (exclusive scan of sizes; [nunits] = total)
    uint8_t *packed;          // compacted stream (device memory or mapped pinned host memory)
    uint64_t packed_cap;      // bytes available at `packed`
    int packed_on_host;       // > 0: `packed` is mapped host memory — k_gather runs with at most this many thread blocks (0 = one per unit)
    const uint64_t *base_ptr; // device-visible address holding this batch's first offset (NULL = 0)
    uint64_t *end_mirror;     // optional second copy of offsets[nunits] (mapped pinned host memory: the next batch of a multi-device stream reads its base there)
    int32_t *overflow;        // set to 1 by k_scan when the batch would exceed packed_cap
    KernelTimer *timer;       // optional
    uint64_t *launch_counter; // optional: += the kernels launch_deflate_pipeline launches (gzpb_launch_count)
    uint32_t in_stride, m_stride, tok_stride, out_stride;   // per-unit strides (bytes / entries / tokens / bytes)
    uint32_t spu, seg;        // sub-units per unit and new positions per sub-unit (gzpb_common.cuh: Geo)
    int check_kind;           // -1 none, 0 CRC-32, 1 Adler-32 (written to `crc`)
};

void upload_deflate_constants();
void read_phase_counters(unsigned long long *out, bool reset);
cudaError_t launch_deflate_pipeline(const DeflateBatch &b, cudaStream_t st);
cudaError_t launch_pack(const DeflateBatch &b, cudaStream_t st);
cudaError_t launch_check_combine(const uint32_t *sums, const uint32_t *unit_len, const uint32_t *unit_dict, uint32_t nunits, int kind,
                                 uint32_t *out3, cudaStream_t st);

}  // namespace gzpb
