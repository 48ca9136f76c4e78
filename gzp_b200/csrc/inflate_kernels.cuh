// inflate_kernels.cuh — launch interface of the block-parallel DEFLATE decoder
// (the ParDecompress worker body, /root/reference/src/par/decompress.rs:163-187).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "deflate_kernels.cuh"

namespace gzpb {

// One gzp block (BGZF / Mgzip member) to decode: raw DEFLATE payload at comp + in_off,
// ISIZE bytes to produce at out + out_off, CRC-32 from the member footer.
struct InflateDesc {
    uint64_t in_off;
    uint64_t out_off;
    uint32_t in_len;
    uint32_t out_len;
    uint32_t crc;
    uint32_t pad;
};

struct InflateBatch {
    uint32_t nblocks;
    const uint8_t *comp;       // compressed bytes of the batch (padded by >= 16 readable bytes)
    const InflateDesc *desc;   // nblocks
    uint8_t *out;              // decoded bytes, blocks at their final offsets
    int32_t *status;           // nblocks: 0 ok, 1 bad data, 2 output overrun, 3 input overrun, 4 CRC mismatch
    uint32_t *crc_found;       // nblocks: CRC-32 of what was decoded
    KernelTimer *timer;
};

void upload_inflate_constants();
cudaError_t launch_inflate(const InflateBatch &b, cudaStream_t st);

}  // namespace gzpb
