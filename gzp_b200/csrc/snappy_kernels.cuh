// snappy_kernels.cuh — launch interface of the Snap (Snappy framed) encoder.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "deflate_kernels.cuh"

namespace gzpb {

struct SnapBatch {
    uint32_t nunits;
    uint32_t cpu;              // 64 KiB chunks per unit
    const uint8_t *in;         // nunits * in_stride
    const uint32_t *unit_len;  // nunits
    uint32_t in_stride;
    uint8_t *out;              // nunits * cpu * out_stride (one slot per chunk)
    uint32_t out_stride;
    uint32_t *out_len;         // nunits * cpu * 2
    KernelTimer *timer;
};

void upload_snappy_constants();
cudaError_t launch_snap(const SnapBatch &b, cudaStream_t st);

}  // namespace gzpb
