// ingest: the tensor half of the reference's input pipeline on the device.
//
// data/face_dataset.py:45-90 turns every decoded uint8 image / mask (H, W, C), optionally flipped left-right
// (`img[:, ::-1, :]`, :66-71), into a float32 (C, H, W) tensor divided by 255 (:77-80) on the HOST, one sample at a time
// in DataLoader workers, and the batch then crosses PCIe as fp32.  Here the batch crosses as uint8 (4x fewer bytes:
// 0.39 MB instead of 1.57 MB per 8 RGB images) and one kernel does transpose + flip + convert + divide for the whole
// batch: reads are coalesced along the HWC rows (one thread per pixel loads its C bytes), writes are coalesced per
// channel plane.  float32(x) / 255.0f is an IEEE division, bit-identical to numpy's `.astype('float32')` + `.div(255)`.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace ffwm {

template <int C>
__global__ void ingest_u8_kernel(const uint8_t* __restrict__ src, const uint8_t* __restrict__ flip, float* __restrict__ dst, int b, int h, int w) {
    const int64_t total = (int64_t)b * h * w;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h), n = (int)(i / ((int64_t)w * h));
        const int xs = (flip && flip[n]) ? w - 1 - x : x;                       // img[:, ::-1, :]
        const uint8_t* p = src + (((int64_t)n * h + y) * w + xs) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) dst[(((int64_t)n * C + c) * h + y) * w + x] = (float)p[c] / 255.0f;
    }
}

}  // namespace ffwm

// dst (B, C, H, W) float32 = transpose(src (B, H, W, C) uint8, optionally flipped left-right per sample) / 255.
// flip: B bytes on the device (non-zero = flip) or NULL.  C in {1, 3}.  Both tensors dense.
extern "C" int ffwm_ingest_u8(const void* src, const void* flip, float* dst, int b, int h, int w, int c, void* stream) {
    using namespace ffwm;
    if (b < 0 || h < 0 || w < 0 || (c != 1 && c != 3)) { set_error("ingest_u8: bad shape (B %d, H %d, W %d, C %d; C must be 1 or 3)", b, h, w, c); return FFWM_ERR_SHAPE; }
    const int64_t total = (int64_t)b * h * w;
    if (total == 0) return FFWM_OK;
    if (!src || !dst) { set_error("ingest_u8: null pointer"); return FFWM_ERR_NULL; }
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (c == 3) ingest_u8_kernel<3><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<const uint8_t*>(flip), dst, b, h, w);
    else ingest_u8_kernel<1><<<blocks, 256, 0, st>>>(static_cast<const uint8_t*>(src), static_cast<const uint8_t*>(flip), dst, b, h, w);
    return check_launch("ingest_u8");
}
