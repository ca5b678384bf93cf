// 2x2 / stride-2 max pooling and its gradient (LightCNN: lightcnn/light_cnn.py:38-42,96-124 `nn.MaxPool2d(2, 2, ceil_mode=True)` and
// `F.max_pool2d(x, 2)`; VGG19: models/losses.py:430-470) as one streaming kernel per direction.
//
// ATen keeps an int64 index map per output (max_pool2d_with_indices: 8 bytes written per output forward, read again backward)
// and its NCHW backward kernel walks output windows per INPUT element: 0.33 + 0.49 ms of the train step for 48 launches
// (profiles/r02w_launches_train_summary.txt).  With non-overlapping 2x2 windows every input element belongs to exactly one
// window, so the backward pass recomputes the argmax from the saved input with the forward's scan (rows, then columns;
// `val > max || isnan(val)`: the FIRST maximum of a window wins, a NaN wins over everything — ATen's rule) and writes the
// window's four gradients directly: no index map, no atomics.  ceil_mode windows are clipped at the border.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace ffwm {

// argmax of the (clipped) window at (yo, xo): returns the position 0..3 (row-major) and the value
__device__ __forceinline__ int pool_argmax(const float* __restrict__ xp, int w, int rows, int cols, float* best) {
    float m = xp[0];
    int arg = 0;
    if (cols > 1) { const float v = xp[1]; if (v > m || v != v) m = v, arg = 1; }
    if (rows > 1) {
        { const float v = xp[w]; if (v > m || v != v) m = v, arg = 2; }
        if (cols > 1) { const float v = xp[w + 1]; if (v > m || v != v) m = v, arg = 3; }
    }
    *best = m;
    return arg;
}

template <bool BWD>
__global__ void __launch_bounds__(256) max_pool2x2_kernel(const float* __restrict__ x, const float* __restrict__ go, float* __restrict__ dst,
                                                          int64_t planes, int h, int w, int ho, int wo) {
    const int64_t total = planes * ho * wo;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % wo), yo = (int)((i / wo) % ho);
        const int64_t p = i / ((int64_t)wo * ho);
        const int y0 = 2 * yo, x0 = 2 * xo, rows = min(2, h - y0), cols = min(2, w - x0);
        const float* xp = x + (p * h + y0) * w + x0;
        float m;
        const int arg = pool_argmax(xp, w, rows, cols, &m);
        if (!BWD) dst[i] = m;
        else {
            const float g = go[i];
            float* dp = dst + (p * h + y0) * w + x0;
            dp[0] = arg == 0 ? g : 0.f;
            if (cols > 1) dp[1] = arg == 1 ? g : 0.f;
            if (rows > 1) {
                dp[w] = arg == 2 ? g : 0.f;
                if (cols > 1) dp[w + 1] = arg == 3 ? g : 0.f;
            }
        }
    }
}

static int pool_check(const char* what, int64_t planes, int h, int w, int ho, int wo) {
    if (planes < 0 || h < 1 || w < 1) { set_error("%s: bad sizes", what); return FFWM_ERR_SHAPE; }
    if (!((ho == h / 2 || ho == (h + 1) / 2) && (wo == w / 2 || wo == (w + 1) / 2)) || ho < 1 || wo < 1) {
        set_error("%s: output %dx%d is neither floor nor ceil of %dx%d / 2", what, ho, wo, h, w);
        return FFWM_ERR_SHAPE;
    }
    return FFWM_OK;
}

}  // namespace ffwm

// out (planes, ho, wo) = 2x2 / stride 2 max pooling of x (planes, h, w), contiguous fp32; ho = floor(h/2) or ceil(h/2) (ceil_mode).
extern "C" int ffwm_max_pool2x2_forward(const float* x, float* out, int64_t planes, int h, int w, int ho, int wo, void* stream) {
    using namespace ffwm;
    int rc;
    if ((rc = pool_check("max_pool2x2_forward", planes, h, w, ho, wo))) return rc;
    if (planes == 0) return FFWM_OK;
    if (!x || !out) { set_error("max_pool2x2_forward: null pointer"); return FFWM_ERR_NULL; }
    const int64_t total = planes * ho * wo;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 32);
    max_pool2x2_kernel<false><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, nullptr, out, planes, h, w, ho, wo);
    return check_launch("max_pool2x2_forward");
}

// grad_x (planes, h, w; OVERWRITTEN) = gradient of the pooling: each window's grad_out goes to its first maximum.
extern "C" int ffwm_max_pool2x2_backward(const float* x, const float* grad_out, float* grad_x, int64_t planes, int h, int w, int ho, int wo,
                                         void* stream) {
    using namespace ffwm;
    int rc;
    if ((rc = pool_check("max_pool2x2_backward", planes, h, w, ho, wo))) return rc;
    if (planes == 0) return FFWM_OK;
    if (!x || !grad_out || !grad_x) { set_error("max_pool2x2_backward: null pointer"); return FFWM_ERR_NULL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (2 * ho < h || 2 * wo < w) {                                   // floor mode with an odd size: the last row / column has no window
        cudaError_t e = cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)planes * h * w, st);
        if (e != cudaSuccess) { set_error("max_pool2x2_backward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return int(e); }
    }
    const int64_t total = planes * ho * wo;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 32);
    max_pool2x2_kernel<true><<<blocks, 256, 0, st>>>(x, grad_out, grad_x, planes, h, w, ho, wo);
    return check_launch("max_pool2x2_backward");
}
