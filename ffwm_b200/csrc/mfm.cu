// Max-feature-map activation of LightCNN (lightcnn/light_cnn.py:13-26: conv/linear to 2*C channels, split,
// elementwise max) as one kernel per direction.
//
// STATUS: default on since round 2 (FFWM_FUSED_MFM=0 in ffwm_b200/light_cnn.py for the A/B); B200 parity in
// tests/test_mfm_gpu.py.
//
// Why: PyTorch runs torch.max(a, b) as one kernel forward and FOUR per operand backward (eq, where, lt,
// masked_fill: derivatives.yaml's tie-splitting formula) — 30 MFM layers x 4 LightCNN passes per train step came to
// 600 launches / 3.4 ms of the r01m launch list (profiles/r01m_launches_train_summary.txt: maximum, CompareFunctor,
// where, masked_fill).  Both directions are pure streaming: forward reads 2 and writes 1 value per output element,
// backward reads 3 and writes 2; the bound is HBM (or L2 for the small maps).
//
// Semantics are ATen's, including the corner cases:
//   forward   a != a ? a : (b != b ? b : max(a, b))                                   (NaN propagates)
//   backward  grad_a = a < b ? 0 : (a == b ? g/2 : g),  grad_b = a > b ? 0 : (a == b ? g/2 : g)   (ties split)
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace ffwm {

__device__ __forceinline__ float mfm_max(float a, float b) { return a != a ? a : (b != b ? b : fmaxf(a, b)); }
__device__ __forceinline__ float mfm_ga(float a, float b, float g) { return a < b ? 0.f : (a == b ? g * 0.5f : g); }
__device__ __forceinline__ float mfm_gb(float a, float b, float g) { return a > b ? 0.f : (a == b ? g * 0.5f : g); }

// x (n, 2, chw) contiguous: half 0 = channels [0, C), half 1 = channels [C, 2C); out (n, chw).  VEC = 4 needs
// chw % 4 == 0 and 16-byte aligned bases (then every half starts 16-byte aligned too).
template <int VEC>
__global__ void __launch_bounds__(256) mfm_forward_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, int64_t chw) {
    const int64_t per = chw / VEC, total = n * per;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, e = (i - b * per) * VEC;
        const float* pa = x + b * 2 * chw + e;
        if (VEC == 4) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(pa)), c = __ldcs(reinterpret_cast<const float4*>(pa + chw));
            __stcs(reinterpret_cast<float4*>(out + b * chw + e), make_float4(mfm_max(a.x, c.x), mfm_max(a.y, c.y), mfm_max(a.z, c.z), mfm_max(a.w, c.w)));
        } else {
            out[b * chw + e] = mfm_max(pa[0], pa[chw]);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) mfm_backward_kernel(const float* __restrict__ x, const float* __restrict__ go,
                                                           float* __restrict__ gx, int64_t n, int64_t chw) {
    const int64_t per = chw / VEC, total = n * per;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per, e = (i - b * per) * VEC;
        const float* pa = x + b * 2 * chw + e;
        float* qa = gx + b * 2 * chw + e;
        if (VEC == 4) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(pa)), c = __ldcs(reinterpret_cast<const float4*>(pa + chw));
            const float4 g = __ldcs(reinterpret_cast<const float4*>(go + b * chw + e));
            __stcs(reinterpret_cast<float4*>(qa), make_float4(mfm_ga(a.x, c.x, g.x), mfm_ga(a.y, c.y, g.y), mfm_ga(a.z, c.z, g.z), mfm_ga(a.w, c.w, g.w)));
            __stcs(reinterpret_cast<float4*>(qa + chw), make_float4(mfm_gb(a.x, c.x, g.x), mfm_gb(a.y, c.y, g.y), mfm_gb(a.z, c.z, g.z), mfm_gb(a.w, c.w, g.w)));
        } else {
            const float a = pa[0], c = pa[chw], g = go[b * chw + e];
            qa[0] = mfm_ga(a, c, g);
            qa[chw] = mfm_gb(a, c, g);
        }
    }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int mfm_grid(int64_t work) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)sm_count() * 8)); }

}  // namespace ffwm

// out (n, chw) = max over the two channel halves of x (n, 2*chw); contiguous fp32.  Replaces the
// `torch.max(out[0], out[1])` of mfm.forward (lightcnn/light_cnn.py:24-26).
extern "C" int ffwm_mfm_forward(const float* x, float* out, int64_t n, int64_t chw, void* stream) {
    using namespace ffwm;
    if (n < 0 || chw < 0) { set_error("mfm_forward: negative size"); return FFWM_ERR_SHAPE; }
    if (n == 0 || chw == 0) return FFWM_OK;
    if (!x || !out) { set_error("mfm_forward: null pointer"); return FFWM_ERR_NULL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (chw % 4 == 0 && aligned16(x) && aligned16(out))
        mfm_forward_kernel<4><<<mfm_grid(n * chw / 4), 256, 0, st>>>(x, out, n, chw);
    else
        mfm_forward_kernel<1><<<mfm_grid(n * chw), 256, 0, st>>>(x, out, n, chw);
    return check_launch("mfm_forward");
}

// grad_x (n, 2*chw), overwritten = gradient of ffwm_mfm_forward for grad_out (n, chw), ATen's tie-splitting rule.
extern "C" int ffwm_mfm_backward(const float* x, const float* grad_out, float* grad_x, int64_t n, int64_t chw, void* stream) {
    using namespace ffwm;
    if (n < 0 || chw < 0) { set_error("mfm_backward: negative size"); return FFWM_ERR_SHAPE; }
    if (n == 0 || chw == 0) return FFWM_OK;
    if (!x || !grad_out || !grad_x) { set_error("mfm_backward: null pointer"); return FFWM_ERR_NULL; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (chw % 4 == 0 && aligned16(x) && aligned16(grad_out) && aligned16(grad_x))
        mfm_backward_kernel<4><<<mfm_grid(n * chw / 4), 256, 0, st>>>(x, grad_out, grad_x, n, chw);
    else
        mfm_backward_kernel<1><<<mfm_grid(n * chw), 256, 0, st>>>(x, grad_out, grad_x, n, chw);
    return check_launch("mfm_backward");
}
