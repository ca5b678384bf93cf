// affine_reg: the affine-regularisation quadratic form of FlowNet pre-training as ONE kernel per direction.
//
// The reference (models/losses.py:181-223, AffineRegularizationLoss.calculate_loss) computes, for one coordinate plane
// g (B,1,H,W) of the sampling grid and the fixed matrix Q = K^T K (K = A (A^T A)^-1 A^T - I, A = [row, col, 1]):
//     results   = conv2d(g, Q reshaped to (kz^2, 1, kz, kz))              kz^2 planes of (H-kz+1, W-kz+1)
//     kernels   = LocalAttnReshape(results)                                K6: (B,1,kz*h',kz*w')
//     grid_H    = BlockExtractor(g, flow = kz//2)                          K4 used as an unfold: (B,1,kz*h',kz*w')
//     result    = avg_pool2d(grid_H * kernels, kz, kz)                     (B,1,h',w')
// i.e. five passes and two (kz*h')(kz*w') intermediates (17.5 MB each at kz = 7) for what is, per window position,
//     result[b,y,x] = w^T Q w / kz^2,   w = the kz x kz window of g at (y, x)
// — the squared residual of the best affine fit of the window.  Here one thread owns one window: it holds the window
// in registers, walks Q (shared memory, broadcast reads) row by row, and writes r = w^T Q w / kz^2.  The backward pass
// recomputes Q w per window and scatters 2 * grad * (Q w) / kz^2 into grad_g with REDs (kz^2 per window; windows of
// neighbouring threads overlap, so the REDs of a warp fall on neighbouring addresses).
// The arithmetic is the reference's (conv = sum over n of Q[m][n] w[n], then sum over m of w[m] * that), in the
// operand's own precision; float and double.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace ffwm {

template <typename T, int KZ>
__global__ void __launch_bounds__(128) affine_reg_fwd_kernel(View<const T> g, const T* __restrict__ q, View<T> out) {
    constexpr int K2 = KZ * KZ;
    __shared__ T sq[K2 * K2];
    for (int i = threadIdx.x; i < K2 * K2; i += blockDim.x) sq[i] = q[i];
    __syncthreads();
    const int64_t total = (int64_t)out.n * out.h * out.w;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % out.w), y = (int)((idx / out.w) % out.h), b = (int)(idx / ((int64_t)out.w * out.h));
        const T* gp = g.p + b * g.sb + (int64_t)y * g.sh + (int64_t)x * g.sw;
        T w[K2];
#pragma unroll
        for (int u = 0; u < KZ; ++u)
#pragma unroll
            for (int v = 0; v < KZ; ++v) w[u * KZ + v] = gp[u * g.sh + v * g.sw];
        T acc = T(0);
#pragma unroll 1
        for (int m = 0; m < K2; ++m) {
            T r = T(0);                                                    // results[m] of the reference's conv2d
#pragma unroll
            for (int n = 0; n < K2; ++n) r += sq[m * K2 + n] * w[n];
            // w[m] with a run-time m: the window also lives in global memory (L1 hit), cheaper than a register select
            acc += r * gp[(m / KZ) * g.sh + (m % KZ) * g.sw];
        }
        out.p[b * out.sb + (int64_t)y * out.sh + (int64_t)x * out.sw] = acc / T(K2);
    }
}

// grad_g[b, y+u, x+v] += grad_out[b,y,x] * (2 / kz^2) * (Q w)[u*kz+v]      (Q is symmetric)
template <typename T, int KZ>
__global__ void __launch_bounds__(128) affine_reg_bwd_kernel(View<const T> g, const T* __restrict__ q, View<const T> gout, View<T> gg) {
    constexpr int K2 = KZ * KZ;
    __shared__ T sq[K2 * K2];
    for (int i = threadIdx.x; i < K2 * K2; i += blockDim.x) sq[i] = q[i];
    __syncthreads();
    const int64_t total = (int64_t)gout.n * gout.h * gout.w;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % gout.w), y = (int)((idx / gout.w) % gout.h), b = (int)(idx / ((int64_t)gout.w * gout.h));
        const T* gp = g.p + b * g.sb + (int64_t)y * g.sh + (int64_t)x * g.sw;
        T w[K2];
#pragma unroll
        for (int u = 0; u < KZ; ++u)
#pragma unroll
            for (int v = 0; v < KZ; ++v) w[u * KZ + v] = gp[u * g.sh + v * g.sw];
        const T s = gout.p[b * gout.sb + (int64_t)y * gout.sh + (int64_t)x * gout.sw] * (T(2) / T(K2));
        T* dp = gg.p + b * gg.sb + (int64_t)y * gg.sh + (int64_t)x * gg.sw;
#pragma unroll 1
        for (int m = 0; m < K2; ++m) {
            T r = T(0);
#pragma unroll
            for (int n = 0; n < K2; ++n) r += sq[m * K2 + n] * w[n];
            red_add(dp + (m / KZ) * gg.sh + (m % KZ) * gg.sw, s * r);
        }
    }
}

template <typename T>
static int affine_reg_run(const ffwm_tensor4* grid, const void* q, const ffwm_tensor4* grad_out, const ffwm_tensor4* out, int kz, bool bwd,
                          cudaStream_t st) {
    View<const T> gv, gov;
    View<T> ov;
    int rc;
    const char* what = bwd ? "affine_reg_backward" : "affine_reg_forward";
    if ((rc = make_view<const T>(grid, "grid", &gv))) return rc;
    if ((rc = make_view<T>(out, bwd ? "grad_grid" : "out", &ov))) return rc;
    if (bwd && (rc = make_view<const T>(grad_out, "grad_out", &gov))) return rc;
    if (!q) { set_error("%s: null Q", what); return FFWM_ERR_NULL; }
    if (kz != 3 && kz != 5 && kz != 7) { set_error("%s: kernel size must be 3, 5 or 7 (got %d)", what, kz); return FFWM_ERR_ARG; }
    const int ho = gv.h - kz + 1, wo = gv.w - kz + 1;
    const View<const T>* mv = bwd ? &gov : reinterpret_cast<const View<const T>*>(&ov);
    if (gv.c != 1 || ho < 1 || wo < 1 || mv->n != gv.n || mv->c != 1 || mv->h != ho || mv->w != wo ||
        (bwd && (ov.n != gv.n || ov.c != 1 || ov.h != gv.h || ov.w != gv.w))) {
        set_error("%s: grid (B,1,H,W) %dx%dx%dx%d needs a (B,1,H-kz+1,W-kz+1) map (got %dx%dx%dx%d)", what, gv.n, gv.c, gv.h, gv.w, mv->n,
                  mv->c, mv->h, mv->w);
        return FFWM_ERR_SHAPE;
    }
    const int64_t total = (int64_t)gv.n * ho * wo;
    if (total == 0) return FFWM_OK;
    const int blocks = (int)std::min<int64_t>((total + 127) / 128, (int64_t)sm_count() * 16);
    const T* qp = static_cast<const T*>(q);
#define FFWM_AR_LAUNCH(KZ)                                                              \
    if (bwd) affine_reg_bwd_kernel<T, KZ><<<blocks, 128, 0, st>>>(gv, qp, gov, ov);     \
    else affine_reg_fwd_kernel<T, KZ><<<blocks, 128, 0, st>>>(gv, qp, ov)
    if (kz == 3) { FFWM_AR_LAUNCH(3); } else if (kz == 5) { FFWM_AR_LAUNCH(5); } else { FFWM_AR_LAUNCH(7); }
#undef FFWM_AR_LAUNCH
    return check_launch(what);
}

}  // namespace ffwm

// out (B,1,H-kz+1,W-kz+1) = w^T Q w / kz^2 per kz x kz window w of grid (B,1,H,W); Q: kz^2 x kz^2 row-major device matrix
// of the same dtype (the reference's K^T K).  kz in {3, 5, 7} (the reference uses {1:7, 2:5, 3:3}, models/flownet_model.py:31).
extern "C" int ffwm_affine_reg_forward(const ffwm_tensor4* grid, const void* q, const ffwm_tensor4* out, int kz, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::affine_reg_run<float>(grid, q, nullptr, out, kz, false, st);
    if (dtype == FFWM_F64) return ffwm::affine_reg_run<double>(grid, q, nullptr, out, kz, false, st);
    ffwm::set_error("affine_reg_forward: bad dtype %d", dtype);
    return FFWM_ERR_ARG;
}

// grad_grid (B,1,H,W; ACCUMULATES: zero-fill it) += d/d grid of sum(grad_out * out).
extern "C" int ffwm_affine_reg_backward(const ffwm_tensor4* grid, const void* q, const ffwm_tensor4* grad_out, const ffwm_tensor4* grad_grid,
                                        int kz, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::affine_reg_run<float>(grid, q, grad_out, grad_grid, kz, true, st);
    if (dtype == FFWM_F64) return ffwm::affine_reg_run<double>(grid, q, grad_out, grad_grid, kz, true, st);
    ffwm::set_error("affine_reg_backward: bad dtype %d", dtype);
    return FFWM_ERR_ARG;
}
