// resample2d: GFLA Gaussian-weighted flow resampling, forward + backward.
//
// Semantics restate cuda/resample2d_package/resample2d_kernel.cu of the
// reference (K1 :20-95, K2 :98-202, K3 :204-330) including its numerics
// quirks.  The execution plans are different.
//
// Large fp32 maps (>= ~74 tiles of 16x16, >= 16 channels, kernel_size 2 or 4, dilation 1) — TILED kernels:
//   forward        resample2d_fwd_roll_kernel      rolling-strip channel-lane gather (roll_gather.cuh,
//                  resample2d_roll.cuh); kernel_size 4 only — with 4 taps the direct kernel is faster
//                  (measured, profiles/README.md)
//   grad_input1    scatter_rows_kernel<Resample2dScatterGeo>   row-owner scatter: window rows bucketed by
//                  destination row, sliding register window, vector REDs (scatter_rows.cuh)
//   grad_input2    gather_quad_kernel<RsQuadPolicy>            accumulate-then-weigh: one FFMA per tap and
//                  channel into per-pixel window accumulators, weights once per pixel (gather_quad.cuh);
//                  deterministic, one store per output element
//   dilation > 1 keeps the previous generation (gather_tiled.cuh, scatter_tiled.cuh), also reachable with
//   FFWM_DISABLE_ROLL / FFWM_SCATTER_TILED / FFWM_DISABLE_QUAD for A/B runs (scripts/roll_ab.py).
// Everything else (small maps, fp64, other kernel sizes, FFWM_DISABLE_TILED=1) — DIRECT kernels:
//   * one thread owns one output PIXEL and walks a slice of the channels, so
//     (dx,dy,sigma), the 4*(ks/2) double-precision exps, the tap offsets and
//     the normaliser are computed once per pixel instead of once per element
//     (the reference redoes them C times, and 3*C times in K3);
//   * backward: K2 and K3 are ONE pass over grad_output.  The scatter into
//     grad_input1 uses RED.ADD (as the reference must); the flow gradient is
//     reduced over channels in registers, across the channel slices of a CTA
//     through shared memory, and stored once — no atomics, deterministic.
#include "common.cuh"
#include "gather_tiled.cuh"
#include "scatter_tiled.cuh"
#include "scatter_rows.cuh"
#include "gather_quad.cuh"

namespace ffwm {

// Per-pixel tap geometry shared by forward and backward.
template <typename T, int HALF>
struct Taps {
    int ix[2 * HALF];   // clamped column index: [2*fx] = left, [2*fx+1] = right
    int iy[2 * HALF];   // clamped row index:    [2*fy] = top,  [2*fy+1] = bottom
    T dxs[2 * HALF];    // distances xL_, xR_ (resample2d_kernel.cu:70-73)
    T dys[2 * HALF];    // distances yT_, yB_
};

template <typename T, int HALF>
__device__ __forceinline__ void tap_geometry(T xf, T yf, T alpha, T beta, int dil, int ih, int iw,
                                             Taps<T, HALF>& t) {
    const T fxf = floor(xf), fyf = floor(yf);
#pragma unroll
    for (int f = 0; f < HALF; ++f) {
        t.iy[2 * f] = clampi(f2i(fyf - f * dil), ih - 1);
        t.iy[2 * f + 1] = clampi(f2i(fyf + (f + 1) * dil), ih - 1);
        t.ix[2 * f] = clampi(f2i(fxf - f * dil), iw - 1);
        t.ix[2 * f + 1] = clampi(f2i(fxf + (f + 1) * dil), iw - 1);
        t.dxs[2 * f] = T(f * dil) + alpha;
        t.dxs[2 * f + 1] = T((1. + f) * dil) - alpha;
        t.dys[2 * f] = T(f * dil) + beta;
        t.dys[2 * f + 1] = T((1. + f) * dil) - beta;
    }
}

// Gaussian weights for the distances above and the reference's normaliser,
// summed in the reference's order (one 4-term group per (fy,fx)).
template <typename T, int HALF>
__device__ __forceinline__ T tap_weights(const T* dxs, const T* dys, T sigma, T* wx, T* wy) {
#pragma unroll
    for (int i = 0; i < 2 * HALF; ++i) {
        wx[i] = gauss<T>(dxs[i], sigma);
        wy[i] = gauss<T>(dys[i], sigma);
    }
    T sum = T(0);
#pragma unroll
    for (int fy = 0; fy < HALF; ++fy)
#pragma unroll
        for (int fx = 0; fx < HALF; ++fx)
            sum += (wy[2 * fy] * wx[2 * fx] + wy[2 * fy] * wx[2 * fx + 1] +
                    wy[2 * fy + 1] * wx[2 * fx] + wy[2 * fy + 1] * wx[2 * fx + 1]);
    return sum;
}

// tap_weights with the single-precision expf (<= 2 ulp) instead of the reference's double-precision exp rounded to
// float: used by the tiled fp32 kernels that form the geometry of a pixel once per tile and channel
// group, where the double-precision evaluation (16 exps per pixel) was 11-13 % of all instructions executed.  The
// difference (1e-7 relative) is far inside the path's 1e-5 / 1e-4 tolerances; the direct kernels keep the double exp.
__device__ __forceinline__ float gauss_fast(float d, float sigma) {
    const float num = -d * d, den = 2.f * sigma * sigma;
    return expf(den == 0.f ? num * 1e8f : num / den);
}
template <int HALF>
__device__ __forceinline__ float tap_weights_fast(const float* dxs, const float* dys, float sigma, float* wx, float* wy) {
#pragma unroll
    for (int i = 0; i < 2 * HALF; ++i) {
        wx[i] = gauss_fast(dxs[i], sigma);
        wy[i] = gauss_fast(dys[i], sigma);
    }
    float sum = 0.f;
#pragma unroll
    for (int fy = 0; fy < HALF; ++fy)
#pragma unroll
        for (int fx = 0; fx < HALF; ++fx)
            sum += (wy[2 * fy] * wx[2 * fx] + wy[2 * fy] * wx[2 * fx + 1] +
                    wy[2 * fy + 1] * wx[2 * fx] + wy[2 * fy + 1] * wx[2 * fx + 1]);
    return sum;
}

// Tap list of one pixel for the tiled scatter (scatter_tiled.cuh): destinations are the clamped
// tap indices, weights are K2's SAFE_DIV(w, sum) with the truncation quirk of SURVEY N2 —
// exactly the values the direct kernel below REDs.
template <int HALF>
struct Resample2dScatterGeo {
    static constexpr int NT = 4 * HALF * HALF;
    static constexpr int RW = 31;
    static constexpr int NW = 2 * HALF;          // scatter_rows.cuh: window width (dilation 1 only)
    static constexpr bool CLAMP = true;          // taps outside the image fold onto the border (N5)
    View<const float> in2;
    int dil, ih, iw;
    // window position (left/top to right/bottom) of the reference's tap index: 2f -> floor - f, 2f+1 -> floor + f + 1
    static __device__ __forceinline__ constexpr int win_pos(int k) { return (k & 1) ? HALF + (k >> 1) : HALF - 1 - (k >> 1); }
    // scatter_rows.cuh: the NW x NW window of a pixel in region coordinates and its separable K2 weights
    // (column weight, row weight / normaliser; truncation quirk of SURVEY N2 as in taps()).  dilation 1.
    __device__ __forceinline__ bool window(int b, int y, int x, int rx0, int ry0, int rw, int rh, int& cb, int& rb, float* wx, float* wy) const {
        constexpr int N2 = 2 * HALF;
        const float* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
        const float dx = f[0], dy = f[in2.sc], sigma = f[2 * in2.sc];
        const float xf = float(x) + dx, yf = float(y) + dy;
        const float fxf = floorf(xf), fyf = floorf(yf);
        const bool near = fxf - float(HALF - 1) >= float(rx0) && fxf + float(HALF) <= float(rx0 + rw - 1) &&
                          fyf - float(HALF - 1) >= float(ry0) && fyf + float(HALF) <= float(ry0 + rh - 1);
        if (!near) return false;
        const float alpha2 = xf - float(f2i(xf)), beta2 = yf - float(f2i(yf));
        float d2x[N2], d2y[N2], qx[N2], qy[N2];
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
            d2x[2 * k] = float(k) + alpha2;
            d2x[2 * k + 1] = float(1. + k) - alpha2;
            d2y[2 * k] = float(k) + beta2;
            d2y[2 * k + 1] = float(1. + k) - beta2;
        }
        const float sum2 = tap_weights_fast<HALF>(d2x, d2y, sigma, qx, qy);
#pragma unroll
        for (int k = 0; k < N2; ++k) {
            wx[win_pos(k)] = qx[k];
            wy[win_pos(k)] = sum2 == 0.f ? qy[k] * 1e8f : qy[k] / sum2;
        }
        cb = int(fxf) - (HALF - 1) - rx0;
        rb = int(fyf) - (HALF - 1) - ry0;
        return true;
    }
    __device__ __forceinline__ void region_origin(int tx0, int ty0, int ml, int& rx0, int& ry0) const {
        rx0 = tx0 - ml;
        ry0 = ty0 - ml;
    }
    __device__ __forceinline__ void taps(int b, int y, int x, int* iy, int* ix, float* w) const {
        constexpr int N2 = 2 * HALF;
        const float* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
        const float dx = f[0], dy = f[in2.sc], sigma = f[2 * in2.sc];
        const float xf = float(x) + dx, yf = float(y) + dy;
        Taps<float, HALF> t;
        tap_geometry<float, HALF>(xf, yf, xf - floorf(xf), yf - floorf(yf), dil, ih, iw, t);
        const float alpha2 = xf - float(f2i(xf)), beta2 = yf - float(f2i(yf));
        float d2x[N2], d2y[N2], qx[N2], qy[N2];
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
            d2x[2 * k] = float(k * dil) + alpha2;
            d2x[2 * k + 1] = float((1. + k) * dil) - alpha2;
            d2y[2 * k] = float(k * dil) + beta2;
            d2y[2 * k + 1] = float((1. + k) * dil) - beta2;
        }
        const float sum2 = tap_weights<float, HALF>(d2x, d2y, sigma, qx, qy);
#pragma unroll
        for (int i = 0; i < N2; ++i)
#pragma unroll
            for (int j = 0; j < N2; ++j) {
                iy[i * N2 + j] = t.iy[i];
                ix[i * N2 + j] = t.ix[j];
                w[i * N2 + j] = float(safe_div<float>(qy[i] * qx[j], sum2));
            }
    }
};

}  // namespace ffwm
#include "resample2d_roll.cuh"
namespace ffwm {

// ---------------------------------------------------------------- forward
template <typename T, int HALF>
__global__ void __launch_bounds__(256)
resample2d_fwd_kernel(View<const T> in1, View<const T> in2, View<T> out, int dil, int c_per_block) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= out.h * out.w) return;
    const int b = blockIdx.z;
    const int y = pix / out.w, x = pix - y * out.w;

    const T* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
    const T dx = ld_stream(f), dy = ld_stream(f + in2.sc), sigma = ld_stream(f + 2 * in2.sc);
    const T xf = T(x) + dx, yf = T(y) + dy;

    Taps<T, HALF> t;
    tap_geometry<T, HALF>(xf, yf, xf - floor(xf), yf - floor(yf), dil, in1.h, in1.w, t);
    T wx[2 * HALF], wy[2 * HALF];
    const T sum = tap_weights<T, HALF>(t.dxs, t.dys, sigma, wx, wy);
    int ox[2 * HALF], oy[2 * HALF];
#pragma unroll
    for (int i = 0; i < 2 * HALF; ++i) { ox[i] = t.ix[i] * in1.sw; oy[i] = t.iy[i] * in1.sh; }

    const int c0 = blockIdx.y * c_per_block;
    const int c1 = min(c0 + c_per_block, out.c);
    const T* src = in1.plane(b, c0);
    T* dst = out.plane(b, c0) + y * out.sh + x * out.sw;
#pragma unroll 2
    for (int c = c0; c < c1; ++c, src += in1.sc, dst += out.sc) {
        T val = T(0);
#pragma unroll
        for (int fy = 0; fy < HALF; ++fy)
#pragma unroll
            for (int fx = 0; fx < HALF; ++fx) {
                const T* rt = src + oy[2 * fy];
                const T* rb = src + oy[2 * fy + 1];
                val += wy[2 * fy] * wx[2 * fx] * __ldg(rt + ox[2 * fx]);
                val += wy[2 * fy] * wx[2 * fx + 1] * __ldg(rt + ox[2 * fx + 1]);
                val += wy[2 * fy + 1] * wx[2 * fx] * __ldg(rb + ox[2 * fx]);
                val += wy[2 * fy + 1] * wx[2 * fx + 1] * __ldg(rb + ox[2 * fx + 1]);
            }
        st_stream(dst, T(safe_div<T>(val, sum)));
    }
}

// ---------------------------------------------------------------- forward, tiled
// gather_tiled.cuh explains the plan: lanes are channels, taps come from a shared-memory slab.
// Per-pixel parameters (formed once per tile, read by broadcast): NT tap offsets, the 2*HALF
// column and row weights, the normaliser.  The arithmetic is the direct kernel's, term by term.
template <int HALF>
struct RsFwdParams {
    static constexpr int N2 = 2 * HALF, NT = N2 * N2;
    static constexpr int PW = ((NT + 2 * N2 + 2 + 3) / 4) * 4;     // offsets, wx, wy, sum, far flag; 16-byte rows
};

template <int HALF>
__global__ void __launch_bounds__(GT_THREADS, 1)
resample2d_fwd_tiled_kernel(View<const float> in1, View<const float> in2, View<float> out, int dil, int ml) {
    using PP = RsFwdParams<HALF>;
    constexpr int N2 = PP::N2, NT = PP::NT, PW = PP::PW;
    extern __shared__ __align__(16) unsigned char gt_smem_raw[];
    float* slab = reinterpret_cast<float*>(gt_smem_raw);               // [32][961]
    float* prm = slab + 32 * GT_RPX;                                   // [256][PW]
    float* stage_all = prm + GT_NPX * PW;                              // [16 warps][32][17]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * GT_TW, ty0 = blockIdx.y * GT_TH, b = blockIdx.z;
    const int rx0 = tx0 - ml, ry0 = ty0 - ml;

    if (tid < GT_NPX) {
        const int y = ty0 + tid / GT_TW, x = tx0 + tid % GT_TW;
        if (y < out.h && x < out.w) {
            const float* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
            const float dx = ld_stream(f), dy = ld_stream(f + in2.sc), sigma = ld_stream(f + 2 * in2.sc);
            const float xf = float(x) + dx, yf = float(y) + dy;
            Taps<float, HALF> t;
            tap_geometry<float, HALF>(xf, yf, xf - floorf(xf), yf - floorf(yf), dil, in1.h, in1.w, t);
            float wx[N2], wy[N2];
            const float sum = tap_weights<float, HALF>(t.dxs, t.dys, sigma, wx, wy);
            float* P = prm + tid * PW;
            int* Pi = reinterpret_cast<int*>(P);
            int anyfar = 0;
#pragma unroll
            for (int i = 0; i < N2; ++i)
#pragma unroll
                for (int j = 0; j < N2; ++j) {
                    const int o = gt_tap_offset(t.iy[i], t.ix[j], ry0, rx0, in1.sh, in1.sw);
                    Pi[i * N2 + j] = o;
                    anyfar |= o < 0;
                }
#pragma unroll
            for (int i = 0; i < N2; ++i) { P[NT + i] = wx[i]; P[NT + N2 + i] = wy[i]; }
            P[NT + 2 * N2] = sum;
            Pi[NT + 2 * N2 + 1] = anyfar;
        }
    }
    float* stage = stage_all + warp * (32 * GT_SPITCH);
    const int y = ty0 + warp;                        // this warp's tile row
    for (int c0 = 0; c0 < out.c; c0 += 32) {
        const int nch = min(32, out.c - c0);
        __syncthreads();                             // parameters written / previous group's slab consumed
        gt_fill_slab(slab, in1, b, c0, nch, ry0, rx0, warp, lane);
        gt_fill_wait();
        __syncthreads();
        if (y < out.h) {
            const float* slab_lane = slab + lane * GT_RPX;
            const float* plane_lane = in1.p + b * in1.sb + (int64_t)(c0 + min(lane, nch - 1)) * in1.sc;
#pragma unroll 2
            for (int px = 0; px < GT_TW; ++px) {
                if (tx0 + px >= out.w) break;        // warp-uniform
                const float* P = prm + (warp * GT_TW + px) * PW;
                const float4* P4 = reinterpret_cast<const float4*>(P);
                float w[2 * N2 + 4];
#pragma unroll
                for (int q = 0; q < (2 * N2 + 2 + 3) / 4; ++q) {
                    const float4 v = P4[NT / 4 + q];
                    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
                const float* wx = w;
                const float* wy = w + N2;
                const float sum = w[2 * N2];
                float val = 0.f;
                if (__float_as_int(w[2 * N2 + 1]) == 0) {
                    // every tap inside the slab: one conflict-free shared load per tap, nothing else
                    int off[NT];
#pragma unroll
                    for (int q = 0; q < NT / 4; ++q) {
                        const float4 v = P4[q];
                        off[4 * q] = __float_as_int(v.x); off[4 * q + 1] = __float_as_int(v.y);
                        off[4 * q + 2] = __float_as_int(v.z); off[4 * q + 3] = __float_as_int(v.w);
                    }
#pragma unroll
                    for (int fy = 0; fy < HALF; ++fy)
#pragma unroll
                        for (int fx = 0; fx < HALF; ++fx) {
                            val += wy[2 * fy] * wx[2 * fx] * slab_lane[off[(2 * fy) * N2 + 2 * fx]];
                            val += wy[2 * fy] * wx[2 * fx + 1] * slab_lane[off[(2 * fy) * N2 + 2 * fx + 1]];
                            val += wy[2 * fy + 1] * wx[2 * fx] * slab_lane[off[(2 * fy + 1) * N2 + 2 * fx]];
                            val += wy[2 * fy + 1] * wx[2 * fx + 1] * slab_lane[off[(2 * fy + 1) * N2 + 2 * fx + 1]];
                        }
                } else {
                    // some tap left the halo: same sum, taps fetched one by one (slab or global)
                    const int* Pi = reinterpret_cast<const int*>(P);
#pragma unroll 1
                    for (int f = 0; f < HALF * HALF; ++f) {
                        const int fy = f / HALF, fx = f - fy * HALF;
                        const float wyt = P[NT + N2 + 2 * fy], wyb = P[NT + N2 + 2 * fy + 1];
                        const float wxl = P[NT + 2 * fx], wxr = P[NT + 2 * fx + 1];
                        val += wyt * wxl * gt_load(slab_lane, plane_lane, Pi[(2 * fy) * N2 + 2 * fx]);
                        val += wyt * wxr * gt_load(slab_lane, plane_lane, Pi[(2 * fy) * N2 + 2 * fx + 1]);
                        val += wyb * wxl * gt_load(slab_lane, plane_lane, Pi[(2 * fy + 1) * N2 + 2 * fx]);
                        val += wyb * wxr * gt_load(slab_lane, plane_lane, Pi[(2 * fy + 1) * N2 + 2 * fx + 1]);
                    }
                }
                stage[lane * GT_SPITCH + px] = float(safe_div<float>(val, sum));
            }
            __syncwarp();
            gt_store_row(stage, out, b, c0, nch, y, tx0, lane);
            __syncwarp();
        }
    }
}

template <int HALF>
static int launch_fwd_tiled(const View<const float>& in1, const View<const float>& in2, const View<float>& out,
                            int dil, int ml, cudaStream_t st) {
    using PP = RsFwdParams<HALF>;
    const size_t smem = sizeof(float) * (32 * GT_RPX + GT_NPX * PP::PW + GT_WARPS * 32 * GT_SPITCH);
    cudaError_t e = cudaFuncSetAttribute(resample2d_fwd_tiled_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("resample2d_fwd_tiled: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(ceil_div(out.w, GT_TW), ceil_div(out.h, GT_TH), out.n);
    resample2d_fwd_tiled_kernel<HALF><<<grid, GT_THREADS, smem, st>>>(in1, in2, out, dil, ml);
    return FFWM_OK;
}

// Any kernel_size (runtime ks/2), no per-pixel arrays: weights are recomputed
// per tap.  Only reached for kernel_size > 8; kept so that every argument the
// reference accepts is served by the device path.
template <typename T>
__global__ void __launch_bounds__(256)
resample2d_fwd_generic_kernel(View<const T> in1, View<const T> in2, View<T> out, int half, int dil) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= out.h * out.w) return;
    const int b = blockIdx.z;
    const int y = pix / out.w, x = pix - y * out.w;
    const T* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
    const T dx = f[0], dy = f[in2.sc], sigma = f[2 * in2.sc];
    const T xf = T(x) + dx, yf = T(y) + dy;
    const T alpha = xf - floor(xf), beta = yf - floor(yf);
    for (int c = blockIdx.y; c < out.c; c += gridDim.y) {
        const T* src = in1.plane(b, c);
        T val = T(0), sum = T(0);
        for (int fy = 0; fy < half; ++fy) {
            const int yT = clampi(f2i(floor(yf) - fy * dil), in1.h - 1) * in1.sh;
            const int yB = clampi(f2i(floor(yf) + (fy + 1) * dil), in1.h - 1) * in1.sh;
            const T yT_P = gauss<T>(T(fy * dil) + beta, sigma);
            const T yB_P = gauss<T>(T((1. + fy) * dil) - beta, sigma);
            for (int fx = 0; fx < half; ++fx) {
                const int xL = clampi(f2i(floor(xf) - fx * dil), in1.w - 1) * in1.sw;
                const int xR = clampi(f2i(floor(xf) + (fx + 1) * dil), in1.w - 1) * in1.sw;
                const T xL_P = gauss<T>(T(fx * dil) + alpha, sigma);
                const T xR_P = gauss<T>(T((1. + fx) * dil) - alpha, sigma);
                val += yT_P * xL_P * src[yT + xL];
                val += yT_P * xR_P * src[yT + xR];
                val += yB_P * xL_P * src[yB + xL];
                val += yB_P * xR_P * src[yB + xR];
                sum += (yT_P * xL_P + yT_P * xR_P + yB_P * xL_P + yB_P * xR_P);
            }
        }
        out.plane(b, c)[y * out.sh + x * out.sw] = T(safe_div<T>(val, sum));
    }
}

// --------------------------------------------------------------- flow gradient, tiled
// K3 as a tiled gather (gather_tiled.cuh): lanes are channels, input1's halo region sits in a
// shared-memory slab, grad_output's tile beside it.  Each lane accumulates the four partial sums
// of the direct kernel (A0, A1, A2, Bs) for the 16 pixels of its warp's tile row over all channel
// groups; one butterfly reduction over the lanes at the end gives the channel sums, and lanes
// 0..15 finish one pixel each.  grad_input1 is produced separately by the tiled scatter.
template <int HALF>
struct RsGradParams {
    static constexpr int N2 = 2 * HALF, NT = N2 * N2;
    static constexpr int PW = ((NT + 6 * N2 + 3 + 3) / 4) * 4;   // offsets, wx wy ax ay bx by, sum, sigma, far flag
};

// One pixel of the flow-gradient sums with at least one tap outside the slab (rare): loops are
// kept rolled and read the parameters straight from shared memory.
template <int HALF>
__device__ __noinline__ void rs_gflow_pixel_slow(const float* P, const float* slab_lane, const float* plane_lane,
                                                 float g, float* out4) {
    constexpr int N2 = 2 * HALF, NT = N2 * N2;
    const int* Pi = reinterpret_cast<const int*>(P);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, bs = 0.f;
#pragma unroll 1
    for (int i = 0; i < N2; ++i) {
        float r0 = 0.f, rB = 0.f, r2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < N2; ++j) {
            const float s = g * gt_load(slab_lane, plane_lane, Pi[i * N2 + j]);
            r0 += P[NT + 2 * N2 + j] * s;
            rB += P[NT + j] * s;
            r2 += P[NT + 4 * N2 + j] * s;
        }
        const float wyi = P[NT + N2 + i];
        a0 += wyi * r0;
        a1 += P[NT + 3 * N2 + i] * rB;
        a2 += P[NT + 5 * N2 + i] * rB + wyi * r2;
        bs += wyi * rB;
    }
    out4[0] = a0; out4[1] = a1; out4[2] = a2; out4[3] = bs;
}

template <int HALF>
__global__ void __launch_bounds__(GT_THREADS, 1)
resample2d_gflow_tiled_kernel(View<const float> in1, View<const float> in2, View<const float> gout,
                              View<float> gin2, int dil, int ml) {
    using PP = RsGradParams<HALF>;
    constexpr int N2 = PP::N2, NT = PP::NT, PW = PP::PW;
    extern __shared__ __align__(16) unsigned char gt_smem_raw[];
    float* slab = reinterpret_cast<float*>(gt_smem_raw);               // [32][961]
    float* G = slab + 32 * GT_RPX;                                     // [32][257]
    float* prm = G + 32 * GT_GPITCH;                                   // [256][PW]
    float* accs = prm + GT_NPX * PW;                                   // [16 warps][16 px][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * GT_TW, ty0 = blockIdx.y * GT_TH, b = blockIdx.z;
    const int rx0 = tx0 - ml, ry0 = ty0 - ml;

    for (int i = tid; i < GT_WARPS * GT_TW * 4; i += GT_THREADS) accs[i] = 0.f;
    if (tid < GT_NPX) {
        const int y = ty0 + tid / GT_TW, x = tx0 + tid % GT_TW;
        if (y < gout.h && x < gout.w) {
            const float* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
            const float dx = f[0], dy = f[in2.sc], sigma = f[2 * in2.sc];
            const float xf = float(x) + dx, yf = float(y) + dy;
            Taps<float, HALF> t;
            tap_geometry<float, HALF>(xf, yf, xf - floorf(xf), yf - floorf(yf), dil, in1.h, in1.w, t);
            float wx[N2], wy[N2];
            const float sum = tap_weights<float, HALF>(t.dxs, t.dys, sigma, wx, wy);
            float* P = prm + tid * PW;
            int* Pi = reinterpret_cast<int*>(P);
            int anyfar = 0;
#pragma unroll
            for (int i = 0; i < N2; ++i)
#pragma unroll
                for (int j = 0; j < N2; ++j) {
                    const int o = gt_tap_offset(t.iy[i], t.ix[j], ry0, rx0, in1.sh, in1.sw);
                    Pi[i * N2 + j] = o;
                    anyfar |= o < 0;
                }
#pragma unroll
            for (int i = 0; i < N2; ++i) {
                const float sgn = (i & 1) ? -1.f : 1.f;            // +xL_, -xR_ / +yT_, -yB_
                P[NT + i] = wx[i];
                P[NT + N2 + i] = wy[i];
                P[NT + 2 * N2 + i] = sgn * t.dxs[i] * wx[i];
                P[NT + 3 * N2 + i] = sgn * t.dys[i] * wy[i];
                P[NT + 4 * N2 + i] = t.dxs[i] * t.dxs[i] * wx[i];
                P[NT + 5 * N2 + i] = t.dys[i] * t.dys[i] * wy[i];
            }
            P[NT + 6 * N2] = sum;
            P[NT + 6 * N2 + 1] = sigma;
            Pi[NT + 6 * N2 + 2] = anyfar;
        }
    }
    float* acc = accs + warp * (GT_TW * 4);
    const int y = ty0 + warp;
    for (int c0 = 0; c0 < gout.c; c0 += 32) {
        const int nch = min(32, gout.c - c0);
        __syncthreads();
        gt_fill_slab(slab, in1, b, c0, nch, ry0, rx0, warp, lane);
        gt_fill_tile(G, gout, b, c0, nch, ty0, tx0, tid);
        gt_fill_wait();
        __syncthreads();
        if (y < gout.h) {
            const float* slab_lane = slab + lane * GT_RPX;
            const float* plane_lane = in1.p + b * in1.sb + (int64_t)(c0 + min(lane, nch - 1)) * in1.sc;
            const float* G_lane = G + lane * GT_GPITCH + warp * GT_TW;
#pragma unroll 1
            for (int p4 = 0; p4 < GT_TW / 4; ++p4) {
                float v[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int px = p4 * 4 + k;
                    float r4[4] = {0.f, 0.f, 0.f, 0.f};
                    if (tx0 + px < gout.w) {                      // warp-uniform
                        const float* P = prm + (warp * GT_TW + px) * PW;
                        const float g = lane < nch ? G_lane[px] : 0.f;
                        if (reinterpret_cast<const int*>(P)[NT + 6 * N2 + 2] == 0) {
                            const float4* P4 = reinterpret_cast<const float4*>(P);
                            int off[NT];
                            float w[6 * N2];
#pragma unroll
                            for (int q = 0; q < NT / 4; ++q) {
                                const float4 u = P4[q];
                                off[4 * q] = __float_as_int(u.x); off[4 * q + 1] = __float_as_int(u.y);
                                off[4 * q + 2] = __float_as_int(u.z); off[4 * q + 3] = __float_as_int(u.w);
                            }
#pragma unroll
                            for (int q = 0; q < (6 * N2) / 4; ++q) {
                                const float4 u = P4[NT / 4 + q];
                                w[4 * q] = u.x; w[4 * q + 1] = u.y; w[4 * q + 2] = u.z; w[4 * q + 3] = u.w;
                            }
                            const float *wx = w, *wy = w + N2, *ax = w + 2 * N2, *ay = w + 3 * N2, *bx = w + 4 * N2, *by = w + 5 * N2;
#pragma unroll
                            for (int i = 0; i < N2; ++i) {
                                float r0 = 0.f, rB = 0.f, r2 = 0.f;
#pragma unroll
                                for (int j = 0; j < N2; ++j) {
                                    const float s = g * slab_lane[off[i * N2 + j]];
                                    r0 += ax[j] * s;
                                    rB += wx[j] * s;
                                    r2 += bx[j] * s;
                                }
                                r4[0] += wy[i] * r0;
                                r4[1] += ay[i] * rB;
                                r4[2] += by[i] * rB + wy[i] * r2;
                                r4[3] += wy[i] * rB;
                            }
                        } else {
                            rs_gflow_pixel_slow<HALF>(P, slab_lane, plane_lane, g, r4);
                        }
                    }
                    v[4 * k] = r4[0]; v[4 * k + 1] = r4[1]; v[4 * k + 2] = r4[2]; v[4 * k + 3] = r4[3];
                }
                const float tot = gt_packed_reduce<16>(v, lane);       // lane l: value index l >> 1
                if ((lane & 1) == 0) acc[p4 * 16 + (lane >> 1)] += tot;
            }
        }
    }
    if (y >= gout.h) return;                          // whole warp; no block barrier follows
    __syncwarp();
    const int x = tx0 + lane;
    if (lane >= GT_TW || x >= gout.w) return;
    const float a0 = acc[lane * 4], a1 = acc[lane * 4 + 1], a2 = acc[lane * 4 + 2], bs = acc[lane * 4 + 3];
    const float* P = prm + (warp * GT_TW + lane) * PW;
    float Wx = 0.f, Wy = 0.f, AX = 0.f, AY = 0.f, BX = 0.f, BY = 0.f;
#pragma unroll
    for (int i = 0; i < N2; ++i) {
        Wx += P[NT + i]; Wy += P[NT + N2 + i];
        AX += P[NT + 2 * N2 + i]; AY += P[NT + 3 * N2 + i];
        BX += P[NT + 4 * N2 + i]; BY += P[NT + 5 * N2 + i];
    }
    const float sum = P[NT + 6 * N2], sigma = P[NT + 6 * N2 + 1];
    const float ms2 = -sigma * sigma, s3 = sigma * sigma * sigma;
    const float G0 = float(safe_div<float>(Wy * AX, ms2));
    const float G1 = float(safe_div<float>(AY * Wx, ms2));
    const float G2 = float(safe_div<float>(BY * Wx + Wy * BX, s3));
    const float g10 = float(safe_div<float>(a0, ms2));
    const float g11 = float(safe_div<float>(a1, ms2));
    const float g12 = float(safe_div<float>(a2, s3));
    const float ss = sum * sum;
    float* o = gin2.p + b * gin2.sb + y * gin2.sh + x * gin2.sw;
    o[0] = float(safe_div<float>(g10, sum) - safe_div<float>(G0 * bs, ss));
    if (gin2.c > 1) o[gin2.sc] = float(safe_div<float>(g11, sum) - safe_div<float>(G1 * bs, ss));
    if (gin2.c > 2) o[2 * gin2.sc] = float(safe_div<float>(g12, sum) - safe_div<float>(G2 * bs, ss));
}

template <int HALF>
static int launch_gflow_tiled(const View<const float>& in1, const View<const float>& in2, const View<const float>& gout,
                              const View<float>& g2, int dil, int ml, cudaStream_t st) {
    using PP = RsGradParams<HALF>;
    const size_t smem = sizeof(float) * (32 * GT_RPX + 32 * GT_GPITCH + GT_NPX * PP::PW + GT_WARPS * GT_TW * 4);
    cudaError_t e = cudaFuncSetAttribute(resample2d_gflow_tiled_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("resample2d_gflow_tiled: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(ceil_div(gout.w, GT_TW), ceil_div(gout.h, GT_TH), gout.n);
    resample2d_gflow_tiled_kernel<HALF><<<grid, GT_THREADS, smem, st>>>(in1, in2, gout, g2, dil, ml);
    return FFWM_OK;
}

// --------------------------------------------------------------- flow gradient, accumulate-then-weigh
// K3 on gather_quad.cuh: M[i][j] = sum_c grad_output[c] * input1[c, window(i,j)] per pixel (window positions,
// left/top to right/bottom), then the direct kernel's four partial sums A0, A1, A2, Bs as bilinear forms of M and
// the reference's final SAFE_DIV expressions, once per pixel.
template <int HALF>
struct RsQuadPolicy {
    static constexpr int NW = 2 * HALF;
    static constexpr bool PAD_ZERO = false;                  // the reference clamps every tap index: edge replication
    View<const float> in1, in2, go;
    View<float> gin2;
    __host__ __device__ __forceinline__ const View<const float>& src() const { return in1; }
    __host__ __device__ __forceinline__ const View<const float>& gout() const { return go; }
    static __device__ __forceinline__ constexpr int win_pos(int k) { return (k & 1) ? HALF + (k >> 1) : HALF - 1 - (k >> 1); }

    // record: [1] [2] first window column / row (unclamped, bounded), [3] sigma [4] alpha [5] beta [6] sum,
    // [8 .. 8+N2) wx, [8+N2 .. 8+2*N2) wy (reference tap-index order)
    __device__ __forceinline__ int geometry(int b, int y, int x, int rx0, int ry0, float* rec) const {
        constexpr int N2 = 2 * HALF;
        const float* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
        const float dx = f[0], dy = f[in2.sc], sigma = f[2 * in2.sc];
        const float xf = float(x) + dx, yf = float(y) + dy;
        const float fxf = floorf(xf), fyf = floorf(yf);
        const float alpha = xf - fxf, beta = yf - fyf;
        float dxs[N2], dys[N2], wx[N2], wy[N2];
#pragma unroll
        for (int k = 0; k < HALF; ++k) {
            dxs[2 * k] = float(k) + alpha;
            dxs[2 * k + 1] = float(1. + k) - alpha;
            dys[2 * k] = float(k) + beta;
            dys[2 * k + 1] = float(1. + k) - beta;
        }
        const float sum = tap_weights_fast<HALF>(dxs, dys, sigma, wx, wy);
        int* ri = reinterpret_cast<int*>(rec);
        ri[1] = min(max(f2i(fxf - float(HALF - 1)), -8), in1.w + 8);
        ri[2] = min(max(f2i(fyf - float(HALF - 1)), -8), in1.h + 8);
        rec[3] = sigma; rec[4] = alpha; rec[5] = beta; rec[6] = sum;
#pragma unroll
        for (int k = 0; k < N2; ++k) { rec[8 + k] = wx[k]; rec[8 + N2 + k] = wy[k]; }
        const bool near = fxf - float(HALF - 1) >= float(rx0) && fxf + float(HALF) <= float(rx0 + GQ_RW - 1) &&
                          fyf - float(HALF - 1) >= float(ry0) && fyf + float(HALF) <= float(ry0 + GQ_RH - 1);
        return near ? (ri[2] - ry0) * GQ_RW + (ri[1] - rx0) : -1;
    }
    __device__ __forceinline__ float far_tap(int iy, int ix, const float* plane) const {
        return __ldg(plane + clampi(iy, in1.h - 1) * in1.sh + clampi(ix, in1.w - 1) * in1.sw);
    }
    __device__ __forceinline__ void finish(const float* M, const float* rec, int b, int y, int x) const {
        constexpr int N2 = 2 * HALF;
        const float sigma = rec[3], alpha = rec[4], beta = rec[5], sum = rec[6];
        float wx[N2], wy[N2];
#pragma unroll
        for (int k = 0; k < N2; ++k) { wx[k] = rec[8 + k]; wy[k] = rec[8 + N2 + k]; }
        RsGradCoef<HALF> co;
        rs_grad_coef<HALF>(alpha, beta, wx, wy, co);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, bs = 0.f;
        float Wx = 0.f, Wy = 0.f, AX = 0.f, AY = 0.f, BX = 0.f, BY = 0.f;
#pragma unroll
        for (int ki = 0; ki < N2; ++ki) {
            const float* Mi = M + win_pos(ki) * NW;
            float r0 = 0.f, rB = 0.f, r2 = 0.f;
#pragma unroll
            for (int kj = 0; kj < N2; ++kj) {
                const float m = Mi[win_pos(kj)];
                r0 += co.ax[kj] * m;
                rB += wx[kj] * m;
                r2 += co.bx[kj] * m;
            }
            a0 += wy[ki] * r0;
            a1 += co.ay[ki] * rB;
            a2 += co.by[ki] * rB + wy[ki] * r2;
            bs += wy[ki] * rB;
            Wx += wx[ki]; Wy += wy[ki];
            AX += co.ax[ki]; AY += co.ay[ki];
            BX += co.bx[ki]; BY += co.by[ki];
        }
        const float ms2 = -sigma * sigma, s3 = sigma * sigma * sigma;
        const float G0 = float(safe_div<float>(Wy * AX, ms2));
        const float G1 = float(safe_div<float>(AY * Wx, ms2));
        const float G2 = float(safe_div<float>(BY * Wx + Wy * BX, s3));
        const float g10 = float(safe_div<float>(a0, ms2));
        const float g11 = float(safe_div<float>(a1, ms2));
        const float g12 = float(safe_div<float>(a2, s3));
        const float ss = sum * sum;
        float* o = gin2.p + b * gin2.sb + y * gin2.sh + x * gin2.sw;
        o[0] = float(safe_div<float>(g10, sum) - safe_div<float>(G0 * bs, ss));
        if (gin2.c > 1) o[gin2.sc] = float(safe_div<float>(g11, sum) - safe_div<float>(G1 * bs, ss));
        if (gin2.c > 2) o[2 * gin2.sc] = float(safe_div<float>(g12, sum) - safe_div<float>(G2 * bs, ss));
    }
};

// --------------------------------------------------------------- backward
// CTA = PX pixels x SL channel slices (PX*SL = 256).  Slice s owns channels
// s, s+SL, s+2SL, ...  Per channel and tap: one gather of input1, one RED
// into grad_input1 (skipped when gin1.p == nullptr), three FMAs for the
// flow-gradient partial sums.
template <typename T, int HALF, int SL>
__global__ void __launch_bounds__(256)
resample2d_bwd_kernel(View<const T> in1, View<const T> in2, View<const T> gout,
                      View<T> gin1, View<T> gin2, int dil) {
    constexpr int PX = 256 / SL;
    constexpr int NT = 2 * HALF;
    __shared__ T red[SL > 1 ? SL : 1][4][PX];

    const int lane_px = threadIdx.x;          // 0..PX-1
    const int slice = threadIdx.y;            // 0..SL-1
    const int pix = blockIdx.x * PX + lane_px;
    const int b = blockIdx.z;
    const bool live = pix < gout.h * gout.w;

    T A0 = T(0), A1 = T(0), A2 = T(0), Bs = T(0);
    T sum = T(0), sigma = T(1);
    T Wx = T(0), Wy = T(0), AX = T(0), AY = T(0), BX = T(0), BY = T(0);
    int y = 0, x = 0;

    if (live) {
        y = pix / gout.w;
        x = pix - y * gout.w;
        const T* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
        const T dx = f[0], dy = f[in2.sc];
        sigma = f[2 * in2.sc];
        const T xf = T(x) + dx, yf = T(y) + dy;

        Taps<T, HALF> t;
        tap_geometry<T, HALF>(xf, yf, xf - floor(xf), yf - floor(yf), dil, in1.h, in1.w, t);
        T wx[NT], wy[NT];
        sum = tap_weights<T, HALF>(t.dxs, t.dys, sigma, wx, wy);

        // K2 weights: alpha = xf - int(xf) (truncation).  They only
        // differ from the floor version for negative non-integer coordinates.
        T qx[NT], qy[NT], sum2 = sum;
        const T alpha2 = xf - T(f2i(xf)), beta2 = yf - T(f2i(yf));
        if (alpha2 != t.dxs[0] || beta2 != t.dys[0]) {
            T d2x[NT], d2y[NT];
#pragma unroll
            for (int k = 0; k < HALF; ++k) {
                d2x[2 * k] = T(k * dil) + alpha2;
                d2x[2 * k + 1] = T((1. + k) * dil) - alpha2;
                d2y[2 * k] = T(k * dil) + beta2;
                d2y[2 * k + 1] = T((1. + k) * dil) - beta2;
            }
            sum2 = tap_weights<T, HALF>(d2x, d2y, sigma, qx, qy);
        } else {
#pragma unroll
            for (int i = 0; i < NT; ++i) { qx[i] = wx[i]; qy[i] = wy[i]; }
        }
        // q[i][j] = SAFE_DIV(w, sum) of K2 :195-198, rounded to T.  The product
        // q*gOut is formed in double by the reference and rounded once; for
        // T-representable q that equals the T product.  sum2 == 0 implies every
        // w == 0 (weights are non-negative), so the EPS branch also yields 0.
        T q[NT][NT];
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) q[i][j] = T(safe_div<T>(qy[i] * qx[j], sum2));

        // Separable coefficients of the three flow-gradient numerators
        // (resample2d_kernel.cu:271-294): sign * distance * weight.
        T ax[NT], ay[NT], bx[NT], by[NT];
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const T sgn_x = (i & 1) ? T(-1) : T(1);   // +xL_, -xR_
            const T sgn_y = (i & 1) ? T(-1) : T(1);   // +yT_, -yB_
            ax[i] = sgn_x * t.dxs[i] * wx[i];
            ay[i] = sgn_y * t.dys[i] * wy[i];
            bx[i] = t.dxs[i] * t.dxs[i] * wx[i];
            by[i] = t.dys[i] * t.dys[i] * wy[i];
            Wx += wx[i]; Wy += wy[i];
            AX += ax[i]; AY += ay[i];
            BX += bx[i]; BY += by[i];
        }

        const bool want1 = gin1.p != nullptr;
        const bool want2 = gin2.p != nullptr;
        const int goff = y * gout.sh + x * gout.sw;
        for (int c = slice; c < gout.c; c += SL) {
            const T g = ld_stream(gout.plane(b, c) + goff);
            const T* src = in1.plane(b, c);
            T* dst = want1 ? gin1.plane(b, c) : nullptr;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                T r0 = T(0), rB = T(0), r2 = T(0);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    if (want2) {
                        const T s = g * __ldg(src + t.iy[i] * in1.sh + t.ix[j] * in1.sw);
                        r0 += ax[j] * s;
                        rB += wx[j] * s;
                        r2 += bx[j] * s;
                    }
                    if (want1) {
                        // gin1 has in1's geometry but its own strides
                        red_add(dst + t.iy[i] * gin1.sh + t.ix[j] * gin1.sw, q[i][j] * g);
                    }
                }
                A0 += wy[i] * r0;
                A1 += ay[i] * rB;
                A2 += by[i] * rB + wy[i] * r2;
                Bs += wy[i] * rB;
            }
        }
    }

    if (gin2.p == nullptr) return;   // uniform across the CTA
    if (SL > 1) {
        red[slice][0][lane_px] = A0;
        red[slice][1][lane_px] = A1;
        red[slice][2][lane_px] = A2;
        red[slice][3][lane_px] = Bs;
        __syncthreads();
        if (slice != 0) return;
#pragma unroll
        for (int s = 1; s < SL; ++s) {
            A0 += red[s][0][lane_px];
            A1 += red[s][1][lane_px];
            A2 += red[s][2][lane_px];
            Bs += red[s][3][lane_px];
        }
    }
    if (!live) return;

    // grad1_c = A_c / div_c ; grad2_c = (sumgrad_c / C) * Bs with
    // sumgrad_c / C = sum_t coef_t * w_t / div_c   (K3 :271-294,318-328).
    const T ms2 = -sigma * sigma, s3 = sigma * sigma * sigma;
    const T G0 = T(safe_div<T>(Wy * AX, ms2));
    const T G1 = T(safe_div<T>(AY * Wx, ms2));
    const T G2 = T(safe_div<T>(BY * Wx + Wy * BX, s3));
    const T g10 = T(safe_div<T>(A0, ms2));
    const T g11 = T(safe_div<T>(A1, ms2));
    const T g12 = T(safe_div<T>(A2, s3));
    const T ss = sum * sum;
    T* o = gin2.p + b * gin2.sb + y * gin2.sh + x * gin2.sw;
    o[0] = T(safe_div<T>(g10, sum) - safe_div<T>(G0 * Bs, ss));
    if (gin2.c > 1) o[gin2.sc] = T(safe_div<T>(g11, sum) - safe_div<T>(G1 * Bs, ss));
    if (gin2.c > 2) o[2 * gin2.sc] = T(safe_div<T>(g12, sum) - safe_div<T>(G2 * Bs, ss));
}

// Generic-ks backward (kernel_size > 8): thread per pixel, channels looped,
// weights recomputed per tap.  Correct for every argument, not tuned.
template <typename T>
__global__ void __launch_bounds__(256)
resample2d_bwd_generic_kernel(View<const T> in1, View<const T> in2, View<const T> gout,
                              View<T> gin1, View<T> gin2, int half, int dil) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= gout.h * gout.w) return;
    const int b = blockIdx.z;
    const int y = pix / gout.w, x = pix - y * gout.w;
    const T* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
    const T dx = f[0], dy = f[in2.sc], sigma = f[2 * in2.sc];
    const T xf = T(x) + dx, yf = T(y) + dy;
    const T alpha = xf - floor(xf), beta = yf - floor(yf);
    const T alpha2 = xf - T(f2i(xf)), beta2 = yf - T(f2i(yf));
    const T ms2 = -sigma * sigma, s3 = sigma * sigma * sigma;

    T sum = T(0), sum2 = T(0), G0 = T(0), G1 = T(0), G2 = T(0);
    for (int fy = 0; fy < half; ++fy)
        for (int fx = 0; fx < half; ++fx) {
            const T xd[2] = {T(fx * dil) + alpha, T((1. + fx) * dil) - alpha};
            const T yd[2] = {T(fy * dil) + beta, T((1. + fy) * dil) - beta};
            const T xd2[2] = {T(fx * dil) + alpha2, T((1. + fx) * dil) - alpha2};
            const T yd2[2] = {T(fy * dil) + beta2, T((1. + fy) * dil) - beta2};
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j) {
                    const T w = gauss<T>(yd[i], sigma) * gauss<T>(xd[j], sigma);
                    sum += w;
                    sum2 += gauss<T>(yd2[i], sigma) * gauss<T>(xd2[j], sigma);
                    G0 += T(safe_div<T>((j ? -xd[j] : xd[j]) * w, ms2));
                    G1 += T(safe_div<T>((i ? -yd[i] : yd[i]) * w, ms2));
                    G2 += T(safe_div<T>((yd[i] * yd[i] + xd[j] * xd[j]) * w, s3));
                }
        }

    T A0 = T(0), A1 = T(0), A2 = T(0), Bs = T(0);
    for (int c = 0; c < gout.c; ++c) {
        const T g = gout.plane(b, c)[y * gout.sh + x * gout.sw];
        const T* src = in1.plane(b, c);
        T* dst = gin1.p ? gin1.plane(b, c) : nullptr;
        for (int fy = 0; fy < half; ++fy)
            for (int fx = 0; fx < half; ++fx) {
                const int yy[2] = {clampi(f2i(floor(yf) - fy * dil), in1.h - 1),
                                   clampi(f2i(floor(yf) + (fy + 1) * dil), in1.h - 1)};
                const int xx[2] = {clampi(f2i(floor(xf) - fx * dil), in1.w - 1),
                                   clampi(f2i(floor(xf) + (fx + 1) * dil), in1.w - 1)};
                const T xd[2] = {T(fx * dil) + alpha, T((1. + fx) * dil) - alpha};
                const T yd[2] = {T(fy * dil) + beta, T((1. + fy) * dil) - beta};
                const T xd2[2] = {T(fx * dil) + alpha2, T((1. + fx) * dil) - alpha2};
                const T yd2[2] = {T(fy * dil) + beta2, T((1. + fy) * dil) - beta2};
                for (int i = 0; i < 2; ++i)
                    for (int j = 0; j < 2; ++j) {
                        const T w = gauss<T>(yd[i], sigma) * gauss<T>(xd[j], sigma);
                        const T s = g * src[yy[i] * in1.sh + xx[j] * in1.sw];
                        A0 += T(safe_div<T>((j ? -xd[j] : xd[j]) * w * s, ms2));
                        A1 += T(safe_div<T>((i ? -yd[i] : yd[i]) * w * s, ms2));
                        A2 += T(safe_div<T>((yd[i] * yd[i] + xd[j] * xd[j]) * w * s, s3));
                        Bs += w * s;
                        if (dst) {
                            const T w2 = gauss<T>(yd2[i], sigma) * gauss<T>(xd2[j], sigma);
                            red_add(dst + yy[i] * gin1.sh + xx[j] * gin1.sw,
                                    T(safe_div<T>(w2, sum2) * double(g)));
                        }
                    }
            }
    }
    if (!gin2.p) return;
    const T ss = sum * sum;
    T* o = gin2.p + b * gin2.sb + y * gin2.sh + x * gin2.sw;
    o[0] = T(safe_div<T>(A0, sum) - safe_div<T>(G0 * Bs, ss));
    if (gin2.c > 1) o[gin2.sc] = T(safe_div<T>(A1, sum) - safe_div<T>(G1 * Bs, ss));
    if (gin2.c > 2) o[2 * gin2.sc] = T(safe_div<T>(A2, sum) - safe_div<T>(G2 * Bs, ss));
}

// ------------------------------------------------------------ host launch
template <typename T, int HALF>
static void launch_fwd(const View<const T>& in1, const View<const T>& in2, const View<T>& out,
                       int dil, cudaStream_t st) {
    const int hw = out.h * out.w;
    const int pix_blocks = ceil_div(hw, 256);
    // enough CTAs for ~8 per SM, but at least 4 channels per CTA so the
    // per-pixel exps stay amortised
    int64_t want = (int64_t)8 * sm_count();
    int chunks = int((want + (int64_t)pix_blocks * out.n - 1) / ((int64_t)pix_blocks * out.n));
    chunks = max(1, min(chunks, ceil_div(out.c, 4)));
    chunks = min(chunks, 65535);
    const int c_per_block = ceil_div(out.c, chunks);
    chunks = ceil_div(out.c, c_per_block);
    dim3 grid(pix_blocks, chunks, out.n);
    resample2d_fwd_kernel<T, HALF><<<grid, 256, 0, st>>>(in1, in2, out, dil, c_per_block);
}

template <typename T, int HALF, int SL>
static void launch_bwd_sl(const View<const T>& in1, const View<const T>& in2, const View<const T>& gout,
                          const View<T>& g1, const View<T>& g2, int dil, cudaStream_t st) {
    constexpr int PX = 256 / SL;
    dim3 grid(ceil_div(gout.h * gout.w, PX), 1, gout.n), block(PX, SL);
    resample2d_bwd_kernel<T, HALF, SL><<<grid, block, 0, st>>>(in1, in2, gout, g1, g2, dil);
}

template <typename T, int HALF>
static void launch_bwd(const View<const T>& in1, const View<const T>& in2, const View<const T>& gout,
                       const View<T>& g1, const View<T>& g2, int dil, cudaStream_t st) {
    const int c = gout.c;
    if (c >= 8) launch_bwd_sl<T, HALF, 8>(in1, in2, gout, g1, g2, dil, st);
    else if (c >= 4) launch_bwd_sl<T, HALF, 4>(in1, in2, gout, g1, g2, dil, st);
    else if (c >= 2) launch_bwd_sl<T, HALF, 2>(in1, in2, gout, g1, g2, dil, st);
    else launch_bwd_sl<T, HALF, 1>(in1, in2, gout, g1, g2, dil, st);
}

template <typename T>
static int resample2d_forward_t(const ffwm_tensor4* a, const ffwm_tensor4* b, const ffwm_tensor4* o,
                                int ks, int dil, cudaStream_t st) {
    View<const T> in1, in2;
    View<T> out;
    int rc;
    if ((rc = make_view<const T>(a, "input1", &in1))) return rc;
    if ((rc = make_view<const T>(b, "input2", &in2))) return rc;
    if ((rc = make_view<T>(o, "output", &out))) return rc;
    if (in2.c < 3) { set_error("resample2d: input2 needs 3 channels (dx,dy,sigma), got %d", in2.c); return FFWM_ERR_SHAPE; }
    if (out.n != in2.n || out.h != in2.h || out.w != in2.w || out.c != in1.c || in1.n < out.n) {
        set_error("resample2d: output (%d,%d,%d,%d) inconsistent with input1 C=%d N=%d / input2 (%d,3,%d,%d)",
                  out.n, out.c, out.h, out.w, in1.c, in1.n, in2.n, in2.h, in2.w);
        return FFWM_ERR_SHAPE;
    }
    if (ks < 0 || dil < 0) { set_error("resample2d: kernel_size=%d dilation=%d", ks, dil); return FFWM_ERR_ARG; }
    if ((int64_t)out.n * out.c * out.h * out.w == 0) return FFWM_OK;
    if (out.n > 65535) { set_error("resample2d: batch %d > 65535", out.n); return FFWM_ERR_TOO_LARGE; }
    if (in1.h == 0 || in1.w == 0) { set_error("resample2d: empty input1 plane"); return FFWM_ERR_SHAPE; }
    const int half = ks / 2;
    if constexpr (sizeof(T) == 4) {
        // rolling-strip gather (roll_gather.cuh): kernel_size 2 and 4, dilation 1
        // measured at the cfg5 point (scripts/roll_ab.py): kernel_size 4 0.89 ms (tiled gather 1.05, direct 1.41);
        // kernel_size 2 0.67 ms against 0.53 ms for the direct kernel, so 4-tap calls stay direct unless forced
        if ((half == 2 || (half == 1 && opt(OPT_FORCE_ROLL))) && dil == 1 && in1.n >= out.n &&
            roll_applicable(out.n, out.c, out.h, out.w, in1, ceil_div(out.c, 32))) {
            const int rc2 = half == 1 ? launch_fwd_roll<1>(in1, in2, out, st) : launch_fwd_roll<2>(in1, in2, out, st);
            if (rc2) return rc2;
            return check_launch("resample2d_forward(roll)");
        }
        const int hmax = (15 - (2 * half - 1) * dil) / 2;
        // 16-tap kernels only: with 4 taps the direct kernel (0.54 ms at the cfg5 point) beats the
        // tiled one (0.78 ms), whose slab fill then outweighs the gather it saves
        if (half == 2 && dil >= 1 && hmax >= 2 && in1.n >= out.n &&
            (int64_t)(in1.h - 1) * in1.sh + (int64_t)(in1.w - 1) * in1.sw < (1 << 30) &&
            gather_tiled_applicable(out.n, out.c, out.h, out.w, in1)) {
            const int ml = hmax + (half - 1) * dil;
            const int rc2 = half == 1 ? launch_fwd_tiled<1>(in1, in2, out, dil, ml, st) : launch_fwd_tiled<2>(in1, in2, out, dil, ml, st);
            if (rc2) return rc2;
            return check_launch("resample2d_forward(tiled)");
        }
    }
    switch (half) {
        case 1: launch_fwd<T, 1>(in1, in2, out, dil, st); break;
        case 2: launch_fwd<T, 2>(in1, in2, out, dil, st); break;
        case 3: launch_fwd<T, 3>(in1, in2, out, dil, st); break;
        case 4: launch_fwd<T, 4>(in1, in2, out, dil, st); break;
        default: {
            dim3 grid(ceil_div(out.h * out.w, 256), min(out.c, 64), out.n);
            resample2d_fwd_generic_kernel<T><<<grid, 256, 0, st>>>(in1, in2, out, half, dil);
        }
    }
    return check_launch("resample2d_forward");
}

template <typename T>
static int resample2d_backward_t(const ffwm_tensor4* a, const ffwm_tensor4* b, const ffwm_tensor4* go,
                                 const ffwm_tensor4* ga, const ffwm_tensor4* gb,
                                 int ks, int dil, cudaStream_t st) {
    View<const T> in1, in2, gout;
    View<T> g1, g2;
    int rc;
    if ((rc = make_view<const T>(a, "input1", &in1))) return rc;
    if ((rc = make_view<const T>(b, "input2", &in2))) return rc;
    if ((rc = make_view<const T>(go, "grad_output", &gout))) return rc;
    if ((rc = make_view<T>(ga, "grad_input1", &g1, true))) return rc;
    if ((rc = make_view<T>(gb, "grad_input2", &g2, true))) return rc;
    if (in2.c < 3) { set_error("resample2d: input2 needs 3 channels"); return FFWM_ERR_SHAPE; }
    if (gout.n != in2.n || gout.h != in2.h || gout.w != in2.w || gout.c != in1.c || in1.n < gout.n) {
        set_error("resample2d_backward: grad_output (%d,%d,%d,%d) inconsistent with inputs", gout.n, gout.c, gout.h, gout.w);
        return FFWM_ERR_SHAPE;
    }
    if (g1.p && (g1.n != in1.n || g1.c != in1.c || g1.h != in1.h || g1.w != in1.w)) {
        set_error("resample2d_backward: grad_input1 shape differs from input1"); return FFWM_ERR_SHAPE;
    }
    if (g2.p && (g2.n != in2.n || g2.h != in2.h || g2.w != in2.w || g2.c < 1 || g2.c > 3)) {
        set_error("resample2d_backward: grad_input2 shape differs from input2"); return FFWM_ERR_SHAPE;
    }
    if (ks < 0 || dil < 0) { set_error("resample2d: kernel_size=%d dilation=%d", ks, dil); return FFWM_ERR_ARG; }
    if ((int64_t)gout.n * gout.h * gout.w == 0) return FFWM_OK;
    if (gout.n > 65535) { set_error("resample2d: batch %d > 65535", gout.n); return FFWM_ERR_TOO_LARGE; }
    if (in1.h == 0 || in1.w == 0) { set_error("resample2d: empty input1 plane"); return FFWM_ERR_SHAPE; }
    const int half = ks / 2;
    if constexpr (sizeof(T) == 4) {
        // grad_input1 through the tiled scatter when the problem is large enough; the fused kernel
        // below then only produces the flow gradient
        const int hmax = (15 - (2 * half - 1) * dil) / 2;
        if (g1.p && (half == 1 || half == 2) && dil >= 1 && hmax >= 2 && scatter_tiled_applicable(gout, g1)) {
            const int ml = hmax + (half - 1) * dil;
            int rc2;
            if (dil == 1 && !opt(OPT_SCATTER_TILED)) {      // row-owner scatter (scatter_rows.cuh)
                if (half == 1) rc2 = launch_scatter_rows(Resample2dScatterGeo<1>{in2, dil, in1.h, in1.w}, gout, g1, st);
                else rc2 = launch_scatter_rows(Resample2dScatterGeo<2>{in2, dil, in1.h, in1.w}, gout, g1, st);
            } else if (half == 1) rc2 = launch_scatter_tiled(Resample2dScatterGeo<1>{in2, dil, in1.h, in1.w}, gout, g1, ml, st);
            else rc2 = launch_scatter_tiled(Resample2dScatterGeo<2>{in2, dil, in1.h, in1.w}, gout, g1, ml, st);
            if (rc2) return rc2;
            if ((rc2 = check_launch("resample2d_backward(tiled scatter)"))) return rc2;
            if (!g2.p) return FFWM_OK;
            g1.p = nullptr;
        }
        // flow gradient: accumulate-then-weigh gather (gather_quad.cuh)
        if (!g1.p && g2.p && (half == 1 || half == 2) && dil == 1 && !opt(OPT_DISABLE_TILED_GFLOW) &&
            gather_quad_applicable(gout.n, gout.c, gout.h, gout.w, in1)) {
            const int rc2 = half == 1 ? launch_gather_quad(RsQuadPolicy<1>{in1, in2, gout, g2}, gout.n, gout.h, gout.w, st)
                                      : launch_gather_quad(RsQuadPolicy<2>{in1, in2, gout, g2}, gout.n, gout.h, gout.w, st);
            if (rc2) return rc2;
            return check_launch("resample2d_backward(quad flow gradient)");
        }
        // dilation > 1: the tiled gather
        if (!g1.p && g2.p && (half == 1 || half == 2) && dil >= 1 && hmax >= 2 &&
            (int64_t)(in1.h - 1) * in1.sh + (int64_t)(in1.w - 1) * in1.sw < (1 << 30) &&
            gather_tiled_applicable(gout.n, gout.c, gout.h, gout.w, in1) && !opt(OPT_DISABLE_TILED_GFLOW)) {
            const int ml = hmax + (half - 1) * dil;
            const int rc2 = half == 1 ? launch_gflow_tiled<1>(in1, in2, gout, g2, dil, ml, st)
                                      : launch_gflow_tiled<2>(in1, in2, gout, g2, dil, ml, st);
            if (rc2) return rc2;
            return check_launch("resample2d_backward(tiled flow gradient)");
        }
    }
    switch (half) {
        case 1: launch_bwd<T, 1>(in1, in2, gout, g1, g2, dil, st); break;
        case 2: launch_bwd<T, 2>(in1, in2, gout, g1, g2, dil, st); break;
        case 3: launch_bwd<T, 3>(in1, in2, gout, g1, g2, dil, st); break;
        case 4: launch_bwd<T, 4>(in1, in2, gout, g1, g2, dil, st); break;
        default: {
            dim3 grid(ceil_div(gout.h * gout.w, 256), 1, gout.n);
            resample2d_bwd_generic_kernel<T><<<grid, 256, 0, st>>>(in1, in2, gout, g1, g2, half, dil);
        }
    }
    return check_launch("resample2d_backward");
}

}  // namespace ffwm

extern "C" int ffwm_resample2d_forward(const ffwm_tensor4* input1, const ffwm_tensor4* input2,
                                       const ffwm_tensor4* output, int kernel_size, int dilation,
                                       int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::resample2d_forward_t<float>(input1, input2, output, kernel_size, dilation, st);
    if (dtype == FFWM_F64) return ffwm::resample2d_forward_t<double>(input1, input2, output, kernel_size, dilation, st);
    ffwm::set_error("resample2d_forward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}

extern "C" int ffwm_resample2d_backward(const ffwm_tensor4* input1, const ffwm_tensor4* input2,
                                        const ffwm_tensor4* grad_output, const ffwm_tensor4* grad_input1,
                                        const ffwm_tensor4* grad_input2, int kernel_size, int dilation,
                                        int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32)
        return ffwm::resample2d_backward_t<float>(input1, input2, grad_output, grad_input1, grad_input2, kernel_size, dilation, st);
    if (dtype == FFWM_F64)
        return ffwm::resample2d_backward_t<double>(input1, input2, grad_output, grad_input1, grad_input2, kernel_size, dilation, st);
    ffwm::set_error("resample2d_backward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}
