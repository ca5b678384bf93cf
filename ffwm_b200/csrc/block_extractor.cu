// block_extractor: k x k bilinear block extraction around a flow field.
//
// Semantics restate cuda/block_extractor/block_extractor_kernel.cu of the
// reference (K4 :20-85, K5 :89-170): for every flow pixel
// (yf,xf) and block offset (i,j) the source is sampled bilinearly at
// (xf + flow_x + j - k/2, yf + flow_y + i - k/2), tap indices clamped after
// the weights are formed.
//
// Execution plans (not the reference's thread-per-element):
//   forward, k = 2,3 (fp32)  block_extractor_fwd_tiled_kernel: the k*k samples of a flow pixel share a
//             (k+1)x(k+1) source window held in registers; every lane stores its k consecutive floats
//             of an output row directly; source rows are prefetched into L2/L1.
//   backward, k = 2,3 (fp32) block_extractor_bwd_window_kernel: the same window folding for the
//             scatter ((k+1)^2 REDs instead of 4*k*k per flow pixel and channel) and the gather.
//   any other k / fp64       one thread per output pixel (forward) or per flow pixel x channel
//             slice (backward); grad_source is a scatter (RED.ADD, as in the reference).
//   grad_flow — k*k*C contributions per address, the reference's worst atomic hotspot — is always
//   reduced in registers, then across the CTA's channel slices through shared memory, and stored
//   once: no atomics, deterministic.
#include <stdlib.h>

#include "common.cuh"

namespace ffwm {

template <typename T>
struct Bilin {
    int xL, xR, yT, yB;          // clamped indices
    T xL_P, xR_P, yT_P, yB_P;    // weights from the unclamped coordinate
};

// block_extractor_kernel.cu:57-76 — same expression order.
template <typename T>
__device__ __forceinline__ Bilin<T> block_tap(T flow_x_raw, T flow_y_raw, int xf, int yf,
                                              int xoff, int yoff, int hs, int ws) {
    const T flow_y = flow_y_raw + yoff;
    const T flow_x = flow_x_raw + xoff;
    const T dy = flow_y + T(yf);
    const T dx = flow_x + T(xf);
    const T fdx = floor(dx), fdy = floor(dy);
    Bilin<T> r;
    r.xL = clampi(f2i(fdx), ws - 1);
    r.xR = clampi(f2i(fdx + 1), ws - 1);
    r.yT = clampi(f2i(fdy), hs - 1);
    r.yB = clampi(f2i(fdy + 1), hs - 1);
    r.xL_P = 1 - (dx - fdx);
    r.xR_P = dx - fdx;
    r.yT_P = 1 - (dy - fdy);
    r.yB_P = dy - fdy;
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256)
block_extractor_fwd_kernel(View<const T> src, View<const T> flow, View<T> out, int k, int c_per_block) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= out.h * out.w) return;
    const int b = blockIdx.z;
    const int y = pix / out.w, x = pix - y * out.w;
    const int yf = y / k, xf = x / k;
    const int yoff = y - yf * k - k / 2, xoff = x - xf * k - k / 2;

    const T* f = flow.p + b * flow.sb + yf * flow.sh + xf * flow.sw;
    const Bilin<T> t = block_tap<T>(__ldg(f), __ldg(f + flow.sc), xf, yf, xoff, yoff, src.h, src.w);
    const int oTL = t.yT * src.sh + t.xL * src.sw, oTR = t.yT * src.sh + t.xR * src.sw;
    const int oBL = t.yB * src.sh + t.xL * src.sw, oBR = t.yB * src.sh + t.xR * src.sw;
    const T wTL = t.xL_P * t.yT_P, wTR = t.xR_P * t.yT_P, wBL = t.xL_P * t.yB_P, wBR = t.xR_P * t.yB_P;

    const int c0 = blockIdx.y * c_per_block;
    const int c1 = min(c0 + c_per_block, out.c);
    const T* s = src.plane(b, c0);
    T* d = out.plane(b, c0) + y * out.sh + x * out.sw;
#pragma unroll 4
    for (int c = c0; c < c1; ++c, s += src.sc, d += out.sc) {
        T sample = T(0);
        sample += wTL * __ldg(s + oTL);
        sample += wTR * __ldg(s + oTR);
        sample += wBL * __ldg(s + oBL);
        sample += wBR * __ldg(s + oBR);
        st_stream(d, sample);
    }
}

// ---- forward, windowed: k is a template parameter -----------------------------------------
// A warp owns 32 consecutive flow pixels of one flow row and walks a slice of channels.
// The k*k samples of one flow pixel share a (k+1)x(k+1) source window (same fractional
// part, integer offsets 0..k-1), so the window is loaded once into registers — (k+1)^2
// gathers instead of 4*k*k.  The output is k*k times larger than the source, so the kernel is
// store-bound: every lane stores its k consecutive floats of an output row directly (the warp
// covers 32*k consecutive floats; k stores of stride k).  Measured at the cfg5 point against
// transposing each row through shared memory into 128-bit stores: 0.434 vs 0.480 ms for the
// reference's flow distribution, equal for incoherent flows — the transposition costs three times
// the L1 wavefronts of the strided stores, and L2 merges the partial sectors.
// Geometry (weights, clamped indices) is formed per tap with the reference's expressions
// (block_tap); the shared window is only used when the clamped tap indices really are
// consecutive, otherwise that pixel gathers its four taps directly.
constexpr int BE_WARPS = 8;

template <typename T, int K>
__global__ void __launch_bounds__(32 * BE_WARPS, 2)
block_extractor_fwd_tiled_kernel(View<const T> src, View<const T> flow, View<T> out, int chunks, int c_per_block) {
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int xf0 = blockIdx.x * 32, xf = xf0 + lane;
    const int yf = blockIdx.y * BE_WARPS + warp;
    const int b = blockIdx.z / chunks, chunk = blockIdx.z - b * chunks;
    if (yf >= flow.h || xf >= flow.w) return;

    // Separable tap geometry: column part depends on j only, row part on i only.
    int cx[K + 1], cy[K + 1];       // element offsets of the shared window's columns / rows
    T xLP[K], xRP[K], yTP[K], yBP[K];
    bool shared_window = true;
    const T* f = flow.p + b * flow.sb + yf * flow.sh + xf * flow.sw;
    const T fx_raw = __ldg(f), fy_raw = __ldg(f + flow.sc);
    {
        int prev_xR = 0, prev_yB = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const Bilin<T> t = block_tap<T>(fx_raw, fy_raw, xf, yf, j - K / 2, j - K / 2, src.h, src.w);
            xLP[j] = t.xL_P; xRP[j] = t.xR_P; yTP[j] = t.yT_P; yBP[j] = t.yB_P;
            cx[j] = t.xL * src.sw; cy[j] = t.yT * src.sh;
            if (j > 0) shared_window = shared_window && prev_xR == t.xL && prev_yB == t.yT;
            prev_xR = t.xR; prev_yB = t.yB;
        }
        cx[K] = prev_xR * src.sw;
        cy[K] = prev_yB * src.sh;
    }

    const int c0 = chunk * c_per_block;
    const int c1 = min(c0 + c_per_block, out.c);
    const T* s = src.plane(b, c0);
    T* obase = out.plane(b, c0) + (yf * K) * out.sh + (xf * K) * out.sw;
    // The first touch of a source row is a DRAM round trip (~1500 cycles under load) and every
    // channel touches new rows: rows are requested far ahead into L2, and the window of channel
    // c+1 is loaded into registers before channel c is computed and stored (software pipeline).
    constexpr int PF_L2 = 16;
    T nxt[K + 1][K + 1];
    if (shared_window) {
#pragma unroll
        for (int n = 0; n <= K; ++n)
#pragma unroll
            for (int m = 0; m <= K; ++m) nxt[n][m] = __ldg(s + cy[n] + cx[m]);
    }
    for (int c = c0; c < c1; ++c, s += src.sc, obase += out.sc) {
        if (shared_window) {
            T win[K + 1][K + 1];
#pragma unroll
            for (int n = 0; n <= K; ++n)
#pragma unroll
                for (int m = 0; m <= K; ++m) win[n][m] = nxt[n][m];
            if (c + 1 < c1) {
                const T* s1 = s + src.sc;
#pragma unroll
                for (int n = 0; n <= K; ++n)
#pragma unroll
                    for (int m = 0; m <= K; ++m) nxt[n][m] = __ldg(s1 + cy[n] + cx[m]);
            }
            if (c + PF_L2 < c1) {
                const T* ps = s + (int64_t)PF_L2 * src.sc + cx[0];
#pragma unroll
                for (int n = 0; n <= K; ++n) asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + cy[n]));
            }
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    T sample = T(0);
                    sample += xLP[j] * yTP[i] * win[i][j];
                    sample += xRP[j] * yTP[i] * win[i][j + 1];
                    sample += xLP[j] * yBP[i] * win[i + 1][j];
                    sample += xRP[j] * yBP[i] * win[i + 1][j + 1];
                    st_stream(obase + i * out.sh + j * out.sw, sample);
                }
        } else {
            // rare: the clamped taps of this pixel are not consecutive (float rounding at an
            // integer boundary); gather the four taps of every sample directly
#pragma unroll 1
            for (int i = 0; i < K; ++i)
#pragma unroll 1
                for (int j = 0; j < K; ++j) {
                    const Bilin<T> t = block_tap<T>(fx_raw, fy_raw, xf, yf, j - K / 2, i - K / 2, src.h, src.w);
                    T sample = T(0);
                    sample += t.xL_P * t.yT_P * __ldg(s + t.yT * src.sh + t.xL * src.sw);
                    sample += t.xR_P * t.yT_P * __ldg(s + t.yT * src.sh + t.xR * src.sw);
                    sample += t.xL_P * t.yB_P * __ldg(s + t.yB * src.sh + t.xL * src.sw);
                    sample += t.xR_P * t.yB_P * __ldg(s + t.yB * src.sh + t.xR * src.sw);
                    st_stream(obase + i * out.sh + j * out.sw, sample);
                }
        }
    }
}

// ---- backward, windowed: k is a template parameter -----------------------------------------
// Same decomposition as the direct kernel below (CTA = 32 flow pixels x 8 channel slices, flow
// gradient reduced in registers + shared memory, stored once), but per channel the k*k samples of a
// flow pixel are folded into their shared (k+1)x(k+1) source window in registers first:
//   * grad_source: (k+1)^2 REDs per (flow pixel, channel) instead of 4*k*k (16 vs 36 for k=3),
//     and lanes are neighbouring flow pixels, so the REDs of a warp hit neighbouring addresses;
//   * flow gradient: the source window is loaded once, (k+1)^2 gathers instead of 4*k*k.
// Pixels whose clamped taps are not consecutive (float rounding at an integer boundary, rare)
// take the per-tap path.
template <int K, int SL>
__global__ void __launch_bounds__(256, 2)
block_extractor_bwd_window_kernel(View<const float> src, View<const float> flow, View<const float> gout,
                                  View<float> gsrc, View<float> gflow) {
    constexpr int PX = 256 / SL;
    __shared__ float red[SL][2][PX];
    const int lane_px = threadIdx.x, slice = threadIdx.y;
    const int fpix = blockIdx.x * PX + lane_px;
    const int b = blockIdx.z;
    const bool live = fpix < flow.h * flow.w;
    const bool want_src = gsrc.p != nullptr, want_flow = gflow.p != nullptr;

    float gx = 0.f, gy = 0.f;
    int yf = 0, xf = 0;
    if (live) {
        yf = fpix / flow.w;
        xf = fpix - yf * flow.w;
        const float* f = flow.p + b * flow.sb + yf * flow.sh + xf * flow.sw;
        const float fx_raw = __ldg(f), fy_raw = __ldg(f + flow.sc);
        int cx[K + 1], cy[K + 1];
        float xLP[K], xRP[K], yTP[K], yBP[K];
        bool shared_window = true;
        {
            int prev_xR = 0, prev_yB = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const Bilin<float> t = block_tap<float>(fx_raw, fy_raw, xf, yf, j - K / 2, j - K / 2, src.h, src.w);
                xLP[j] = t.xL_P; xRP[j] = t.xR_P; yTP[j] = t.yT_P; yBP[j] = t.yB_P;
                cx[j] = t.xL; cy[j] = t.yT;
                if (j > 0) shared_window = shared_window && prev_xR == t.xL && prev_yB == t.yT;
                prev_xR = t.xR; prev_yB = t.yB;
            }
            cx[K] = prev_xR;
            cy[K] = prev_yB;
        }
        const int gbase = (yf * K) * gout.sh + (xf * K) * gout.sw;
        if (shared_window) {
            for (int c = slice; c < gout.c; c += SL) {
                const float* gp = gout.plane(b, c) + gbase;
                float g[K][K];
#pragma unroll
                for (int i = 0; i < K; ++i)
#pragma unroll
                    for (int j = 0; j < K; ++j) g[i][j] = ld_stream(gp + i * gout.sh + j * gout.sw);
                if (want_flow) {
                    const float* s = src.plane(b, c);
                    float win[K + 1][K + 1];
#pragma unroll
                    for (int n = 0; n <= K; ++n)
#pragma unroll
                        for (int m = 0; m <= K; ++m) win[n][m] = __ldg(s + cy[n] * src.sh + cx[m] * src.sw);
#pragma unroll
                    for (int i = 0; i < K; ++i)
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            const float xL_yT = win[i][j], xR_yT = win[i][j + 1], xL_yB = win[i + 1][j], xR_yB = win[i + 1][j + 1];
                            gy += g[i][j] * (-xLP[j] * xL_yT - xRP[j] * xR_yT + xLP[j] * xL_yB + xRP[j] * xR_yB);
                            gx += g[i][j] * (-yTP[i] * xL_yT - yBP[i] * xL_yB + yTP[i] * xR_yT + yBP[i] * xR_yB);
                        }
                }
                if (want_src) {
                    float acc[K + 1][K + 1];
#pragma unroll
                    for (int n = 0; n <= K; ++n)
#pragma unroll
                        for (int m = 0; m <= K; ++m) acc[n][m] = 0.f;
#pragma unroll
                    for (int i = 0; i < K; ++i)
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            acc[i][j] += g[i][j] * xLP[j] * yTP[i];
                            acc[i][j + 1] += g[i][j] * xRP[j] * yTP[i];
                            acc[i + 1][j] += g[i][j] * xLP[j] * yBP[i];
                            acc[i + 1][j + 1] += g[i][j] * xRP[j] * yBP[i];
                        }
                    float* d = gsrc.plane(b, c);
#pragma unroll
                    for (int n = 0; n <= K; ++n)
#pragma unroll
                        for (int m = 0; m <= K; ++m) red_add(d + cy[n] * gsrc.sh + cx[m] * gsrc.sw, acc[n][m]);
                }
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < K; ++i)
#pragma unroll 1
                for (int j = 0; j < K; ++j) {
                    const Bilin<float> t = block_tap<float>(fx_raw, fy_raw, xf, yf, j - K / 2, i - K / 2, src.h, src.w);
                    for (int c = slice; c < gout.c; c += SL) {
                        const float grad = ld_stream(gout.plane(b, c) + gbase + i * gout.sh + j * gout.sw);
                        if (want_src) {
                            float* d = gsrc.plane(b, c);
                            red_add(d + t.yT * gsrc.sh + t.xL * gsrc.sw, grad * t.xL_P * t.yT_P);
                            red_add(d + t.yT * gsrc.sh + t.xR * gsrc.sw, grad * t.xR_P * t.yT_P);
                            red_add(d + t.yB * gsrc.sh + t.xL * gsrc.sw, grad * t.xL_P * t.yB_P);
                            red_add(d + t.yB * gsrc.sh + t.xR * gsrc.sw, grad * t.xR_P * t.yB_P);
                        }
                        if (want_flow) {
                            const float* s = src.plane(b, c);
                            const float xL_yT = __ldg(s + t.yT * src.sh + t.xL * src.sw), xR_yT = __ldg(s + t.yT * src.sh + t.xR * src.sw);
                            const float xL_yB = __ldg(s + t.yB * src.sh + t.xL * src.sw), xR_yB = __ldg(s + t.yB * src.sh + t.xR * src.sw);
                            gy += grad * (-t.xL_P * xL_yT - t.xR_P * xR_yT + t.xL_P * xL_yB + t.xR_P * xR_yB);
                            gx += grad * (-t.yT_P * xL_yT - t.yB_P * xL_yB + t.yT_P * xR_yT + t.yB_P * xR_yB);
                        }
                    }
                }
        }
    }
    if (!want_flow) return;
    red[slice][0][lane_px] = gx;
    red[slice][1][lane_px] = gy;
    __syncthreads();
    if (slice != 0) return;
#pragma unroll
    for (int s = 1; s < SL; ++s) {
        gx += red[s][0][lane_px];
        gy += red[s][1][lane_px];
    }
    if (!live) return;
    float* o = gflow.p + b * gflow.sb + yf * gflow.sh + xf * gflow.sw;
    o[0] = gx;
    o[gflow.sc] = gy;
}

template <typename T, int SL>
__global__ void __launch_bounds__(256)
block_extractor_bwd_kernel(View<const T> src, View<const T> flow, View<const T> gout,
                           View<T> gsrc, View<T> gflow, int k) {
    constexpr int PX = 256 / SL;
    __shared__ T red[SL > 1 ? SL : 1][2][PX];

    const int lane_px = threadIdx.x, slice = threadIdx.y;
    const int fpix = blockIdx.x * PX + lane_px;
    const int b = blockIdx.z;
    const bool live = fpix < flow.h * flow.w;
    const bool want_src = gsrc.p != nullptr, want_flow = gflow.p != nullptr;

    T gx = T(0), gy = T(0);
    int yf = 0, xf = 0;
    if (live) {
        yf = fpix / flow.w;
        xf = fpix - yf * flow.w;
        const T* f = flow.p + b * flow.sb + yf * flow.sh + xf * flow.sw;
        const T fx_raw = __ldg(f), fy_raw = __ldg(f + flow.sc);
        for (int i = 0; i < k; ++i) {
            for (int j = 0; j < k; ++j) {
                const Bilin<T> t = block_tap<T>(fx_raw, fy_raw, xf, yf, j - k / 2, i - k / 2, src.h, src.w);
                const int goff = (yf * k + i) * gout.sh + (xf * k + j) * gout.sw;
                const int sTL = t.yT * src.sh + t.xL * src.sw, sTR = t.yT * src.sh + t.xR * src.sw;
                const int sBL = t.yB * src.sh + t.xL * src.sw, sBR = t.yB * src.sh + t.xR * src.sw;
                const int dTL = t.yT * gsrc.sh + t.xL * gsrc.sw, dTR = t.yT * gsrc.sh + t.xR * gsrc.sw;
                const int dBL = t.yB * gsrc.sh + t.xL * gsrc.sw, dBR = t.yB * gsrc.sh + t.xR * gsrc.sw;
#pragma unroll 2
                for (int c = slice; c < gout.c; c += SL) {
                    const T grad = ld_stream(gout.plane(b, c) + goff);
                    if (want_src) {
                        T* d = gsrc.plane(b, c);
                        red_add(d + dTL, grad * t.xL_P * t.yT_P);
                        red_add(d + dTR, grad * t.xR_P * t.yT_P);
                        red_add(d + dBL, grad * t.xL_P * t.yB_P);
                        red_add(d + dBR, grad * t.xR_P * t.yB_P);
                    }
                    if (want_flow) {
                        const T* s = src.plane(b, c);
                        const T xL_yT = __ldg(s + sTL), xR_yT = __ldg(s + sTR);
                        const T xL_yB = __ldg(s + sBL), xR_yB = __ldg(s + sBR);
                        gy += grad * (-t.xL_P * xL_yT - t.xR_P * xR_yT + t.xL_P * xL_yB + t.xR_P * xR_yB);
                        gx += grad * (-t.yT_P * xL_yT - t.yB_P * xL_yB + t.yT_P * xR_yT + t.yB_P * xR_yB);
                    }
                }
            }
        }
    }

    if (!want_flow) return;
    if (SL > 1) {
        red[slice][0][lane_px] = gx;
        red[slice][1][lane_px] = gy;
        __syncthreads();
        if (slice != 0) return;
#pragma unroll
        for (int s = 1; s < SL; ++s) {
            gx += red[s][0][lane_px];
            gy += red[s][1][lane_px];
        }
    }
    if (!live) return;
    T* o = gflow.p + b * gflow.sb + yf * gflow.sh + xf * gflow.sw;
    o[0] = gx;
    o[gflow.sc] = gy;
}

template <typename T, int SL>
static void launch_bwd_sl(const View<const T>& src, const View<const T>& flow, const View<const T>& gout,
                          const View<T>& gs, const View<T>& gf, int k, cudaStream_t st) {
    constexpr int PX = 256 / SL;
    dim3 grid(ceil_div(flow.h * flow.w, PX), 1, gout.n), block(PX, SL);
    block_extractor_bwd_kernel<T, SL><<<grid, block, 0, st>>>(src, flow, gout, gs, gf, k);
}

template <typename T>
static int block_extractor_forward_t(const ffwm_tensor4* a, const ffwm_tensor4* b, const ffwm_tensor4* o,
                                     int k, cudaStream_t st) {
    View<const T> src, flow;
    View<T> out;
    int rc;
    if ((rc = make_view<const T>(a, "source", &src))) return rc;
    if ((rc = make_view<const T>(b, "flow_field", &flow))) return rc;
    if ((rc = make_view<T>(o, "output", &out))) return rc;
    if (k < 1) { set_error("block_extractor: kernel_size=%d", k); return FFWM_ERR_ARG; }
    if (flow.c != 2) { set_error("block_extractor: flow_field needs 2 channels, got %d", flow.c); return FFWM_ERR_SHAPE; }
    if (out.n != flow.n || src.n < out.n || out.c != src.c ||
        (int64_t)out.h != (int64_t)k * flow.h || (int64_t)out.w != (int64_t)k * flow.w) {
        set_error("block_extractor: output (%d,%d,%d,%d) != (B,C,k*Hf,k*Wf) for k=%d, flow (%d,2,%d,%d), source C=%d",
                  out.n, out.c, out.h, out.w, k, flow.n, flow.h, flow.w, src.c);
        return FFWM_ERR_SHAPE;
    }
    if ((int64_t)out.n * out.c * out.h * out.w == 0) return FFWM_OK;
    if (out.n > 65535) { set_error("block_extractor: batch %d > 65535", out.n); return FFWM_ERR_TOO_LARGE; }
    if (src.h == 0 || src.w == 0) { set_error("block_extractor: empty source plane"); return FFWM_ERR_SHAPE; }
    if constexpr (sizeof(T) == 4) if (k == 2 || k == 3) {
        // (a rolling-strip channel-lane variant with the source staged in shared memory was measured at 0.64 ms
        // against 0.30 ms for this kernel at the cfg5 point and removed)
        const int tx = ceil_div(flow.w, 32), ty = ceil_div(flow.h, BE_WARPS);
        // ~8 CTAs per SM overall, at least 4 channels per CTA so the per-pixel geometry stays amortised
        int64_t want = (int64_t)8 * sm_count();
        int64_t tiles = (int64_t)tx * ty * out.n;
        int chunks = int((want + tiles - 1) / tiles);
        chunks = max(1, min(chunks, ceil_div(out.c, 4)));
        chunks = min(chunks, max(1, 65535 / out.n));
        const int c_per_block = ceil_div(out.c, chunks);
        chunks = ceil_div(out.c, c_per_block);
        if (ty <= 65535 && (int64_t)out.n * chunks <= 65535) {
            dim3 grid(tx, ty, out.n * chunks), block(32, BE_WARPS);
            switch (k) {
                case 2: block_extractor_fwd_tiled_kernel<T, 2><<<grid, block, 0, st>>>(src, flow, out, chunks, c_per_block); break;
                default: block_extractor_fwd_tiled_kernel<T, 3><<<grid, block, 0, st>>>(src, flow, out, chunks, c_per_block); break;
            }
            return check_launch("block_extractor_forward");
        }
    }
    const int pix_blocks = ceil_div((int64_t)out.h * out.w, 256);
    int64_t want = (int64_t)8 * sm_count();
    int chunks = int((want + (int64_t)pix_blocks * out.n - 1) / ((int64_t)pix_blocks * out.n));
    chunks = max(1, min(min(chunks, ceil_div(out.c, 4)), 65535));
    const int c_per_block = ceil_div(out.c, chunks);
    chunks = ceil_div(out.c, c_per_block);
    dim3 grid(pix_blocks, chunks, out.n);
    block_extractor_fwd_kernel<T><<<grid, 256, 0, st>>>(src, flow, out, k, c_per_block);
    return check_launch("block_extractor_forward");
}

template <typename T>
static int block_extractor_backward_t(const ffwm_tensor4* a, const ffwm_tensor4* b, const ffwm_tensor4* go,
                                      const ffwm_tensor4* ga, const ffwm_tensor4* gb, int k, cudaStream_t st) {
    View<const T> src, flow, gout;
    View<T> gs, gf;
    int rc;
    if ((rc = make_view<const T>(a, "source", &src))) return rc;
    if ((rc = make_view<const T>(b, "flow_field", &flow))) return rc;
    if ((rc = make_view<const T>(go, "grad_output", &gout))) return rc;
    if ((rc = make_view<T>(ga, "grad_source", &gs, true))) return rc;
    if ((rc = make_view<T>(gb, "grad_flow_field", &gf, true))) return rc;
    if (k < 1) { set_error("block_extractor: kernel_size=%d", k); return FFWM_ERR_ARG; }
    if (flow.c != 2) { set_error("block_extractor: flow_field needs 2 channels, got %d", flow.c); return FFWM_ERR_SHAPE; }
    if (gout.n != flow.n || src.n < gout.n || gout.c != src.c ||
        (int64_t)gout.h != (int64_t)k * flow.h || (int64_t)gout.w != (int64_t)k * flow.w) {
        set_error("block_extractor_backward: grad_output (%d,%d,%d,%d) != (B,C,k*Hf,k*Wf)", gout.n, gout.c, gout.h, gout.w);
        return FFWM_ERR_SHAPE;
    }
    if (gs.p && (gs.n != src.n || gs.c != src.c || gs.h != src.h || gs.w != src.w)) {
        set_error("block_extractor_backward: grad_source shape differs from source"); return FFWM_ERR_SHAPE;
    }
    if (gf.p && (gf.n != flow.n || gf.c != 2 || gf.h != flow.h || gf.w != flow.w)) {
        set_error("block_extractor_backward: grad_flow_field shape differs from flow_field"); return FFWM_ERR_SHAPE;
    }
    if ((int64_t)gout.n * gout.h * gout.w == 0) return FFWM_OK;
    if (gout.n > 65535) { set_error("block_extractor: batch %d > 65535", gout.n); return FFWM_ERR_TOO_LARGE; }
    if (src.h == 0 || src.w == 0) { set_error("block_extractor: empty source plane"); return FFWM_ERR_SHAPE; }
    if constexpr (sizeof(T) == 4) {
        if ((k == 2 || k == 3) && gout.c >= 8 && !opt(OPT_DISABLE_TILED)) {
            constexpr int SL = 8, PX = 256 / SL;
            dim3 grid(ceil_div(flow.h * flow.w, PX), 1, gout.n), block(PX, SL);
            if (k == 2) block_extractor_bwd_window_kernel<2, SL><<<grid, block, 0, st>>>(src, flow, gout, gs, gf);
            else block_extractor_bwd_window_kernel<3, SL><<<grid, block, 0, st>>>(src, flow, gout, gs, gf);
            return check_launch("block_extractor_backward(window)");
        }
    }
    const int c = gout.c;
    if (c >= 8) launch_bwd_sl<T, 8>(src, flow, gout, gs, gf, k, st);
    else if (c >= 4) launch_bwd_sl<T, 4>(src, flow, gout, gs, gf, k, st);
    else if (c >= 2) launch_bwd_sl<T, 2>(src, flow, gout, gs, gf, k, st);
    else launch_bwd_sl<T, 1>(src, flow, gout, gs, gf, k, st);
    return check_launch("block_extractor_backward");
}

}  // namespace ffwm

extern "C" int ffwm_block_extractor_forward(const ffwm_tensor4* source, const ffwm_tensor4* flow,
                                            const ffwm_tensor4* output, int kernel_size, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::block_extractor_forward_t<float>(source, flow, output, kernel_size, st);
    if (dtype == FFWM_F64) return ffwm::block_extractor_forward_t<double>(source, flow, output, kernel_size, st);
    ffwm::set_error("block_extractor_forward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}

extern "C" int ffwm_block_extractor_backward(const ffwm_tensor4* source, const ffwm_tensor4* flow,
                                             const ffwm_tensor4* grad_output, const ffwm_tensor4* grad_source,
                                             const ffwm_tensor4* grad_flow, int kernel_size, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32)
        return ffwm::block_extractor_backward_t<float>(source, flow, grad_output, grad_source, grad_flow, kernel_size, st);
    if (dtype == FFWM_F64)
        return ffwm::block_extractor_backward_t<double>(source, flow, grad_output, grad_source, grad_flow, kernel_size, st);
    ffwm::set_error("block_extractor_backward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}
