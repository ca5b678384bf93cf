// resample2d forward on the rolling-strip gather (roll_gather.cuh), kernel_size 2 or 4, dilation 1, fp32
// (a rolling flow-gradient kernel was measured at 1.74 ms against 1.11 ms for gather_quad.cuh and removed).  Included by resample2d.cu after Taps / tap_geometry.
//
// Numerics: the reference evaluates exp() in double and rounds to float; these kernels
// use the single-precision expf (<= 2 ulp) and multiply by a reciprocal of the normaliser instead of
// dividing — both far inside the path's 1e-5 forward / 1e-4 backward tolerance (the direct kernels
// keep the double-precision evaluation).  Tap order of the sums is the reference's.
#pragma once
#include "roll_gather.cuh"

namespace ffwm {

// SAFE_DIV(-d*d, 2*sigma*sigma) then exp, in single precision.
__device__ __forceinline__ float rs_gauss_fast(float d, float sigma) {
    const float num = -d * d, den = 2.f * sigma * sigma;
    return expf(den == 0.f ? num * 1e8f : num / den);
}

// Window position (0 .. 2*HALF-1, left/top to right/bottom) of the reference's tap index:
// index 2f is floor - f, index 2f+1 is floor + f + 1 (resample2d_kernel.cu:62-68).
template <int HALF>
__host__ __device__ constexpr int rs_win_pos(int k) { return (k & 1) ? HALF + (k >> 1) : HALF - 1 - (k >> 1); }

// Per-pixel geometry of the rolling kernels, formed by one lane.
template <int HALF>
struct RsRollGeo {
    static constexpr int N2 = 2 * HALF;
    float wx[N2], wy[N2];      // Gaussian weights, reference tap-index order
    float alpha, beta, sum, sigma;
    int off[N2];               // ring offset of window row i at the window's first column (fast pixels)
    bool fast;
};

template <int HALF>
__device__ __forceinline__ void rs_roll_geometry(float dx, float dy, float sigma, int x, int y, int rx0, int ystep,
                                                 RsRollGeo<HALF>& g) {
    constexpr int N2 = 2 * HALF;
    const float xf = float(x) + dx, yf = float(y) + dy;
    const float fxf = floorf(xf), fyf = floorf(yf);
    g.alpha = xf - fxf;
    g.beta = yf - fyf;
    g.sigma = sigma;
    float dxs[N2], dys[N2];
#pragma unroll
    for (int f = 0; f < HALF; ++f) {
        dxs[2 * f] = float(f) + g.alpha;
        dxs[2 * f + 1] = float(1. + f) - g.alpha;
        dys[2 * f] = float(f) + g.beta;
        dys[2 * f + 1] = float(1. + f) - g.beta;
    }
#pragma unroll
    for (int i = 0; i < N2; ++i) {
        g.wx[i] = rs_gauss_fast(dxs[i], sigma);
        g.wy[i] = rs_gauss_fast(dys[i], sigma);
    }
    float sum = 0.f;
#pragma unroll
    for (int fy = 0; fy < HALF; ++fy)
#pragma unroll
        for (int fx = 0; fx < HALF; ++fx)
            sum += (g.wy[2 * fy] * g.wx[2 * fx] + g.wy[2 * fy] * g.wx[2 * fx + 1] +
                    g.wy[2 * fy + 1] * g.wx[2 * fx] + g.wy[2 * fy + 1] * g.wx[2 * fx + 1]);
    g.sum = sum;
    // the whole window inside the ring of this step?  (float compares: NaN -> slow path)
    g.fast = fxf - float(HALF - 1) >= float(rx0) && fxf + float(HALF) <= float(rx0 + RG_RW - 1) &&
             fyf - float(HALF - 1) >= float(ystep - RG_M) && fyf + float(HALF) <= float(ystep + RG_SH + RG_M - 1);
    if (g.fast) {
        const int cb = int(fxf) - (HALF - 1) - rx0, r0 = int(fyf) - (HALF - 1);
#pragma unroll
        for (int i = 0; i < N2; ++i) g.off[i] = ((r0 + i) & (RG_RING - 1)) * RG_RW + cb;
    } else {
#pragma unroll
        for (int i = 0; i < N2; ++i) g.off[i] = 0;
    }
}

// A pixel whose window left the ring: every lane reads its own channel plane with the reference's
// clamped tap indices.  Returns the un-normalised sum; rare (|displacement| >= 7 px).
template <int HALF>
__device__ __noinline__ float2 rs_roll_fwd_slow(const View<const float>& in1, const View<const float>& in2,
                                                const float* plane_lane, int b, int y, int x) {
    constexpr int N2 = 2 * HALF;
    const float* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
    const float dx = __ldg(f), dy = __ldg(f + in2.sc), sigma = __ldg(f + 2 * in2.sc);
    const float xf = float(x) + dx, yf = float(y) + dy;
    Taps<float, HALF> t;
    tap_geometry<float, HALF>(xf, yf, xf - floorf(xf), yf - floorf(yf), 1, in1.h, in1.w, t);
    float wx[N2], wy[N2];
#pragma unroll
    for (int i = 0; i < N2; ++i) { wx[i] = rs_gauss_fast(t.dxs[i], sigma); wy[i] = rs_gauss_fast(t.dys[i], sigma); }
    float sum = 0.f, val = 0.f;
#pragma unroll
    for (int fy = 0; fy < HALF; ++fy)
#pragma unroll
        for (int fx = 0; fx < HALF; ++fx) {
            sum += (wy[2 * fy] * wx[2 * fx] + wy[2 * fy] * wx[2 * fx + 1] + wy[2 * fy + 1] * wx[2 * fx] + wy[2 * fy + 1] * wx[2 * fx + 1]);
            const float* rt = plane_lane + t.iy[2 * fy] * in1.sh;
            const float* rb = plane_lane + t.iy[2 * fy + 1] * in1.sh;
            val += wy[2 * fy] * wx[2 * fx] * __ldg(rt + t.ix[2 * fx] * in1.sw);
            val += wy[2 * fy] * wx[2 * fx + 1] * __ldg(rt + t.ix[2 * fx + 1] * in1.sw);
            val += wy[2 * fy + 1] * wx[2 * fx] * __ldg(rb + t.ix[2 * fx] * in1.sw);
            val += wy[2 * fy + 1] * wx[2 * fx + 1] * __ldg(rb + t.ix[2 * fx + 1] * in1.sw);
        }
    return make_float2(val, sum == 0.f ? 1e8f : 1.f / sum);
}

// ---------------------------------------------------------------- forward
// Record per pixel (PW floats, 16-byte rows): [0] off0|off1<<16  [1] off2|off3<<16 (HALF=2)
// [2] 1/sum  [3] slow flag  [4..4+N2) wx  [4+N2..4+2*N2) wy
template <int HALF>
struct RsRollFwdRec { static constexpr int N2 = 2 * HALF, PW = 4 + 2 * N2; };

template <int HALF>
__global__ void __launch_bounds__(RG_THREADS, 1)
resample2d_fwd_roll_kernel(View<const float> in1, View<const float> in2, View<float> out) {
    using RR = RsRollFwdRec<HALF>;
    constexpr int N2 = RR::N2, PW = RR::PW;
    extern __shared__ __align__(16) unsigned char rg_smem_raw[];
    float* slab = reinterpret_cast<float*>(rg_smem_raw);                       // [32][1025]
    float* prm_all = slab + 32 * RG_CHP;                                       // [16 warps][32 px][PW]
    float* stage_all = prm_all + RG_WARPS * 32 * PW;                           // [16 warps][32][9]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * RG_SW, c0 = blockIdx.y * 32, b = blockIdx.z;
    const int nch = min(32, out.c - c0);
    const int rx0 = x0 - RG_M;
    const int nsteps = (out.h + RG_SH - 1) / RG_SH;
    const int wrow = warp >> 1, xw0 = x0 + (warp & 1) * RG_PXW;                // this warp's row in the step / first column
    float* prm = prm_all + warp * (32 * PW);
    float* stage = stage_all + warp * (32 * RG_SPITCH);
    const float* slab_lane = slab + lane * RG_CHP;
    const float* plane_lane = in1.p + b * in1.sb + (int64_t)(c0 + min(lane, nch - 1)) * in1.sc;

    rg_fill_rows<false>(slab, in1, b, c0, nch, rx0, -RG_M, RG_M + 2 * RG_SH, warp, lane);   // rows [-8, 16)

    // geometry lane mapping: lane -> (step inside the block, pixel)
    const int gsub = lane >> 3, gx = xw0 + (lane & 7);
    float ndx = 0.f, ndy = 0.f, nsg = 1.f;
    auto load_flow = [&](int s_base) {
        const int y = (s_base + gsub) * RG_SH + wrow;
        ndx = ndy = 0.f; nsg = 1.f;
        if (y < out.h && gx < out.w) {
            const float* f = in2.p + b * in2.sb + y * in2.sh + gx * in2.sw;
            ndx = ld_stream(f); ndy = ld_stream(f + in2.sc); nsg = ld_stream(f + 2 * in2.sc);
        }
    };
    load_flow(0);

    for (int s = 0; s < nsteps; ++s) {
        if ((s & (RG_BLK - 1)) == 0) {
            __syncwarp();
            RsRollGeo<HALF> g;
            const int ystep = (s + gsub) * RG_SH;
            rs_roll_geometry<HALF>(ndx, ndy, nsg, gx, ystep + wrow, rx0, ystep, g);
            float* P = prm + lane * PW;
            int* Pi = reinterpret_cast<int*>(P);
            Pi[0] = g.off[0] | (g.off[1] << 16);
            if (HALF == 2) Pi[1] = g.off[N2 - 2] | (g.off[N2 - 1] << 16);
            P[2] = g.sum == 0.f ? 1e8f : 1.f / g.sum;
            Pi[3] = g.fast ? 0 : 1;
#pragma unroll
            for (int i = 0; i < N2; ++i) { P[4 + i] = g.wx[i]; P[4 + N2 + i] = g.wy[i]; }
            load_flow(s + RG_BLK);
            __syncwarp();
        }
        rg_cp_async_wait_all();
        __syncthreads();
        if (s + 1 < nsteps) rg_fill_rows<false>(slab, in1, b, c0, nch, rx0, (s + 1) * RG_SH + RG_M, RG_SH, warp, lane);
        const int y = s * RG_SH + wrow;
        if (y < out.h) {                                     // warp-uniform
#pragma unroll 2
            for (int px = 0; px < RG_PXW; ++px) {
                if (xw0 + px >= out.w) break;                // warp-uniform
                const float4* P4 = reinterpret_cast<const float4*>(prm + ((s & (RG_BLK - 1)) * RG_PXW + px) * PW);
                const float4 h = P4[0];
                float w[2 * N2];
#pragma unroll
                for (int q = 0; q < (2 * N2) / 4; ++q) {
                    const float4 v = P4[1 + q];
                    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
                const float* wx = w;
                const float* wy = w + N2;
                float val = 0.f, inv = h.z;
                if (__float_as_int(h.w) == 0) {
                    const float* rowp[N2];
                    rowp[0] = slab_lane + (__float_as_int(h.x) & 0xffff);
                    rowp[1] = slab_lane + (__float_as_int(h.x) >> 16);
                    if (HALF == 2) {
                        rowp[N2 - 2] = slab_lane + (__float_as_int(h.y) & 0xffff);
                        rowp[N2 - 1] = slab_lane + (__float_as_int(h.y) >> 16);
                    }
#pragma unroll
                    for (int fy = 0; fy < HALF; ++fy)
#pragma unroll
                        for (int fx = 0; fx < HALF; ++fx) {
                            const float* rt = rowp[rs_win_pos<HALF>(2 * fy)];
                            const float* rb = rowp[rs_win_pos<HALF>(2 * fy + 1)];
                            val += wy[2 * fy] * wx[2 * fx] * rt[rs_win_pos<HALF>(2 * fx)];
                            val += wy[2 * fy] * wx[2 * fx + 1] * rt[rs_win_pos<HALF>(2 * fx + 1)];
                            val += wy[2 * fy + 1] * wx[2 * fx] * rb[rs_win_pos<HALF>(2 * fx)];
                            val += wy[2 * fy + 1] * wx[2 * fx + 1] * rb[rs_win_pos<HALF>(2 * fx + 1)];
                        }
                } else {
                    const float2 sv = rs_roll_fwd_slow<HALF>(in1, in2, plane_lane, b, y, xw0 + px);
                    val = sv.x;
                    inv = sv.y;
                }
                stage[lane * RG_SPITCH + px] = val * inv;
            }
            __syncwarp();
            rg_store_row(stage, out, b, c0, nch, y, xw0, lane);
            __syncwarp();
        }
    }
}

template <int HALF>
static int launch_fwd_roll(const View<const float>& in1, const View<const float>& in2, const View<float>& out, cudaStream_t st) {
    using RR = RsRollFwdRec<HALF>;
    const size_t smem = sizeof(float) * (32 * RG_CHP + RG_WARPS * 32 * RR::PW + RG_WARPS * 32 * RG_SPITCH);
    cudaError_t e = cudaFuncSetAttribute(resample2d_fwd_roll_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("resample2d_fwd_roll: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(ceil_div(out.w, RG_SW), ceil_div(out.c, 32), out.n);
    resample2d_fwd_roll_kernel<HALF><<<grid, RG_THREADS, smem, st>>>(in1, in2, out);
    return FFWM_OK;
}

// ---------------------------------------------------------------- flow-gradient coefficients (gather_quad policy)
// the derived coefficient tables of one pixel (reference tap-index order)
template <int HALF>
struct RsGradCoef {
    static constexpr int N2 = 2 * HALF;
    float ax[N2], ay[N2], bx[N2], by[N2];
};
template <int HALF>
__device__ __forceinline__ void rs_grad_coef(float alpha, float beta, const float* wx, const float* wy, RsGradCoef<HALF>& k) {
#pragma unroll
    for (int f = 0; f < HALF; ++f) {
        const float dl = float(f) + alpha, dr = float(1. + f) - alpha;
        const float dt = float(f) + beta, db = float(1. + f) - beta;
        k.ax[2 * f] = dl * wx[2 * f];          k.ax[2 * f + 1] = -dr * wx[2 * f + 1];
        k.ay[2 * f] = dt * wy[2 * f];          k.ay[2 * f + 1] = -db * wy[2 * f + 1];
        k.bx[2 * f] = dl * dl * wx[2 * f];     k.bx[2 * f + 1] = dr * dr * wx[2 * f + 1];
        k.by[2 * f] = dt * dt * wy[2 * f];     k.by[2 * f + 1] = db * db * wy[2 * f + 1];
    }
}

}  // namespace ffwm
