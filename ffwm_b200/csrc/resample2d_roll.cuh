// resample2d on the rolling-strip gather (roll_gather.cuh): forward and flow gradient, kernel_size
// 2 or 4, dilation 1, fp32.  Included by resample2d.cu after Taps / tap_geometry.
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

// ---------------------------------------------------------------- flow gradient (K3)
// CTA = (strip, 128-row segment, batch); it walks the segment once per group of 32 channels.  Each
// lane accumulates the four partial sums of the direct kernel (A0, A1, A2, Bs) for a pixel; a packed
// butterfly sums them over the lanes (channels), the per-pixel totals of the channel groups are
// added up in warp-private shared memory, and the lanes 0..7 finish the 8 pixels of the step while
// the last group passes.  Deterministic, one store per output element.
// Record per pixel: [0] off0|off1<<16 [1] off2|off3<<16 [2] slow flag [3] sigma [4] alpha [5] beta
// [6] sum [7] - [8..8+N2) wx [8+N2..8+2*N2) wy
template <int HALF>
struct RsRollGradRec { static constexpr int N2 = 2 * HALF, PW = 8 + 2 * N2; };

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

template <int HALF>
__device__ __noinline__ float4 rs_roll_grad_slow(const View<const float>& in1, const View<const float>& in2,
                                                 const float* plane_lane, int b, int y, int x, float g) {
    constexpr int N2 = 2 * HALF;
    const float* f = in2.p + b * in2.sb + y * in2.sh + x * in2.sw;
    const float dx = __ldg(f), dy = __ldg(f + in2.sc), sigma = __ldg(f + 2 * in2.sc);
    const float xf = float(x) + dx, yf = float(y) + dy;
    const float alpha = xf - floorf(xf), beta = yf - floorf(yf);
    Taps<float, HALF> t;
    tap_geometry<float, HALF>(xf, yf, alpha, beta, 1, in1.h, in1.w, t);
    float wx[N2], wy[N2];
#pragma unroll
    for (int i = 0; i < N2; ++i) { wx[i] = rs_gauss_fast(t.dxs[i], sigma); wy[i] = rs_gauss_fast(t.dys[i], sigma); }
    RsGradCoef<HALF> k;
    rs_grad_coef<HALF>(alpha, beta, wx, wy, k);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, bs = 0.f;
#pragma unroll 1
    for (int i = 0; i < N2; ++i) {
        float r0 = 0.f, rB = 0.f, r2 = 0.f;
#pragma unroll
        for (int j = 0; j < N2; ++j) {
            const float v = __ldg(plane_lane + t.iy[i] * in1.sh + t.ix[j] * in1.sw);
            r0 += k.ax[j] * v;
            rB += wx[j] * v;
            r2 += k.bx[j] * v;
        }
        a0 += wy[i] * r0;
        a1 += k.ay[i] * rB;
        a2 += k.by[i] * rB + wy[i] * r2;
        bs += wy[i] * rB;
    }
    return make_float4(a0 * g, a1 * g, a2 * g, bs * g);
}

template <int HALF>
__global__ void __launch_bounds__(RG_THREADS, 1)
resample2d_gflow_roll_kernel(View<const float> in1, View<const float> in2, View<const float> gout, View<float> gin2, int seg_rows) {
    using RR = RsRollGradRec<HALF>;
    constexpr int N2 = RR::N2, PW = RR::PW;
    constexpr int SEG_STEPS = RG_SEG / RG_SH;
    extern __shared__ __align__(16) unsigned char rg_smem_raw[];
    float* slab = reinterpret_cast<float*>(rg_smem_raw);                       // [32][1025]
    float* prm_all = slab + 32 * RG_CHP;                                       // [16 warps][32 px][PW]
    float* stage_all = prm_all + RG_WARPS * 32 * PW;                           // [16 warps][32][9]  grad_output
    float* accs_all = stage_all + RG_WARPS * 32 * RG_SPITCH;                   // [16 warps][16 steps][8 px][4]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * RG_SW, yseg = blockIdx.y * seg_rows, b = blockIdx.z;
    const int rx0 = x0 - RG_M;
    const int seg_h = min(seg_rows, gout.h - yseg);
    const int nsteps = (seg_h + RG_SH - 1) / RG_SH;
    const int wrow = warp >> 1, xw0 = x0 + (warp & 1) * RG_PXW;
    float* prm = prm_all + warp * (32 * PW);
    float* stage = stage_all + warp * (32 * RG_SPITCH);
    float* acc = accs_all + warp * (SEG_STEPS * RG_PXW * 4);
    const float* slab_lane = slab + lane * RG_CHP;
    const int gsub = lane >> 3, gx = xw0 + (lane & 7);

    for (int i = lane; i < SEG_STEPS * RG_PXW * 4; i += 32) acc[i] = 0.f;

    for (int c0 = 0; c0 < gout.c; c0 += 32) {
        const int nch = min(32, gout.c - c0);
        const bool last_group = c0 + 32 >= gout.c;
        const float* plane_lane = in1.p + b * in1.sb + (int64_t)(c0 + min(lane, nch - 1)) * in1.sc;
        __syncthreads();                                   // the previous group's last step is done with the ring
        rg_fill_rows<false>(slab, in1, b, c0, nch, rx0, yseg - RG_M, RG_M + 2 * RG_SH, warp, lane);

        float ndx = 0.f, ndy = 0.f, nsg = 1.f;
        auto load_flow = [&](int s_base) {
            const int y = yseg + (s_base + gsub) * RG_SH + wrow;
            ndx = ndy = 0.f; nsg = 1.f;
            if (s_base + gsub < nsteps && y < gout.h && gx < gout.w) {
                const float* f = in2.p + b * in2.sb + y * in2.sh + gx * in2.sw;
                ndx = __ldg(f); ndy = __ldg(f + in2.sc); nsg = __ldg(f + 2 * in2.sc);
            }
        };
        load_flow(0);
        float gl[8];
        rg_load_row(gl, gout, b, c0, nch, yseg + wrow, xw0, lane);

        for (int s = 0; s < nsteps; ++s) {
            if ((s & (RG_BLK - 1)) == 0) {
                __syncwarp();
                RsRollGeo<HALF> g;
                const int ystep = yseg + (s + gsub) * RG_SH;
                rs_roll_geometry<HALF>(ndx, ndy, nsg, gx, ystep + wrow, rx0, ystep, g);
                float* P = prm + lane * PW;
                int* Pi = reinterpret_cast<int*>(P);
                Pi[0] = g.off[0] | (g.off[1] << 16);
                if (HALF == 2) Pi[1] = g.off[N2 - 2] | (g.off[N2 - 1] << 16);
                Pi[2] = g.fast ? 0 : 1;
                P[3] = g.sigma; P[4] = g.alpha; P[5] = g.beta; P[6] = g.sum;
#pragma unroll
                for (int i = 0; i < N2; ++i) { P[8 + i] = g.wx[i]; P[8 + N2 + i] = g.wy[i]; }
                load_flow(s + RG_BLK);
                __syncwarp();
            }
            rg_cp_async_wait_all();
            __syncthreads();
            if (s + 1 < nsteps) rg_fill_rows<false>(slab, in1, b, c0, nch, rx0, yseg + (s + 1) * RG_SH + RG_M, RG_SH, warp, lane);
            const int y = yseg + s * RG_SH + wrow;
            // grad_output of this step: registers -> warp-private staging; next step's row -> registers
            rg_stage_row(stage, gl, lane);
            __syncwarp();
            if (s + 1 < nsteps) rg_load_row(gl, gout, b, c0, nch, y + RG_SH, xw0, lane);
            if (y < gout.h) {                                // warp-uniform
#pragma unroll 1
                for (int p4 = 0; p4 < RG_PXW / 4; ++p4) {
                    float v[16];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int px = p4 * 4 + kk;
                        float r4[4] = {0.f, 0.f, 0.f, 0.f};
                        if (xw0 + px < gout.w) {             // warp-uniform
                            const float4* P4 = reinterpret_cast<const float4*>(prm + ((s & (RG_BLK - 1)) * RG_PXW + px) * PW);
                            const float4 h = P4[0], h2 = P4[1];
                            const float g = stage[lane * RG_SPITCH + px];
                            if (__float_as_int(h.z) == 0) {
                                float w[2 * N2];
#pragma unroll
                                for (int q = 0; q < (2 * N2) / 4; ++q) {
                                    const float4 u = P4[2 + q];
                                    w[4 * q] = u.x; w[4 * q + 1] = u.y; w[4 * q + 2] = u.z; w[4 * q + 3] = u.w;
                                }
                                const float* wx = w;
                                const float* wy = w + N2;
                                RsGradCoef<HALF> k;
                                rs_grad_coef<HALF>(h2.x, h2.y, wx, wy, k);
                                const float* rowp[N2];
                                rowp[0] = slab_lane + (__float_as_int(h.x) & 0xffff);
                                rowp[1] = slab_lane + (__float_as_int(h.x) >> 16);
                                if (HALF == 2) {
                                    rowp[N2 - 2] = slab_lane + (__float_as_int(h.y) & 0xffff);
                                    rowp[N2 - 1] = slab_lane + (__float_as_int(h.y) >> 16);
                                }
#pragma unroll
                                for (int i = 0; i < N2; ++i) {
                                    const float* rp = rowp[rs_win_pos<HALF>(i)];
                                    float r0 = 0.f, rB = 0.f, r2 = 0.f;
#pragma unroll
                                    for (int j = 0; j < N2; ++j) {
                                        const float sv = rp[rs_win_pos<HALF>(j)];
                                        r0 += k.ax[j] * sv;
                                        rB += wx[j] * sv;
                                        r2 += k.bx[j] * sv;
                                    }
                                    r4[0] += wy[i] * r0;
                                    r4[1] += k.ay[i] * rB;
                                    r4[2] += k.by[i] * rB + wy[i] * r2;
                                    r4[3] += wy[i] * rB;
                                }
                                r4[0] *= g; r4[1] *= g; r4[2] *= g; r4[3] *= g;
                            } else {
                                const float4 sv = rs_roll_grad_slow<HALF>(in1, in2, plane_lane, b, y, xw0 + px, g);
                                r4[0] = sv.x; r4[1] = sv.y; r4[2] = sv.z; r4[3] = sv.w;
                            }
                        }
                        v[4 * kk] = r4[0]; v[4 * kk + 1] = r4[1]; v[4 * kk + 2] = r4[2]; v[4 * kk + 3] = r4[3];
                    }
                    const float tot = gt_packed_reduce<16>(v, lane);       // lane l: value index l >> 1
                    if ((lane & 1) == 0) acc[(s * RG_PXW + p4 * 4) * 4 + (lane >> 1)] += tot;
                }
                if (last_group) {
                    __syncwarp();
                    const int x = xw0 + lane;
                    if (lane < RG_PXW && x < gout.w) {
                        const float* A = acc + (s * RG_PXW + lane) * 4;
                        const float a0 = A[0], a1 = A[1], a2 = A[2], bs = A[3];
                        const float* P = prm + ((s & (RG_BLK - 1)) * RG_PXW + lane) * PW;
                        const float sigma = P[3], alpha = P[4], beta = P[5], sum = P[6];
                        float wx[N2], wy[N2];
#pragma unroll
                        for (int i = 0; i < N2; ++i) { wx[i] = P[8 + i]; wy[i] = P[8 + N2 + i]; }
                        RsGradCoef<HALF> k;
                        rs_grad_coef<HALF>(alpha, beta, wx, wy, k);
                        float Wx = 0.f, Wy = 0.f, AX = 0.f, AY = 0.f, BX = 0.f, BY = 0.f;
#pragma unroll
                        for (int i = 0; i < N2; ++i) {
                            Wx += wx[i]; Wy += wy[i];
                            AX += k.ax[i]; AY += k.ay[i];
                            BX += k.bx[i]; BY += k.by[i];
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
                }
            }
            __syncwarp();
        }
    }
}

template <int HALF>
static int launch_gflow_roll(const View<const float>& in1, const View<const float>& in2, const View<const float>& gout,
                             const View<float>& g2, cudaStream_t st) {
    using RR = RsRollGradRec<HALF>;
    const size_t smem = sizeof(float) * (32 * RG_CHP + RG_WARPS * 32 * RR::PW + RG_WARPS * 32 * RG_SPITCH +
                                         RG_WARPS * (RG_SEG / RG_SH) * RG_PXW * 4);
    cudaError_t e = cudaFuncSetAttribute(resample2d_gflow_roll_kernel<HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("resample2d_gflow_roll: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    const int seg = roll_segment_rows(gout.n, gout.h, gout.w);
    dim3 grid(ceil_div(gout.w, RG_SW), ceil_div(gout.h, seg), gout.n);
    resample2d_gflow_roll_kernel<HALF><<<grid, RG_THREADS, smem, st>>>(in1, in2, gout, g2, seg);
    return FFWM_OK;
}

}  // namespace ffwm
