// conv_few: direct fp32 convolutions for the layers of the path with a DEGENERATE channel count — at most 4 input channels
// (LightCNN's 5x5 stem on the grey image, lightcnn/light_cnn.py:96-100; the generator's 7x7 stem, the FlowNets' conv0, VGG's
// conv1_1 and the discriminators' first convolution on RGB, models/base_networks.py:59-75,230,397-399, models/losses.py:430)
// or at most 4 output channels at stride 1 (the data gradient of LightCNN's stem; the 2-channel flow heads and 3-channel
// reconstructions, models/base_networks.py:45-57,241).  On the tensor cores those layers pad K or N to a tile that is 75-97 %
// zeros and still pay the full operand staging: the stem and its data gradient ran at 5-15 TFLOP/s, 331 us per launch at
// batch 16 (profiles/r02z_launches_train_summary.txt: conv_gen_tc_kernel<1, 2> grid 2048).  Here they are plain FFMA kernels
// (exact fp32 accumulation) whose cost is the output write (few inputs) or the input read (few outputs):
//
//   few inputs   block = 32 x 8 output pixels x 32 output channels; the input tile (all <= 4 channels, with halo) and the
//                [taps][32] weight slice live in shared memory; a thread keeps 32 accumulators, reads each input value once
//                per tap (conflict-free LDS) and the 32 weights of that tap as 8 broadcast LDS.128.
//   few outputs  block = 128 x 8 output pixels, thread = a strip of 4 pixels; input channels arrive in chunks of 8 through
//                shared memory; per (channel, kernel row) a thread loads its 4 + kw - 1 input values with 128-bit LDS and
//                the kw x Cout weights as broadcasts: 0.65 LDS per tap instead of 2.
// Weights are read through (in_major, flip): the same kernels serve forward passes (weight[out][in], taps as stored) and data
// gradients (weight[in][out], taps reversed) without materialising a flipped copy.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace ffwm {

struct FewGeo {
    int n, ci, co, hi, wi, ho, wo, kh, kw, stride, pad;
    int64_t w_o, w_i, w_ky, w_kx;      // element strides of the LOGICAL weight (out, in, ky, kx); w_base: offset of its first tap
    int64_t w_base;
};

__device__ __forceinline__ float few_weight(const float* __restrict__ w, const FewGeo& g, int o, int i, int ky, int kx) {
    return __ldg(w + g.w_base + o * g.w_o + i * g.w_i + ky * g.w_ky + kx * g.w_kx);
}

// ---------------------------------------------------------------- at most 4 input channels
constexpr int FI_TW = 32, FI_TH = 8, FI_CO = 32;

__global__ void __launch_bounds__(FI_TW * FI_TH)
conv_few_in_kernel(View<const float> x, const float* __restrict__ w, const float* __restrict__ bias, View<float> out, const __grid_constant__ FewGeo g) {
    extern __shared__ __align__(16) float few_smem[];
    const int K = g.ci * g.kh * g.kw;
    const int th = (FI_TH - 1) * g.stride + g.kh, tw = (FI_TW - 1) * g.stride + g.kw, tpitch = tw | 1;
    float* sw = few_smem;                                  // [K][32]
    float* sx = few_smem + K * FI_CO;                      // [ci][th][tpitch]
    const int cog = (g.co + FI_CO - 1) / FI_CO;
    const int n = blockIdx.z / cog, co0 = (blockIdx.z % cog) * FI_CO;
    const int ox0 = blockIdx.x * FI_TW, oy0 = blockIdx.y * FI_TH;
    const int tid = threadIdx.y * FI_TW + threadIdx.x;
    for (int e = tid; e < K * FI_CO; e += FI_TW * FI_TH) {
        const int c = e % FI_CO, k = e / FI_CO;
        const int kx = k % g.kw, ky = (k / g.kw) % g.kh, i = k / (g.kw * g.kh);
        sw[e] = co0 + c < g.co ? few_weight(w, g, co0 + c, i, ky, kx) : 0.f;
    }
    const int iy0 = oy0 * g.stride - g.pad, ix0 = ox0 * g.stride - g.pad;
    for (int e = tid; e < g.ci * th * tw; e += FI_TW * FI_TH) {
        const int xx = e % tw, yy = (e / tw) % th, i = e / (tw * th);
        const int iy = iy0 + yy, ix = ix0 + xx;
        float v = 0.f;
        if ((unsigned)iy < (unsigned)g.hi && (unsigned)ix < (unsigned)g.wi) v = __ldg(x.p + n * x.sb + i * x.sc + (int64_t)iy * x.sh + (int64_t)ix * x.sw);
        sx[(i * th + yy) * tpitch + xx] = v;
    }
    __syncthreads();
    float acc[FI_CO];
#pragma unroll
    for (int c = 0; c < FI_CO; ++c) acc[c] = 0.f;
    const float* xp = sx + (threadIdx.y * g.stride) * tpitch + threadIdx.x * g.stride;
    int k = 0;
    for (int i = 0; i < g.ci; ++i)
        for (int ky = 0; ky < g.kh; ++ky)
            for (int kx = 0; kx < g.kw; ++kx, ++k) {
                const float v = xp[(i * th + ky) * tpitch + kx];
                const float4* wp = reinterpret_cast<const float4*>(sw + k * FI_CO);
#pragma unroll
                for (int q = 0; q < FI_CO / 4; ++q) {
                    const float4 w4 = wp[q];
                    acc[4 * q] = fmaf(v, w4.x, acc[4 * q]);
                    acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
                }
            }
    const int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
    if (oy < g.ho && ox < g.wo) {
        float* op = out.p + n * out.sb + (int64_t)oy * out.sh + (int64_t)ox * out.sw;
#pragma unroll
        for (int c = 0; c < FI_CO; ++c)
            if (co0 + c < g.co) op[(co0 + c) * out.sc] = acc[c] + (bias ? __ldg(bias + co0 + c) : 0.f);
    }
}

// ---------------------------------------------------------------- at most 4 output channels, stride 1
constexpr int FO_TW = 128, FO_TH = 8, FO_CI = 8, FO_MAXCO = 4;

template <int CO>
__global__ void __launch_bounds__(FO_TW / 4 * FO_TH)
conv_few_out_kernel(View<const float> x, const float* __restrict__ w, const float* __restrict__ bias, View<float> out, const __grid_constant__ FewGeo g) {
    extern __shared__ __align__(16) float few_smem[];
    const int th = FO_TH + g.kh - 1, tw = FO_TW + g.kw - 1, tpitch = (tw + 3) & ~3;
    const int taps = g.kh * g.kw;
    float* sx = few_smem;                                  // [8 ci][th][tpitch]
    float* sw = few_smem + FO_CI * th * tpitch;            // [8 ci][taps][4]
    const int n = blockIdx.z, ox0 = blockIdx.x * FO_TW, oy0 = blockIdx.y * FO_TH;
    const int tid = threadIdx.y * (FO_TW / 4) + threadIdx.x, nthr = FO_TW / 4 * FO_TH;
    const int iy0 = oy0 - g.pad, ix0 = ox0 - g.pad;
    float acc[CO][4];
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[o][p] = 0.f;
    for (int c0 = 0; c0 < g.ci; c0 += FO_CI) {
        const int nc = min(FO_CI, g.ci - c0);
        __syncthreads();                                   // the previous chunk has been consumed
        // one warp per tile row, lanes along the row: one division per row instead of three per element (the first version's
        // index arithmetic cost four times the FFMAs of the chunk)
        for (int r = tid >> 5; r < nc * th; r += nthr >> 5) {
            const int c = r / th, yy = r - c * th, iy = iy0 + yy;
            const bool rowok = (unsigned)iy < (unsigned)g.hi;
            const float* xr = x.p + n * x.sb + (int64_t)(c0 + c) * x.sc + (int64_t)iy * x.sh;
            float* sr = sx + r * tpitch;
            for (int xx = tid & 31; xx < tpitch; xx += 32) {
                const int ix = ix0 + xx;
                sr[xx] = (rowok && xx < tw && (unsigned)ix < (unsigned)g.wi) ? __ldg(xr + (int64_t)ix * x.sw) : 0.f;
            }
        }
        for (int e = tid; e < nc * taps * 4; e += nthr) {
            const int o = e & 3, t = (e >> 2) % taps, c = (e >> 2) / taps;
            sw[e] = o < g.co ? few_weight(w, g, o, c0 + c, t / g.kw, t % g.kw) : 0.f;
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c)
            for (int ky = 0; ky < g.kh; ++ky) {
                // the strip's 4 + kw - 1 (<= 10) input values of this row: three aligned 128-bit loads
                const float4* rp = reinterpret_cast<const float4*>(sx + (c * th + threadIdx.y + ky) * tpitch + 4 * threadIdx.x);
                const float4 r0 = rp[0], r1 = rp[1];
                float v[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, 0.f, 0.f, 0.f, 0.f};
                if (g.kw > 5) { const float4 r2 = rp[2]; v[8] = r2.x; v[9] = r2.y; v[10] = r2.z; v[11] = r2.w; }
                const float4* wp = reinterpret_cast<const float4*>(sw + (c * taps + ky * g.kw) * 4);
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    if (kx < g.kw) {
                        const float4 w4 = wp[kx];
                        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int o = 0; o < CO; ++o)
#pragma unroll
                            for (int p = 0; p < 4; ++p) acc[o][p] = fmaf(v[kx + p], wv[o], acc[o][p]);
                    }
                }
            }
    }
    const int oy = oy0 + threadIdx.y, ox = ox0 + 4 * threadIdx.x;
    if (oy < g.ho) {
#pragma unroll
        for (int o = 0; o < CO; ++o) {
            const float b = bias ? __ldg(bias + o) : 0.f;
            float* op = out.p + n * out.sb + o * out.sc + (int64_t)oy * out.sh + (int64_t)ox * out.sw;
#pragma unroll
            for (int p = 0; p < 4; ++p)
                if (ox + p < g.wo) op[p * out.sw] = acc[o][p] + b;
        }
    }
}

}  // namespace ffwm

// out = conv2d(x, W', bias, stride, pad) for W' with at most 4 input channels, or at most 4 output channels at stride 1
// (returns FFWM_ERR_ARG otherwise: the caller keeps the tensor-core path).  W' is `weight` read through (in_major, flip):
//   in_major = 0: W'[o][i] = weight[o][i];  in_major = 1: W'[o][i] = weight[i][o];  flip: taps reversed (ky -> kh-1-ky, kx -> kw-1-kx)
// so that the data gradient of a stride-1 convolution is ffwm_conv_few(grad_out, weight, 1, 1, NULL, grad_in, 1, k-1-pad).
// Exact fp32 FFMA accumulation; any tensor strides; kernels up to 7x7; stride 1 or 2 (few inputs only).
extern "C" int ffwm_conv_few(const ffwm_tensor4* x, const ffwm_tensor4* weight, int in_major, int flip, const float* bias, const ffwm_tensor4* out,
                             int stride, int pad, void* stream) {
    using namespace ffwm;
    View<const float> xv;
    View<float> ov;
    int rc;
    if ((rc = make_view<const float>(x, "x", &xv))) return rc;
    if ((rc = make_view<float>(out, "out", &ov))) return rc;
    if (!weight || !weight->data) { set_error("conv_few: null weight"); return FFWM_ERR_NULL; }
    FewGeo g;
    g.kh = (int)weight->size[2], g.kw = (int)weight->size[3];
    g.co = (int)weight->size[in_major ? 1 : 0], g.ci = (int)weight->size[in_major ? 0 : 1];
    g.n = xv.n, g.hi = xv.h, g.wi = xv.w, g.ho = ov.h, g.wo = ov.w, g.stride = stride, g.pad = pad;
    if (xv.c != g.ci || ov.c != g.co || ov.n != xv.n || g.kh < 1 || g.kw < 1 || g.kh > 7 || g.kw > 7 || (stride != 1 && stride != 2) || pad < 0 ||
        g.ho != (g.hi + 2 * pad - g.kh) / stride + 1 || g.wo != (g.wi + 2 * pad - g.kw) / stride + 1 || g.ho < 1 || g.wo < 1) {
        set_error("conv_few: inconsistent shapes (x %dx%dx%dx%d, weight %dx%dx%dx%d, out %dx%dx%dx%d, stride %d, pad %d)", xv.n, xv.c, xv.h, xv.w,
                  g.co, g.ci, g.kh, g.kw, ov.n, ov.c, ov.h, ov.w, stride, pad);
        return FFWM_ERR_SHAPE;
    }
    g.w_o = weight->stride[in_major ? 1 : 0], g.w_i = weight->stride[in_major ? 0 : 1];
    g.w_ky = weight->stride[2], g.w_kx = weight->stride[3], g.w_base = 0;
    if (flip) {
        g.w_base = (g.kh - 1) * g.w_ky + (g.kw - 1) * g.w_kx;
        g.w_ky = -g.w_ky, g.w_kx = -g.w_kx;
    }
    if ((int64_t)ov.n * ov.c * ov.h * ov.w == 0) return FFWM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* wp = static_cast<const float*>(weight->data);
    if (g.ci <= 4) {
        const int K = g.ci * g.kh * g.kw;
        const int th = (FI_TH - 1) * stride + g.kh, tw = (FI_TW - 1) * stride + g.kw;
        const size_t smem = sizeof(float) * ((size_t)K * FI_CO + (size_t)g.ci * th * (tw | 1));
        const int cog = (g.co + FI_CO - 1) / FI_CO;
        if ((int64_t)g.n * cog > 65535) { set_error("conv_few: grid too large"); return FFWM_ERR_TOO_LARGE; }
        cudaError_t e = cudaFuncSetAttribute(conv_few_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (e != cudaSuccess || smem > 96 * 1024) { set_error("conv_few: shared memory (%zu bytes)", smem); return FFWM_ERR_TOO_LARGE; }
        conv_few_in_kernel<<<dim3((g.wo + FI_TW - 1) / FI_TW, (g.ho + FI_TH - 1) / FI_TH, g.n * cog), dim3(FI_TW, FI_TH), smem, st>>>(xv, wp, bias, ov, g);
        return check_launch("conv_few (few inputs)");
    }
    if (g.co <= FO_MAXCO && stride == 1) {
        const int th = FO_TH + g.kh - 1, tw = FO_TW + g.kw - 1, tpitch = (tw + 3) & ~3;
        const size_t smem = sizeof(float) * ((size_t)FO_CI * th * tpitch + (size_t)FO_CI * g.kh * g.kw * 4);
        if (g.n > 65535) { set_error("conv_few: batch too large"); return FFWM_ERR_TOO_LARGE; }
        const dim3 grid((g.wo + FO_TW - 1) / FO_TW, (g.ho + FO_TH - 1) / FO_TH, g.n), block(FO_TW / 4, FO_TH);
#define FFWM_FEW_OUT(CO)                                                                                              \
        {                                                                                                             \
            cudaError_t e = cudaFuncSetAttribute(conv_few_out_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
            if (e != cudaSuccess || smem > 96 * 1024) { set_error("conv_few: shared memory (%zu bytes)", smem); return FFWM_ERR_TOO_LARGE; } \
            conv_few_out_kernel<CO><<<grid, block, smem, st>>>(xv, wp, bias, ov, g);                                  \
        }
        switch (g.co) { case 1: FFWM_FEW_OUT(1) break; case 2: FFWM_FEW_OUT(2) break; case 3: FFWM_FEW_OUT(3) break; default: FFWM_FEW_OUT(4) break; }
#undef FFWM_FEW_OUT
        return check_launch("conv_few (few outputs)");
    }
    set_error("conv_few: needs <= 4 input channels, or <= 4 output channels at stride 1 (got %d -> %d, stride %d)", g.ci, g.co, stride);
    return FFWM_ERR_ARG;
}
