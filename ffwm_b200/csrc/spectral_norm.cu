// spectral_norm: torch.nn.utils.spectral_norm (one power iteration, dim 0) for ALL layers of a network in three
// launches forward and two backward.
//
// The reference wraps every convolution of the generator and the discriminator in spectral_norm
// (models/base_networks.py:204-246 `_maybe_sn`, :397-410): before each forward pass of a layer,
//     v = normalize(W^T u), u = normalize(W v)         (no_grad, buffers updated in place)
//     sigma = u . (W v),  weight = W / sigma            (W = weight_orig.reshape(out, -1); differentiable in W only)
// — per layer a dozen small library kernels and as many again in the backward pass.  Batched per weight SHAPE
// (ffwm_b200/spectral.py, round 2) that still is ~28 launches per shape group: ~780 of the train step's 3 900 launches.
// Here a device-resident table describes all L layers (pointers to weight_orig / u / v, sizes, offsets into flat scratch
// buffers) and every kernel runs over the blocks of all layers at once:
//   forward   sn_wtu_kernel    t = W^T u                       block = (layer, 32 columns), rows interleaved over 8 warps
//             sn_wv_kernel     v = t / max(|t|, eps), s = W v  block = (layer, 8 rows), |t| re-reduced per block (L2 hits)
//             sn_scale_kernel  u = s / max(|s|, eps), sigma = |s|^2 / max(|s|, eps), out = W / sigma   block = (layer, 4096 elements)
//   backward  sn_bwd_dot_kernel    partial sums of r = sum(g * W) per block
//             sn_bwd_apply_kernel  dW = g / sigma - (r / sigma^2) u v^T        (u, v: the values the forward used, saved per call)
// `update = 0` (eval mode / no power iteration): v and u are the stored buffers, sigma = u . (W v).
// All reductions run in a fixed order: deterministic; fp32 arithmetic, double for the long sums.
#include <stdint.h>

#include "common.cuh"

namespace ffwm {

constexpr int SN_THREADS = 256;
constexpr int SN_CHUNK = 4096;                   // elements per block of the elementwise kernels
constexpr int SN_FIELDS = 8;                     // table: W, u, v, h, w, element offset, t offset (sum of w), s offset (sum of h)

struct SnLayer {
    const float* W;
    float *u, *v;
    int h, w;
    int64_t off_e, off_t, off_s;
};

// table: [L][8] int64, then three int64 prefix arrays of L + 1 block offsets (kernels 1, 2, 3)
__device__ __forceinline__ SnLayer sn_layer(const int64_t* __restrict__ table, int l) {
    const int64_t* p = table + (int64_t)l * SN_FIELDS;
    SnLayer s;
    s.W = reinterpret_cast<const float*>(p[0]);
    s.u = reinterpret_cast<float*>(p[1]);
    s.v = reinterpret_cast<float*>(p[2]);
    s.h = (int)p[3];
    s.w = (int)p[4];
    s.off_e = p[5];
    s.off_t = p[6];
    s.off_s = p[7];
    return s;
}

// layer of block b in prefix array `which` (0, 1, 2); *first = its first block
__device__ __forceinline__ int sn_find(const int64_t* __restrict__ table, int layers, int which, int b, int* first) {
    const int64_t* pre = table + (int64_t)layers * SN_FIELDS + (int64_t)which * (layers + 1);
    int lo = 0, hi = layers - 1;
    while (lo < hi) {                                                  // last l with pre[l] <= b
        const int mid = (lo + hi + 1) >> 1;
        if (pre[mid] <= b) lo = mid; else hi = mid - 1;
    }
    *first = (int)pre[lo];
    return lo;
}

__device__ __forceinline__ double sn_block_sum(double a, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    a = 0.0;
#pragma unroll
    for (int i = 0; i < SN_THREADS / 32; ++i) a += sh[i];
    return a;
}

// block = (layer, 32 columns): lanes are consecutive columns (128-byte rows of W), the 8 warps interleave the rows and their
// partial sums are added in warp order through shared memory (first version: 256 columns per block, every thread walking all
// h rows — 14 blocks for the 384 x 3456 layer, 91 us per launch)
__global__ void __launch_bounds__(SN_THREADS) sn_wtu_kernel(const int64_t* __restrict__ table, int layers, float* __restrict__ t) {
    __shared__ float part[SN_THREADS / 32][33];
    int first;
    const int l = sn_find(table, layers, 0, blockIdx.x, &first);
    const SnLayer L = sn_layer(table, l);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = (blockIdx.x - first) * 32 + lane;
    float acc = 0.f;
    if (j < L.w) {
        const float* wp = L.W + j;
#pragma unroll 4
        for (int i = warp; i < L.h; i += SN_THREADS / 32) acc = fmaf(__ldg(wp + (int64_t)i * L.w), __ldg(L.u + i), acc);
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && j < L.w) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < SN_THREADS / 32; ++k) a += part[k][lane];
        t[L.off_t + j] = a;
    }
}

__global__ void __launch_bounds__(SN_THREADS) sn_wv_kernel(const int64_t* __restrict__ table, int layers, int update, float eps,
                                                           const float* __restrict__ t, float* __restrict__ s, float* __restrict__ v_saved) {
    __shared__ double sh[SN_THREADS / 32];
    int first;
    const int l = sn_find(table, layers, 1, blockIdx.x, &first);
    const SnLayer L = sn_layer(table, l);
    const int rb = blockIdx.x - first, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* vsrc = update ? t + L.off_t : L.v;
    float inv = 1.f;
    if (update) {
        double q = 0.0;
        for (int j = threadIdx.x; j < L.w; j += SN_THREADS) { const float x = vsrc[j]; q += (double)x * x; }
        q = sn_block_sum(q, sh);
        inv = 1.f / fmaxf((float)sqrt(q), eps);
    }
    if (rb == 0)
        for (int j = threadIdx.x; j < L.w; j += SN_THREADS) {
            const float x = vsrc[j] * inv;
            if (update) L.v[j] = x;
            v_saved[L.off_t + j] = x;
        }
    const int i = rb * (SN_THREADS / 32) + warp;
    if (i < L.h) {
        const float* wp = L.W + (int64_t)i * L.w;
        float acc = 0.f;
        for (int j = lane; j < L.w; j += 32) acc = fmaf(__ldg(wp + j), vsrc[j] * inv, acc);
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s[L.off_s + i] = acc;
    }
}

__global__ void __launch_bounds__(SN_THREADS) sn_scale_kernel(const int64_t* __restrict__ table, int layers, int update, float eps,
                                                              const float* __restrict__ s, float* __restrict__ u_saved, float* __restrict__ sigma,
                                                              float* __restrict__ out) {
    __shared__ double sh[SN_THREADS / 32];
    int first;
    const int l = sn_find(table, layers, 2, blockIdx.x, &first);
    const SnLayer L = sn_layer(table, l);
    const int cb = blockIdx.x - first;
    const float* sp = s + L.off_s;
    double q = 0.0;
    if (update) for (int i = threadIdx.x; i < L.h; i += SN_THREADS) { const float x = sp[i]; q += (double)x * x; }
    else for (int i = threadIdx.x; i < L.h; i += SN_THREADS) q += (double)L.u[i] * sp[i];
    q = sn_block_sum(q, sh);
    float sg, inv_u = 1.f;
    if (update) {
        const float un = fmaxf((float)sqrt(q), eps);
        inv_u = 1.f / un;
        sg = (float)(q / (double)un);                                  // u . (W v) with u = s / max(|s|, eps)
    } else sg = (float)q;
    if (cb == 0) {
        for (int i = threadIdx.x; i < L.h; i += SN_THREADS) {
            const float x = update ? sp[i] * inv_u : L.u[i];
            if (update) L.u[i] = x;
            u_saved[L.off_s + i] = x;
        }
        if (threadIdx.x == 0) sigma[l] = sg;
    }
    const int64_t n = (int64_t)L.h * L.w, e0 = (int64_t)cb * SN_CHUNK;
    const int len = (int)(n - e0 < SN_CHUNK ? n - e0 : SN_CHUNK);
    for (int e = threadIdx.x; e < len; e += SN_THREADS) out[L.off_e + e0 + e] = __ldg(L.W + e0 + e) / sg;
}

__global__ void __launch_bounds__(SN_THREADS) sn_bwd_dot_kernel(const int64_t* __restrict__ table, int layers, const float* __restrict__ g,
                                                                double* __restrict__ partial) {
    __shared__ double sh[SN_THREADS / 32];
    int first;
    const int l = sn_find(table, layers, 2, blockIdx.x, &first);
    const SnLayer L = sn_layer(table, l);
    const int64_t n = (int64_t)L.h * L.w, e0 = (int64_t)(blockIdx.x - first) * SN_CHUNK;
    const int len = (int)(n - e0 < SN_CHUNK ? n - e0 : SN_CHUNK);
    float a = 0.f;
    for (int e = threadIdx.x; e < len; e += SN_THREADS) a = fmaf(g[L.off_e + e0 + e], __ldg(L.W + e0 + e), a);
    const double q = sn_block_sum((double)a, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = q;
}

__global__ void __launch_bounds__(SN_THREADS) sn_bwd_apply_kernel(const int64_t* __restrict__ table, int layers, const float* __restrict__ g,
                                                                  const float* __restrict__ u_saved, const float* __restrict__ v_saved,
                                                                  const float* __restrict__ sigma, const double* __restrict__ partial,
                                                                  float* __restrict__ gw) {
    __shared__ double sh[SN_THREADS / 32];
    int first;
    const int l = sn_find(table, layers, 2, blockIdx.x, &first);
    const SnLayer L = sn_layer(table, l);
    const int64_t* pre = table + (int64_t)layers * SN_FIELDS + 2 * (int64_t)(layers + 1);
    const int nblk = (int)(pre[l + 1] - pre[l]);
    double r = 0.0;
    for (int i = threadIdx.x; i < nblk; i += SN_THREADS) r += partial[first + i];
    r = sn_block_sum(r, sh);
    const float sg = sigma[l], inv = 1.f / sg, coef = (float)(r / ((double)sg * (double)sg));
    const int64_t n = (int64_t)L.h * L.w, e0 = (int64_t)(blockIdx.x - first) * SN_CHUNK;
    const int len = (int)(n - e0 < SN_CHUNK ? n - e0 : SN_CHUNK);
    const float* up = u_saved + L.off_s;
    const float* vp = v_saved + L.off_t;
    for (int e = threadIdx.x; e < len; e += SN_THREADS) {
        const int64_t idx = e0 + e;
        const int i = (int)(idx / L.w), j = (int)(idx - (int64_t)i * L.w);
        gw[L.off_e + idx] = g[L.off_e + idx] * inv - coef * up[i] * vp[j];
    }
}

}  // namespace ffwm

// table (device): `layers` rows of 8 int64 {weight_orig pointer (h x w row-major fp32), u pointer (h), v pointer (w), h, w, offset
// of the layer in the flat element buffers, in the flat t / v buffers (sum of w), in the flat s / u buffers (sum of h)}, then
// three arrays of layers + 1 block prefix sums: ceil(w / 32), ceil(h / 8), ceil(h * w / 4096) blocks per layer; blocks1..3 are
// their totals.  out (sum of h*w floats) = W / sigma per layer; t, v_saved (sum of w), s, u_saved (sum of h), sigma (layers).
// update = 1: one power iteration, u and v are overwritten (training mode); 0: u, v as stored.
extern "C" int ffwm_spectral_norm_forward(const void* table, int layers, int update, float eps, float* out, float* t, float* s,
                                          float* u_saved, float* v_saved, float* sigma, int blocks1, int blocks2, int blocks3, void* stream) {
    using namespace ffwm;
    if (!table || !out || !t || !s || !u_saved || !v_saved || !sigma) { set_error("spectral_norm_forward: null pointer"); return FFWM_ERR_NULL; }
    if (layers < 1 || blocks1 < 1 || blocks2 < 1 || blocks3 < 1) { set_error("spectral_norm_forward: empty table"); return FFWM_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t* tb = static_cast<const int64_t*>(table);
    int rc;
    if (update) {
        sn_wtu_kernel<<<blocks1, SN_THREADS, 0, st>>>(tb, layers, t);
        if ((rc = check_launch("spectral_norm_forward (W^T u)"))) return rc;
    }
    sn_wv_kernel<<<blocks2, SN_THREADS, 0, st>>>(tb, layers, update, eps, t, s, v_saved);
    if ((rc = check_launch("spectral_norm_forward (W v)"))) return rc;
    sn_scale_kernel<<<blocks3, SN_THREADS, 0, st>>>(tb, layers, update, eps, s, u_saved, sigma, out);
    return check_launch("spectral_norm_forward (scale)");
}

// grad_w (flat, like out) = d/dW of sum(grad_out * W / sigma) with sigma = u . (W v), u and v constants (the saved ones);
// partial: blocks3 doubles of scratch.
extern "C" int ffwm_spectral_norm_backward(const void* table, int layers, const float* grad_out, const float* u_saved, const float* v_saved,
                                           const float* sigma, float* grad_w, void* partial, int blocks3, void* stream) {
    using namespace ffwm;
    if (!table || !grad_out || !u_saved || !v_saved || !sigma || !grad_w || !partial) { set_error("spectral_norm_backward: null pointer"); return FFWM_ERR_NULL; }
    if (layers < 1 || blocks3 < 1) { set_error("spectral_norm_backward: empty table"); return FFWM_ERR_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t* tb = static_cast<const int64_t*>(table);
    int rc;
    sn_bwd_dot_kernel<<<blocks3, SN_THREADS, 0, st>>>(tb, layers, grad_out, static_cast<double*>(partial));
    if ((rc = check_launch("spectral_norm_backward (dot)"))) return rc;
    sn_bwd_apply_kernel<<<blocks3, SN_THREADS, 0, st>>>(tb, layers, grad_out, u_saved, v_saved, sigma, static_cast<const double*>(partial), grad_w);
    return check_launch("spectral_norm_backward (apply)");
}
