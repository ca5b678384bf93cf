// batch_norm: training-mode BatchNorm2d (+ the LeakyReLU that follows it, + the residual add of a ResidualBlock) as TWO
// streaming kernels per direction.
//
// Every conv unit of FlowNet, the FFWM generator and the discriminator is conv -> nn.BatchNorm2d -> LeakyReLU(0.2)
// (reference models/base_networks.py:15-45 conv/deconv/i_conv, :179-206 _block/ResidualBlock, :397-410 make_net): 112
// batch-norm layers per train step.  The library kernels PyTorch dispatches to give ONE thread block to a channel
// (`bn_fw_tr_1C11_kernel_NCHW`, `bn_bw_1C11_kernel_new`: grid = C), so a 64-channel 128x128 map is normalised by 64 of
// the 148 SMs and the pass runs far below HBM speed: 7.2 ms of the 57.9 ms step (profiles/r02s_launches_train_summary.txt),
// plus 1.4 ms of separate LeakyReLU forward / backward kernels.  Here the work is split over (plane, chunk) blocks:
//
//   forward   bn_stats_kernel   every block sums (x - pivot) and (x - pivot)^2 over its <= 4096-element chunk (pivot = the
//                               channel's first element: shifted sums keep E[d^2] - E[d]^2 well conditioned), fp32 per
//                               thread over <= 16 values, double from there on; one (sum, sum of squares) pair per block
//                               goes to the workspace.
//             bn_apply_kernel   every block re-reduces its channel's partials (N * chunks pairs, L2 hits, fixed order ->
//                               all blocks of a channel get bit-identical statistics, run-to-run deterministic), then
//                               streams y = lrelu((x - mean) * (gamma * invstd) + beta [+ residual]).  The block (n = 0,
//                               chunk = 0) of each channel also writes save_mean / save_invstd and the running statistics
//                               (momentum update, unbiased variance — torch.nn.functional.batch_norm's training semantics).
//   backward  bn_bwd_stats_kernel   ge = grad_out * lrelu'(y) (y's sign recomputed from x with the forward's exact
//                               expression, or read from the saved output when a residual was added), partial sums of ge
//                               and ge * (x - mean); writes ge as the residual branch's gradient when there is one.
//             bn_bwd_apply_kernel   grad_x = gamma * invstd * (ge - mean(ge) - (x - mean) * invstd^2 * mean(ge * (x - mean)));
//                               the first block of each channel writes grad_gamma / grad_beta.
//
// Traffic: forward 2 reads + 1 write of the map (the second read hits L2 for maps under ~60 MB), backward 4 reads + 1
// write; no LeakyReLU pass in either direction.  Bound: HBM.  fp32 only (the train step's dtype); contiguous NCHW.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace ffwm {

constexpr int BN_THREADS = 256;
constexpr int BN_CHUNK = 4096;      // elements per block: 4 float4 per thread

struct BnGeo {
    int n, c, hw;
    int chunk, chunks;              // elements per chunk (multiple of 4 when vec), chunks per plane
    int vec;                        // hw % 4 == 0 and every base pointer 16-byte aligned
    float eps, momentum, slope;     // slope 1 = no activation
};

// sum of (a, b) over the block, result in every thread; fixed order
__device__ __forceinline__ void bn_block_sum2(double& a, double& b, double* sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        sh[2 * (threadIdx.x >> 5)] = a;
        sh[2 * (threadIdx.x >> 5) + 1] = b;
    }
    __syncthreads();
    a = 0.0;
    b = 0.0;
#pragma unroll
    for (int i = 0; i < BN_THREADS / 32; ++i) {
        a += sh[2 * i];
        b += sh[2 * i + 1];
    }
}

__device__ __forceinline__ void bn_channel_totals(const double* __restrict__ partial, int c, int per_channel, double* sh, double& a, double& b) {
    a = 0.0;
    b = 0.0;
    const double* p = partial + (int64_t)c * per_channel * 2;
    for (int i = threadIdx.x; i < per_channel; i += BN_THREADS) {
        a += p[2 * i];
        b += p[2 * i + 1];
    }
    bn_block_sum2(a, b, sh);
}

// the one expression both directions use for the normalised value (explicit roundings: no contraction differences)
__device__ __forceinline__ float bn_value(float x, float mean, float k, float beta) { return __fmaf_rn(__fsub_rn(x, mean), k, beta); }
__device__ __forceinline__ float bn_act(float v, float slope) { return v > 0.f ? v : __fmul_rn(v, slope); }

struct BnBlock {
    int plane, c, n, chunk, len;
    int64_t base;
};
__device__ __forceinline__ BnBlock bn_block(const BnGeo& g) {
    BnBlock b;
    b.plane = blockIdx.x;
    b.chunk = blockIdx.y;
    b.c = b.plane % g.c;
    b.n = b.plane / g.c;
    b.base = (int64_t)b.plane * g.hw + (int64_t)b.chunk * g.chunk;
    b.len = min(g.chunk, g.hw - b.chunk * g.chunk);
    return b;
}

__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const float* __restrict__ x, BnGeo g, double* __restrict__ partial) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const BnBlock k = bn_block(g);
    const float pivot = __ldg(x + (int64_t)k.c * g.hw);
    float s1 = 0.f, s2 = 0.f;
    if (g.vec) {
        const float4* p = reinterpret_cast<const float4*>(x + k.base);
#pragma unroll 4
        for (int i = threadIdx.x; i < k.len / 4; i += BN_THREADS) {
            const float4 v = __ldg(p + i);
            const float d0 = v.x - pivot, d1 = v.y - pivot, d2 = v.z - pivot, d3 = v.w - pivot;
            s1 += (d0 + d1) + (d2 + d3);
            s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
        }
    } else {
        for (int i = threadIdx.x; i < k.len; i += BN_THREADS) {
            const float d = __ldg(x + k.base + i) - pivot;
            s1 += d;
            s2 = fmaf(d, d, s2);
        }
    }
    double a = s1, b = s2;
    bn_block_sum2(a, b, sh);
    if (threadIdx.x == 0) {
        double* q = partial + ((int64_t)k.c * (g.n * g.chunks) + k.n * g.chunks + k.chunk) * 2;
        q[0] = a;
        q[1] = b;
    }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ residual, BnGeo g, const double* __restrict__ partial,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ running_mean,
                float* __restrict__ running_var, float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ y) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const BnBlock k = bn_block(g);
    double a, b;
    bn_channel_totals(partial, k.c, g.n * g.chunks, sh, a, b);
    const double cnt = (double)g.n * g.hw;
    const double m = a / cnt;
    const double var = fmax(b / cnt - m * m, 0.0);
    const float mean = (float)((double)__ldg(x + (int64_t)k.c * g.hw) + m);
    const float invstd = (float)(1.0 / sqrt(var + (double)g.eps));
    if (k.n == 0 && k.chunk == 0 && threadIdx.x == 0) {
        save_mean[k.c] = mean;
        save_invstd[k.c] = invstd;
        if (running_mean) running_mean[k.c] = (1.f - g.momentum) * running_mean[k.c] + g.momentum * mean;
        if (running_var) running_var[k.c] = (1.f - g.momentum) * running_var[k.c] + g.momentum * (float)(var * (cnt / (cnt - 1.0)));
    }
    const float kk = __fmul_rn(__ldg(gamma + k.c), invstd), bt = __ldg(beta + k.c), slope = g.slope;
    if (g.vec) {
        const float4* p = reinterpret_cast<const float4*>(x + k.base);
        const float4* r = residual ? reinterpret_cast<const float4*>(residual + k.base) : nullptr;
        float4* q = reinterpret_cast<float4*>(y + k.base);
#pragma unroll 4
        for (int i = threadIdx.x; i < k.len / 4; i += BN_THREADS) {
            const float4 v = __ldg(p + i);
            float4 o = make_float4(bn_value(v.x, mean, kk, bt), bn_value(v.y, mean, kk, bt), bn_value(v.z, mean, kk, bt), bn_value(v.w, mean, kk, bt));
            if (r) {
                const float4 rr = __ldg(r + i);
                o.x = __fadd_rn(o.x, rr.x);
                o.y = __fadd_rn(o.y, rr.y);
                o.z = __fadd_rn(o.z, rr.z);
                o.w = __fadd_rn(o.w, rr.w);
            }
            if (slope != 1.f) o = make_float4(bn_act(o.x, slope), bn_act(o.y, slope), bn_act(o.z, slope), bn_act(o.w, slope));
            q[i] = o;
        }
    } else {
        for (int i = threadIdx.x; i < k.len; i += BN_THREADS) {
            float o = bn_value(__ldg(x + k.base + i), mean, kk, bt);
            if (residual) o = __fadd_rn(o, __ldg(residual + k.base + i));
            y[k.base + i] = slope != 1.f ? bn_act(o, slope) : o;
        }
    }
}

// ge = grad_out * lrelu'(y): the sign of y from the saved output (residual blocks) or recomputed from x
__device__ __forceinline__ float bn_ge(float dy, float x, float yout, bool have_y, float mean, float kk, float bt, float slope) {
    if (slope == 1.f) return dy;
    const float v = have_y ? yout : bn_value(x, mean, kk, bt);
    return v > 0.f ? dy : __fmul_rn(dy, slope);
}

__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ yout, BnGeo g,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
                    const float* __restrict__ save_invstd, float* __restrict__ grad_residual, double* __restrict__ partial) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const BnBlock k = bn_block(g);
    const float mean = __ldg(save_mean + k.c), kk = __fmul_rn(__ldg(gamma + k.c), __ldg(save_invstd + k.c)), bt = __ldg(beta + k.c);
    const float slope = g.slope;
    const bool have_y = yout != nullptr;
    float s1 = 0.f, s2 = 0.f;
    if (g.vec) {
        const float4* px = reinterpret_cast<const float4*>(x + k.base);
        const float4* pd = reinterpret_cast<const float4*>(dy + k.base);
        const float4* py = have_y ? reinterpret_cast<const float4*>(yout + k.base) : nullptr;
        float4* pr = grad_residual ? reinterpret_cast<float4*>(grad_residual + k.base) : nullptr;
#pragma unroll 4
        for (int i = threadIdx.x; i < k.len / 4; i += BN_THREADS) {
            const float4 v = __ldg(px + i), d = __ldg(pd + i);
            const float4 yo = have_y ? __ldg(py + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 e = make_float4(bn_ge(d.x, v.x, yo.x, have_y, mean, kk, bt, slope), bn_ge(d.y, v.y, yo.y, have_y, mean, kk, bt, slope),
                                         bn_ge(d.z, v.z, yo.z, have_y, mean, kk, bt, slope), bn_ge(d.w, v.w, yo.w, have_y, mean, kk, bt, slope));
            if (pr) pr[i] = e;
            s1 += (e.x + e.y) + (e.z + e.w);
            s2 = fmaf(e.x, v.x - mean, fmaf(e.y, v.y - mean, fmaf(e.z, v.z - mean, fmaf(e.w, v.w - mean, s2))));
        }
    } else {
        for (int i = threadIdx.x; i < k.len; i += BN_THREADS) {
            const float v = __ldg(x + k.base + i);
            const float e = bn_ge(__ldg(dy + k.base + i), v, have_y ? __ldg(yout + k.base + i) : 0.f, have_y, mean, kk, bt, slope);
            if (grad_residual) grad_residual[k.base + i] = e;
            s1 += e;
            s2 = fmaf(e, v - mean, s2);
        }
    }
    double a = s1, b = s2;
    bn_block_sum2(a, b, sh);
    if (threadIdx.x == 0) {
        double* q = partial + ((int64_t)k.c * (g.n * g.chunks) + k.n * g.chunks + k.chunk) * 2;
        q[0] = a;
        q[1] = b;
    }
}

// ge: grad_residual when the stats pass wrote it (residual blocks), else recomputed from (x, dy)
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ ge_in, BnGeo g,
                    const double* __restrict__ partial, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float* __restrict__ grad_x,
                    float* __restrict__ grad_gamma, float* __restrict__ grad_beta) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const BnBlock k = bn_block(g);
    double a, b;
    bn_channel_totals(partial, k.c, g.n * g.chunks, sh, a, b);
    const float mean = __ldg(save_mean + k.c), invstd = __ldg(save_invstd + k.c);
    const float kk = __fmul_rn(__ldg(gamma + k.c), invstd), bt = __ldg(beta + k.c), slope = g.slope;
    const double cnt = (double)g.n * g.hw;
    if (k.n == 0 && k.chunk == 0 && threadIdx.x == 0) {
        if (grad_gamma) grad_gamma[k.c] = (float)(b * (double)invstd);
        if (grad_beta) grad_beta[k.c] = (float)a;
    }
    const float ma = (float)(a / cnt), mb = (float)(b * (double)invstd * (double)invstd / cnt);
    const bool have_ge = ge_in != nullptr;
    const float* src = have_ge ? ge_in : dy;
    if (g.vec) {
        const float4* px = reinterpret_cast<const float4*>(x + k.base);
        const float4* pd = reinterpret_cast<const float4*>(src + k.base);
        float4* q = reinterpret_cast<float4*>(grad_x + k.base);
#pragma unroll 4
        for (int i = threadIdx.x; i < k.len / 4; i += BN_THREADS) {
            const float4 v = __ldg(px + i), d = __ldg(pd + i);
            float4 e = d;
            if (!have_ge)
                e = make_float4(bn_ge(d.x, v.x, 0.f, false, mean, kk, bt, slope), bn_ge(d.y, v.y, 0.f, false, mean, kk, bt, slope),
                                bn_ge(d.z, v.z, 0.f, false, mean, kk, bt, slope), bn_ge(d.w, v.w, 0.f, false, mean, kk, bt, slope));
            q[i] = make_float4(kk * (e.x - ma - (v.x - mean) * mb), kk * (e.y - ma - (v.y - mean) * mb), kk * (e.z - ma - (v.z - mean) * mb),
                               kk * (e.w - ma - (v.w - mean) * mb));
        }
    } else {
        for (int i = threadIdx.x; i < k.len; i += BN_THREADS) {
            const float v = __ldg(x + k.base + i), d = __ldg(src + k.base + i);
            const float e = have_ge ? d : bn_ge(d, v, 0.f, false, mean, kk, bt, slope);
            grad_x[k.base + i] = kk * (e - ma - (v - mean) * mb);
        }
    }
}

// ---- small maps: one block per channel, the channel in registers, one launch per direction ---------------------------------
// Up to 8 192 values per channel (N * H*W; most of FlowNet below 32x32, the generator's 16x16 / 32x32 stages): thread t holds
// the float4 chunks t, t + 256, ... of its channel (R <= 8 of them), so the statistics are an exact two-pass mean / variance
// and the map crosses HBM once per direction; the two-kernel path above would spend its time in launch latency here.
constexpr int BN_SMALL_MAX = 8192;

template <int R>
__device__ __forceinline__ void bn_small_load(const float* __restrict__ p, const BnGeo& g, int c, float4 (&v)[R]) {
    const int hw4 = g.hw >> 2, total4 = g.n * hw4;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int e = threadIdx.x + k * BN_THREADS;
        if (e < total4) {
            const int n = e / hw4, o = e - n * hw4;
            v[k] = __ldg(reinterpret_cast<const float4*>(p + ((int64_t)n * g.c + c) * g.hw) + o);
        } else v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int R>
__device__ __forceinline__ void bn_small_store(float* __restrict__ p, const BnGeo& g, int c, const float4 (&v)[R]) {
    const int hw4 = g.hw >> 2, total4 = g.n * hw4;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int e = threadIdx.x + k * BN_THREADS;
        if (e < total4) {
            const int n = e / hw4, o = e - n * hw4;
            reinterpret_cast<float4*>(p + ((int64_t)n * g.c + c) * g.hw)[o] = v[k];
        }
    }
}

template <int R>
__global__ void __launch_bounds__(BN_THREADS)
bn_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ residual, BnGeo g, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                    float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ y) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const int c = blockIdx.x, total4 = g.n * (g.hw >> 2);
    float4 v[R];
    bn_small_load<R>(x, g, c, v);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    double a = s, b = 0.0;
    bn_block_sum2(a, b, sh);
    const double cnt = (double)g.n * g.hw;
    const float mean = (float)(a / cnt);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k)
        if (threadIdx.x + k * BN_THREADS < total4) {
            const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;
            q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q))));
        }
    a = q, b = 0.0;
    bn_block_sum2(a, b, sh);
    const double var = a / cnt;
    const float invstd = (float)(1.0 / sqrt(var + (double)g.eps));
    if (threadIdx.x == 0) {
        save_mean[c] = mean;
        save_invstd[c] = invstd;
        if (running_mean) running_mean[c] = (1.f - g.momentum) * running_mean[c] + g.momentum * mean;
        if (running_var) running_var[c] = (1.f - g.momentum) * running_var[c] + g.momentum * (float)(var * (cnt / (cnt - 1.0)));
    }
    const float kk = __fmul_rn(__ldg(gamma + c), invstd), bt = __ldg(beta + c), slope = g.slope;
    float4 r[R];
    if (residual) bn_small_load<R>(residual, g, c, r);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        float4 o = make_float4(bn_value(v[k].x, mean, kk, bt), bn_value(v[k].y, mean, kk, bt), bn_value(v[k].z, mean, kk, bt), bn_value(v[k].w, mean, kk, bt));
        if (residual) o = make_float4(__fadd_rn(o.x, r[k].x), __fadd_rn(o.y, r[k].y), __fadd_rn(o.z, r[k].z), __fadd_rn(o.w, r[k].w));
        if (slope != 1.f) o = make_float4(bn_act(o.x, slope), bn_act(o.y, slope), bn_act(o.z, slope), bn_act(o.w, slope));
        v[k] = o;
    }
    bn_small_store<R>(y, g, c, v);
}

template <int R>
__global__ void __launch_bounds__(BN_THREADS)
bn_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ yout, BnGeo g,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ save_mean,
                    const float* __restrict__ save_invstd, float* __restrict__ grad_x, float* __restrict__ grad_residual,
                    float* __restrict__ grad_gamma, float* __restrict__ grad_beta) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const int c = blockIdx.x;
    const float mean = __ldg(save_mean + c), invstd = __ldg(save_invstd + c);
    const float kk = __fmul_rn(__ldg(gamma + c), invstd), bt = __ldg(beta + c), slope = g.slope;
    const bool have_y = yout != nullptr;
    float4 v[R], e[R];
    bn_small_load<R>(x, g, c, v);
    bn_small_load<R>(dy, g, c, e);                                     // padding chunks: dy = 0 -> ge = 0
    if (slope != 1.f) {
        float4 yo[R];
        if (have_y) bn_small_load<R>(yout, g, c, yo);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const float4 w = have_y ? yo[k] : make_float4(0.f, 0.f, 0.f, 0.f);
            e[k] = make_float4(bn_ge(e[k].x, v[k].x, w.x, have_y, mean, kk, bt, slope), bn_ge(e[k].y, v[k].y, w.y, have_y, mean, kk, bt, slope),
                               bn_ge(e[k].z, v[k].z, w.z, have_y, mean, kk, bt, slope), bn_ge(e[k].w, v[k].w, w.w, have_y, mean, kk, bt, slope));
        }
    }
    if (grad_residual) bn_small_store<R>(grad_residual, g, c, e);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        s1 += (e[k].x + e[k].y) + (e[k].z + e[k].w);
        s2 = fmaf(e[k].x, v[k].x - mean, fmaf(e[k].y, v[k].y - mean, fmaf(e[k].z, v[k].z - mean, fmaf(e[k].w, v[k].w - mean, s2))));
    }
    double a = s1, b = s2;
    bn_block_sum2(a, b, sh);
    const double cnt = (double)g.n * g.hw;
    if (threadIdx.x == 0) {
        if (grad_gamma) grad_gamma[c] = (float)(b * (double)invstd);
        if (grad_beta) grad_beta[c] = (float)a;
    }
    const float ma = (float)(a / cnt), mb = (float)(b * (double)invstd * (double)invstd / cnt);
#pragma unroll
    for (int k = 0; k < R; ++k)
        v[k] = make_float4(kk * (e[k].x - ma - (v[k].x - mean) * mb), kk * (e[k].y - ma - (v[k].y - mean) * mb),
                           kk * (e[k].z - ma - (v[k].z - mean) * mb), kk * (e[k].w - ma - (v[k].w - mean) * mb));
    bn_small_store<R>(grad_x, g, c, v);
}

static bool bn_small(const BnGeo& g) { return g.vec && (int64_t)g.n * g.hw <= BN_SMALL_MAX && !opt(OPT_BN_NO_SMALL); }
static int bn_small_r(const BnGeo& g) {
    const int need = (g.n * (g.hw >> 2) + BN_THREADS - 1) / BN_THREADS;
    return need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : 8;
}

// ---- per-channel sum of an (N, C, H*W) map: the bias gradient of a convolution (grad_out.sum((0, 2, 3))) ------------------
// at::sum gives this reduction C x 4 blocks (64 us for the 102 MB gradients of dres2, 1.6 TB/s); here every (plane, chunk)
// block writes one double, and one warp per channel adds the N * chunks partials in a fixed order.
__global__ void __launch_bounds__(BN_THREADS) channel_sum_partial_kernel(const float* __restrict__ x, BnGeo g, double* __restrict__ partial) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const BnBlock k = bn_block(g);
    float s1 = 0.f, s2 = 0.f;                                          // two chains: shorter dependency, better rounding
    if (g.vec) {
        const float4* p = reinterpret_cast<const float4*>(x + k.base);
#pragma unroll 4
        for (int i = threadIdx.x; i < k.len / 4; i += BN_THREADS) {
            const float4 v = __ldg(p + i);
            s1 += v.x + v.y;
            s2 += v.z + v.w;
        }
    } else {
        for (int i = threadIdx.x; i < k.len; i += BN_THREADS) s1 += __ldg(x + k.base + i);
    }
    double a = (double)s1 + (double)s2, b = 0.0;
    bn_block_sum2(a, b, sh);
    if (threadIdx.x == 0) partial[(int64_t)k.c * (g.n * g.chunks) + k.n * g.chunks + k.chunk] = a;
}

// small maps (<= 8192 values per channel): one block per channel, one launch
__global__ void __launch_bounds__(BN_THREADS) channel_sum_small_kernel(const float* __restrict__ x, BnGeo g, float* __restrict__ out) {
    __shared__ double sh[2 * BN_THREADS / 32];
    const int c = blockIdx.x, hw4 = g.hw >> 2, total4 = g.n * hw4;
    float s1 = 0.f, s2 = 0.f;
    for (int e = threadIdx.x; e < total4; e += BN_THREADS) {
        const int n = e / hw4, o = e - n * hw4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + ((int64_t)n * g.c + c) * g.hw) + o);
        s1 += v.x + v.y;
        s2 += v.z + v.w;
    }
    double a = (double)s1 + (double)s2, b = 0.0;
    bn_block_sum2(a, b, sh);
    if (threadIdx.x == 0) out[c] = (float)a;
}

__global__ void __launch_bounds__(BN_THREADS) channel_sum_final_kernel(const double* __restrict__ partial, int c, int per_channel, float* __restrict__ out) {
    const int ch = blockIdx.x * (BN_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ch >= c) return;
    double a = 0.0;
    for (int i = lane; i < per_channel; i += 32) a += partial[(int64_t)ch * per_channel + i];
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[ch] = (float)a;
}

static int bn_geo(const char* what, int n, int c, int64_t hw, float eps, float momentum, float slope, std::initializer_list<const void*> ptrs,
                  BnGeo* g) {
    if (n < 1 || c < 1 || hw < 1) { set_error("%s: empty map (%d x %d x %lld)", what, n, c, (long long)hw); return FFWM_ERR_SHAPE; }
    if (hw >= (1ll << 31) || (int64_t)n * c >= (1ll << 31)) { set_error("%s: plane or plane count beyond 2^31", what); return FFWM_ERR_TOO_LARGE; }
    if ((int64_t)n * hw < 2) { set_error("%s: more than one value per channel is needed in training mode", what); return FFWM_ERR_SHAPE; }
    g->n = n;
    g->c = c;
    g->hw = (int)hw;
    g->vec = hw % 4 == 0;
    for (const void* p : ptrs)
        if (p && (reinterpret_cast<uintptr_t>(p) & 15)) g->vec = 0;
    g->chunk = BN_CHUNK;
    g->chunks = (int)((hw + g->chunk - 1) / g->chunk);
    if (g->chunks > 65535) {                      // gridDim.y limit: longer chunks
        g->chunk = (int)(((hw + 65534) / 65535 + 3) / 4 * 4);
        g->chunks = (int)((hw + g->chunk - 1) / g->chunk);
    }
    g->eps = eps;
    g->momentum = momentum;
    g->slope = slope;
    return FFWM_OK;
}

static int64_t bn_workspace(const BnGeo& g) { return (int64_t)g.c * g.n * g.chunks * 2 * (int64_t)sizeof(double); }

}  // namespace ffwm

extern "C" int64_t ffwm_batch_norm_workspace_bytes(int n, int c, int64_t hw) {
    ffwm::BnGeo g;
    if (ffwm::bn_geo("batch_norm_workspace_bytes", n, c, hw, 0.f, 0.f, 1.f, {}, &g)) return -1;
    return ffwm::bn_workspace(g);
}

extern "C" int ffwm_batch_norm_forward(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                                       float* running_var, float momentum, float eps, float act_slope, float* y, float* save_mean,
                                       float* save_invstd, int n, int c, int64_t hw, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace ffwm;
    if (!x || !gamma || !beta || !y || !save_mean || !save_invstd || !workspace) { set_error("batch_norm_forward: null pointer"); return FFWM_ERR_NULL; }
    BnGeo g;
    int rc;
    if ((rc = bn_geo("batch_norm_forward", n, c, hw, eps, momentum, act_slope, {x, residual, y}, &g))) return rc;
    if (workspace_bytes < bn_workspace(g) || (reinterpret_cast<uintptr_t>(workspace) & 7)) {
        set_error("batch_norm_forward: workspace of %lld bytes (8-byte aligned) needed, got %lld", (long long)bn_workspace(g), (long long)workspace_bytes);
        return FFWM_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (bn_small(g)) {
#define FFWM_BN_SF(R) bn_small_fwd_kernel<R><<<c, BN_THREADS, 0, st>>>(x, residual, g, gamma, beta, running_mean, running_var, save_mean, save_invstd, y)
        switch (bn_small_r(g)) { case 1: FFWM_BN_SF(1); break; case 2: FFWM_BN_SF(2); break; case 4: FFWM_BN_SF(4); break; default: FFWM_BN_SF(8); }
#undef FFWM_BN_SF
        return check_launch("batch_norm_forward (small)");
    }
    const dim3 grid(n * c, g.chunks);
    double* partial = static_cast<double*>(workspace);
    bn_stats_kernel<<<grid, BN_THREADS, 0, st>>>(x, g, partial);
    if ((rc = check_launch("batch_norm_forward (statistics)"))) return rc;
    bn_apply_kernel<<<grid, BN_THREADS, 0, st>>>(x, residual, g, partial, gamma, beta, running_mean, running_var, save_mean, save_invstd, y);
    return check_launch("batch_norm_forward (apply)");
}

extern "C" int ffwm_batch_norm_backward(const float* x, const float* grad_out, const float* y_out, const float* gamma, const float* beta,
                                        const float* save_mean, const float* save_invstd, float act_slope, float* grad_x,
                                        float* grad_residual, float* grad_gamma, float* grad_beta, int n, int c, int64_t hw, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
    using namespace ffwm;
    if (!x || !grad_out || !gamma || !beta || !save_mean || !save_invstd || !grad_x || !workspace) {
        set_error("batch_norm_backward: null pointer");
        return FFWM_ERR_NULL;
    }
    if (grad_residual && act_slope != 1.f && !y_out) {
        set_error("batch_norm_backward: the saved output is needed for the activation's sign when a residual was added");
        return FFWM_ERR_NULL;
    }
    BnGeo g;
    int rc;
    if ((rc = bn_geo("batch_norm_backward", n, c, hw, 0.f, 0.f, act_slope, {x, grad_out, y_out, grad_x, grad_residual}, &g))) return rc;
    if (workspace_bytes < bn_workspace(g) || (reinterpret_cast<uintptr_t>(workspace) & 7)) {
        set_error("batch_norm_backward: workspace of %lld bytes (8-byte aligned) needed, got %lld", (long long)bn_workspace(g), (long long)workspace_bytes);
        return FFWM_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (bn_small(g)) {
#define FFWM_BN_SB(R) bn_small_bwd_kernel<R><<<c, BN_THREADS, 0, st>>>(x, grad_out, y_out, g, gamma, beta, save_mean, save_invstd, grad_x, grad_residual, grad_gamma, grad_beta)
        switch (bn_small_r(g)) { case 1: FFWM_BN_SB(1); break; case 2: FFWM_BN_SB(2); break; case 4: FFWM_BN_SB(4); break; default: FFWM_BN_SB(8); }
#undef FFWM_BN_SB
        return check_launch("batch_norm_backward (small)");
    }
    const dim3 grid(n * c, g.chunks);
    double* partial = static_cast<double*>(workspace);
    bn_bwd_stats_kernel<<<grid, BN_THREADS, 0, st>>>(x, grad_out, y_out, g, gamma, beta, save_mean, save_invstd, grad_residual, partial);
    if ((rc = check_launch("batch_norm_backward (sums)"))) return rc;
    bn_bwd_apply_kernel<<<grid, BN_THREADS, 0, st>>>(x, grad_out, grad_residual, g, partial, gamma, beta, save_mean, save_invstd, grad_x, grad_gamma,
                                                      grad_beta);
    return check_launch("batch_norm_backward (apply)");
}

// out[c] = sum over (n, hw) of x[n][c][hw]; x contiguous (N, C, H*W) fp32; workspace: ffwm_batch_norm_workspace_bytes(n, c, hw).
extern "C" int ffwm_channel_sum(const float* x, float* out, int n, int c, int64_t hw, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace ffwm;
    if (!x || !out || !workspace) { set_error("channel_sum: null pointer"); return FFWM_ERR_NULL; }
    if (n < 1 || c < 1 || hw < 1) { set_error("channel_sum: empty map"); return FFWM_ERR_SHAPE; }
    BnGeo g;
    int rc;
    if ((int64_t)n * hw < 2) {                                        // bn_geo insists on two values per channel
        if ((rc = bn_geo("channel_sum", 2, c, hw, 0.f, 0.f, 1.f, {x}, &g))) return rc;
        g.n = 1;
    } else if ((rc = bn_geo("channel_sum", n, c, hw, 0.f, 0.f, 1.f, {x}, &g))) return rc;
    if (workspace_bytes < (int64_t)g.c * g.n * g.chunks * (int64_t)sizeof(double) || (reinterpret_cast<uintptr_t>(workspace) & 7)) {
        set_error("channel_sum: workspace too small or misaligned");
        return FFWM_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g.vec && (int64_t)g.n * g.hw <= BN_SMALL_MAX && !opt(OPT_BN_NO_SMALL)) {
        channel_sum_small_kernel<<<c, BN_THREADS, 0, st>>>(x, g, out);
        return check_launch("channel_sum (small)");
    }
    double* partial = static_cast<double*>(workspace);
    channel_sum_partial_kernel<<<dim3(g.n * c, g.chunks), BN_THREADS, 0, st>>>(x, g, partial);
    if ((rc = check_launch("channel_sum (partials)"))) return rc;
    channel_sum_final_kernel<<<(c + BN_THREADS / 32 - 1) / (BN_THREADS / 32), BN_THREADS, 0, st>>>(partial, c, g.n * g.chunks, out);
    return check_launch("channel_sum");
}
