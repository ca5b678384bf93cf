// local_attn_reshape: (B,k*k,H,W) -> (B,1,k*H,k*W) gather and its inverse.
//
// Semantics restate cuda/local_attn_reshape/local_attn_reshape_kernel.cu
// (K6 :20-61, K7 :65-108): out[b,0,y,x] = in[b,(y%k)*k + x%k, y/k, x/k].
// Pure data movement, so results are bit-exact.  The backward is the inverse
// gather: no atomics and no zero-filled destination (the reference scatters
// with atomicAdd into a memset buffer although every address is hit once).
//
// Execution plan (HBM-bound: 8 bytes of traffic per element, nothing else):
// the op is a k-way interleave along x.  A warp owns 32 consecutive input
// columns of one input row; for each of the k output rows it interleaves the
// k planes through a private shared-memory row (stride-k accesses, conflict
// free for odd k), so that BOTH the k*k plane reads and the k output-row
// writes are fully coalesced 128-byte transactions.  k is a template
// parameter (2..8; the reference uses 3, 5, 7): no runtime division anywhere.
// Other k fall back to the simple per-pixel kernels at the bottom.
#include "common.cuh"

namespace ffwm {

constexpr int LAR_WARPS = 8;

template <typename T, int K, bool FWD>
__global__ void __launch_bounds__(32 * LAR_WARPS)
lar_tiled_kernel(View<const T> src, View<T> dst) {
    // FWD: src = inputs (B,K*K,H,W), dst = output (B,Co,K*H,K*W)
    // BWD: src = grad_output (B,Cg,K*H,K*W), dst = grad_inputs (B,K*K,H,W)
    __shared__ T stage[LAR_WARPS][K * 32];
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int h = FWD ? src.h : dst.h, w = FWD ? src.w : dst.w;
    const int xs0 = blockIdx.x * 32;
    const int ys = blockIdx.y * LAR_WARPS + warp;
    const int b = blockIdx.z;
    if (ys >= h) return;                       // whole warp leaves together
    const int xs = xs0 + lane;
    const int nx = min(32, w - xs0);           // valid input columns of this warp
    T* row = stage[warp];
    if (FWD) {
        const T* ip = src.p + b * src.sb + ys * src.sh + xs * src.sw;
        T v[K * K];                            // all k*k plane loads in flight before any use
        if (lane < nx) {
#pragma unroll
            for (int q = 0; q < K * K; ++q) v[q] = ld_stream(ip + (int64_t)q * src.sc);
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            if (lane < nx) {
#pragma unroll
                for (int j = 0; j < K; ++j) row[lane * K + j] = v[i * K + j];
            }
            __syncwarp();
            // the reference ignores the output channel index: every channel gets the map
            for (int c = 0; c < dst.c; ++c) {
                T* op = dst.p + b * dst.sb + c * dst.sc + (ys * K + i) * dst.sh + (xs0 * K) * dst.sw;
#pragma unroll
                for (int m = 0; m < K; ++m) {
                    const int idx = m * 32 + lane;
                    if (idx < nx * K) st_stream(op + idx * dst.sw, row[idx]);
                }
            }
            __syncwarp();
        }
    } else {
        T* op = dst.p + b * dst.sb + ys * dst.sh + xs * dst.sw;
        T g[K * K];                            // k rows x k chunks of 32 columns, all loads issued first
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const int goff = (ys * K + i) * src.sh + (xs0 * K) * src.sw;
#pragma unroll
            for (int m = 0; m < K; ++m) {
                const int idx = m * 32 + lane;
                T acc = T(0);
                if (idx < nx * K) {
                    acc = ld_stream(src.p + b * src.sb + goff + idx * src.sw);
                    for (int c = 1; c < src.c; ++c) acc += ld_stream(src.p + b * src.sb + c * src.sc + goff + idx * src.sw);
                }
                g[i * K + m] = acc;
            }
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int m = 0; m < K; ++m) row[m * 32 + lane] = g[i * K + m];
            __syncwarp();
            if (lane < nx) {
#pragma unroll
                for (int j = 0; j < K; ++j) st_stream(op + (int64_t)(i * K + j) * dst.sc, row[lane * K + j]);
            }
            __syncwarp();
        }
    }
}

// ---- any k: one thread per output (fwd) / input (bwd) pixel -------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
local_attn_reshape_fwd_kernel(View<const T> in, View<T> out, int k) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= out.h * out.w) return;
    const int b = blockIdx.z;
    const int y = pix / out.w, x = pix - y * out.w;
    const int ys = y / k, xs = x / k;
    const int cs = (y - ys * k) * k + (x - xs * k);
    const T v = ld_stream(in.plane(b, cs) + ys * in.sh + xs * in.sw);
    T* d = out.p + b * out.sb + y * out.sh + x * out.sw;
    for (int c = 0; c < out.c; ++c, d += out.sc) st_stream(d, v);
}

template <typename T>
__global__ void __launch_bounds__(256)
local_attn_reshape_bwd_kernel(View<const T> gout, View<T> gin, int k) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= gin.h * gin.w) return;
    const int b = blockIdx.z;
    const int ys = pix / gin.w, xs = pix - ys * gin.w;
    T* d = gin.p + b * gin.sb + ys * gin.sh + xs * gin.sw;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            const int off = (ys * k + i) * gout.sh + (xs * k + j) * gout.sw;
            T acc = T(0);
            for (int c = 0; c < gout.c; ++c) acc += __ldg(gout.plane(b, c) + off);
            st_stream(d + (int64_t)(i * k + j) * gin.sc, acc);
        }
}

template <typename T, bool FWD>
static bool launch_tiled(const View<const T>& src, const View<T>& dst, int k, int h, int w, int n, cudaStream_t st) {
    dim3 grid(ceil_div(w, 32), ceil_div(h, LAR_WARPS), n), block(32, LAR_WARPS);
    if (grid.y > 65535) return false;
    switch (k) {
        case 2: lar_tiled_kernel<T, 2, FWD><<<grid, block, 0, st>>>(src, dst); return true;
        case 3: lar_tiled_kernel<T, 3, FWD><<<grid, block, 0, st>>>(src, dst); return true;
        case 4: lar_tiled_kernel<T, 4, FWD><<<grid, block, 0, st>>>(src, dst); return true;
        case 5: lar_tiled_kernel<T, 5, FWD><<<grid, block, 0, st>>>(src, dst); return true;
        case 6: lar_tiled_kernel<T, 6, FWD><<<grid, block, 0, st>>>(src, dst); return true;
        case 7: lar_tiled_kernel<T, 7, FWD><<<grid, block, 0, st>>>(src, dst); return true;
        case 8: lar_tiled_kernel<T, 8, FWD><<<grid, block, 0, st>>>(src, dst); return true;
        default: return false;
    }
}

template <typename T>
static int lar_forward_t(const ffwm_tensor4* a, const ffwm_tensor4* o, int k, cudaStream_t st) {
    View<const T> in;
    View<T> out;
    int rc;
    if ((rc = make_view<const T>(a, "inputs", &in))) return rc;
    if ((rc = make_view<T>(o, "output", &out))) return rc;
    if (k < 1) { set_error("local_attn_reshape: kernel_size=%d", k); return FFWM_ERR_ARG; }
    if (in.c != k * k) { set_error("local_attn_reshape: inputs has %d channels, need k*k=%d", in.c, k * k); return FFWM_ERR_SHAPE; }
    if (out.n != in.n || (int64_t)out.h != (int64_t)k * in.h || (int64_t)out.w != (int64_t)k * in.w || out.c < 1) {
        set_error("local_attn_reshape: output (%d,%d,%d,%d) != (B,1,k*H,k*W)", out.n, out.c, out.h, out.w);
        return FFWM_ERR_SHAPE;
    }
    if ((int64_t)out.n * out.h * out.w == 0) return FFWM_OK;
    if (out.n > 65535) { set_error("local_attn_reshape: batch %d > 65535", out.n); return FFWM_ERR_TOO_LARGE; }
    if (!launch_tiled<T, true>(in, out, k, in.h, in.w, in.n, st)) {
        dim3 grid(ceil_div((int64_t)out.h * out.w, 256), 1, out.n);
        local_attn_reshape_fwd_kernel<T><<<grid, 256, 0, st>>>(in, out, k);
    }
    return check_launch("local_attn_reshape_forward");
}

template <typename T>
static int lar_backward_t(const ffwm_tensor4* go, const ffwm_tensor4* gi, int k, cudaStream_t st) {
    View<const T> gout;
    View<T> gin;
    int rc;
    if ((rc = make_view<const T>(go, "grad_output", &gout))) return rc;
    if ((rc = make_view<T>(gi, "grad_inputs", &gin))) return rc;
    if (k < 1) { set_error("local_attn_reshape: kernel_size=%d", k); return FFWM_ERR_ARG; }
    if (gin.c != k * k) { set_error("local_attn_reshape: grad_inputs has %d channels, need k*k=%d", gin.c, k * k); return FFWM_ERR_SHAPE; }
    if (gout.n != gin.n || (int64_t)gout.h != (int64_t)k * gin.h || (int64_t)gout.w != (int64_t)k * gin.w) {
        set_error("local_attn_reshape_backward: grad_output (%d,%d,%d,%d) != (B,*,k*H,k*W)", gout.n, gout.c, gout.h, gout.w);
        return FFWM_ERR_SHAPE;
    }
    if ((int64_t)gin.n * gin.h * gin.w == 0) return FFWM_OK;
    if (gin.n > 65535) { set_error("local_attn_reshape: batch %d > 65535", gin.n); return FFWM_ERR_TOO_LARGE; }
    if (gout.c < 1 || !launch_tiled<T, false>(gout, gin, k, gin.h, gin.w, gin.n, st)) {
        dim3 grid(ceil_div((int64_t)gin.h * gin.w, 256), 1, gin.n);
        local_attn_reshape_bwd_kernel<T><<<grid, 256, 0, st>>>(gout, gin, k);
    }
    return check_launch("local_attn_reshape_backward");
}

}  // namespace ffwm

extern "C" int ffwm_local_attn_reshape_forward(const ffwm_tensor4* inputs, const ffwm_tensor4* output,
                                               int kernel_size, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::lar_forward_t<float>(inputs, output, kernel_size, st);
    if (dtype == FFWM_F64) return ffwm::lar_forward_t<double>(inputs, output, kernel_size, st);
    ffwm::set_error("local_attn_reshape_forward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}

extern "C" int ffwm_local_attn_reshape_backward(const ffwm_tensor4* grad_output, const ffwm_tensor4* grad_inputs,
                                                int kernel_size, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::lar_backward_t<float>(grad_output, grad_inputs, kernel_size, st);
    if (dtype == FFWM_F64) return ffwm::lar_backward_t<double>(grad_output, grad_inputs, kernel_size, st);
    ffwm::set_error("local_attn_reshape_backward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}
