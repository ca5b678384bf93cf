// local_attn_reshape: (B,k*k,H,W) -> (B,1,k*H,k*W) gather and its inverse.
//
// Semantics restate cuda/local_attn_reshape/local_attn_reshape_kernel.cu
// (K6 :20-61, K7 :65-108): out[b,0,y,x] = in[b,(y%k)*k + x%k, y/k, x/k].
// Pure data movement, so results are bit-exact.  Both directions are written
// as gathers with coalesced, streamed stores; the backward therefore needs no
// atomics and no zero-filled destination (the reference scatters with
// atomicAdd into a memset buffer although every address is hit once).
#include "common.cuh"

namespace ffwm {

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
    // the reference ignores the output channel index: every channel gets the map
    T* d = out.p + b * out.sb + y * out.sh + x * out.sw;
    for (int c = 0; c < out.c; ++c, d += out.sc) st_stream(d, v);
}

// One thread per (b, ys, xs) walks the k*k planes: stores are coalesced per
// plane, the strided reads of neighbouring planes share L1 lines.
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
    dim3 grid(ceil_div((int64_t)out.h * out.w, 256), 1, out.n);
    local_attn_reshape_fwd_kernel<T><<<grid, 256, 0, st>>>(in, out, k);
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
    dim3 grid(ceil_div((int64_t)gin.h * gin.w, 256), 1, gin.n);
    local_attn_reshape_bwd_kernel<T><<<grid, 256, 0, st>>>(gout, gin, k);
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
