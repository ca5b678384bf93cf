// grid_warp: the bilinear feature warp that FFWM's generator really uses.
//
// WarpNet.forward (models/base_networks.py:168-173) and
// PerceptualCorrectness.bilinear_warp (models/losses.py:392-396) call
//   F.grid_sample(images, flow.permute(0,2,3,1), mode='bilinear')
// i.e. zeros padding, align_corners=False, flow = absolute sampling grid in
// [-1,1] with channel 0 = x and channel 1 = y.
// The arithmetic restated here is ATen's grid_sampler_2d:
//   ix = ((gx+1)*W-1)/2, corners (x0,y0)=floor, weights (x1-ix)(y1-iy)...,
//   taps outside the image contribute nothing; d ix / d gx = W/2.
//
// Execution plan: the flow is read straight from the network's (B,2,H,W)
// layout (no permuted copy).
//   Large fp32 maps of equal input/output size — TILED backward: grad_images is the row-owner
//   scatter (scatter_rows.cuh), the flow gradient the accumulate-then-weigh gather
//   (gather_quad.cuh).  The model's own maps ((8,64,128,128), (8,64,64,64)) take this path, the
//   32x32 crops do not.  The forward pass is always the direct kernel (nothing measured beats it
//   for four taps, see grid_warp_forward_t).
//   Otherwise — DIRECT kernels: one thread owns one output pixel and walks a
//   slice of channels with the four weights/offsets/validity bits in registers.
//   Backward is one fused pass: grad_images is a scatter (RED.ADD), grad_flow
//   is reduced over channels in registers + across the CTA's channel slices in
//   shared memory and stored once (ATen does the same per-pixel reduction but
//   serially over all C in one thread).
#include "common.cuh"
#include "scatter_tiled.cuh"
#include "scatter_rows.cuh"
#include "gather_quad.cuh"

namespace ffwm {

template <typename T>
struct Corner {
    int o[4];      // element offsets of nw, ne, sw, se inside one plane (0 when invalid)
    bool v[4];     // in-bounds flags
    T w[4];        // nw, ne, sw, se weights
    T wx0, wx1, wy0, wy1;
};

template <typename T>
__device__ __forceinline__ Corner<T> corners(T gx, T gy, int hi, int wi) {
    Corner<T> c;
    const T ix = ((gx + 1) * wi - 1) / 2;
    const T iy = ((gy + 1) * hi - 1) / 2;
    const T fx0 = floor(ix), fy0 = floor(iy);
    const int x0 = f2i(fx0), y0 = f2i(fy0);
    const int x1 = x0 + 1, y1 = y0 + 1;
    c.wx1 = T(x1) - ix; c.wx0 = ix - T(x0);
    c.wy1 = T(y1) - iy; c.wy0 = iy - T(y0);
    c.w[0] = c.wx1 * c.wy1; c.w[1] = c.wx0 * c.wy1;
    c.w[2] = c.wx1 * c.wy0; c.w[3] = c.wx0 * c.wy0;
    const bool vx0 = x0 >= 0 && x0 < wi, vx1 = x1 >= 0 && x1 < wi;
    const bool vy0 = y0 >= 0 && y0 < hi, vy1 = y1 >= 0 && y1 < hi;
    c.v[0] = vy0 && vx0; c.v[1] = vy0 && vx1; c.v[2] = vy1 && vx0; c.v[3] = vy1 && vx1;
    c.o[0] = y0; c.o[1] = x0; c.o[2] = y1; c.o[3] = x1;   // raw indices, turned into offsets by the caller
    return c;
}

template <typename T>
__device__ __forceinline__ void corner_offsets(const Corner<T>& c, int sh, int sw, int* o) {
    const int y0 = c.o[0], x0 = c.o[1], y1 = c.o[2], x1 = c.o[3];
    o[0] = c.v[0] ? y0 * sh + x0 * sw : 0;
    o[1] = c.v[1] ? y0 * sh + x1 * sw : 0;
    o[2] = c.v[2] ? y1 * sh + x0 * sw : 0;
    o[3] = c.v[3] ? y1 * sh + x1 * sw : 0;
}

// Tap list of one output pixel for the tiled scatter (scatter_tiled.cuh): the four bilinear
// corners; corners outside the image are skipped (zeros padding).
struct GridWarpScatterGeo {
    static constexpr int NT = 4;
    static constexpr int RW = 31;
    static constexpr int NW = 2;                 // scatter_rows.cuh: 2 x 2 window
    static constexpr bool CLAMP = false;         // zeros padding: taps outside the image are dropped
    View<const float> flow;
    int hi, wi;
    // scatter_rows.cuh: the 2x2 window in region coordinates; column weights (wx1, wx0), row weights (wy1, wy0)
    __device__ __forceinline__ bool window(int b, int y, int x, int rx0, int ry0, int rw, int rh, int& cb, int& rb, float* wx, float* wy) const {
        const float* f = flow.p + b * flow.sb + y * flow.sh + x * flow.sw;
        const float gx = __ldg(f), gy = __ldg(f + flow.sc);
        const float ix = ((gx + 1) * wi - 1) / 2;
        const float iy = ((gy + 1) * hi - 1) / 2;
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const bool near = fx0 >= float(rx0) && fx0 + 1.f <= float(rx0 + rw - 1) &&
                          fy0 >= float(ry0) && fy0 + 1.f <= float(ry0 + rh - 1);
        if (!near) return false;
        const int x0 = int(fx0), y0 = int(fy0);
        wx[0] = float(x0 + 1) - ix; wx[1] = ix - float(x0);
        wy[0] = float(y0 + 1) - iy; wy[1] = iy - float(y0);
        cb = x0 - rx0;
        rb = y0 - ry0;
        return true;
    }
    __device__ __forceinline__ void region_origin(int tx0, int ty0, int ml, int& rx0, int& ry0) const {
        rx0 = tx0 - ml;
        ry0 = ty0 - ml;
    }
    __device__ __forceinline__ void taps(int b, int y, int x, int* iy, int* ix, float* w) const {
        const float* f = flow.p + b * flow.sb + y * flow.sh + x * flow.sw;
        const Corner<float> cr = corners<float>(__ldg(f), __ldg(f + flow.sc), hi, wi);
        const int y0 = cr.o[0], x0 = cr.o[1], y1 = cr.o[2], x1 = cr.o[3];
        iy[0] = cr.v[0] ? y0 : -1; ix[0] = x0;
        iy[1] = cr.v[1] ? y0 : -1; ix[1] = x1;
        iy[2] = cr.v[2] ? y1 : -1; ix[2] = x0;
        iy[3] = cr.v[3] ? y1 : -1; ix[3] = x1;
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = cr.w[k];
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
grid_warp_fwd_kernel(View<const T> img, View<const T> flow, View<T> out, int c_per_block) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= out.h * out.w) return;
    const int b = blockIdx.z;
    const int y = pix / out.w, x = pix - y * out.w;
    const T* f = flow.p + b * flow.sb + y * flow.sh + x * flow.sw;
    const Corner<T> cr = corners<T>(ld_stream(f), ld_stream(f + flow.sc), img.h, img.w);
    int o[4];
    corner_offsets(cr, img.sh, img.sw, o);
    // zero weight + offset 0 for an out-of-image tap: branch-free zeros padding
    const T w0 = cr.v[0] ? cr.w[0] : T(0), w1 = cr.v[1] ? cr.w[1] : T(0);
    const T w2 = cr.v[2] ? cr.w[2] : T(0), w3 = cr.v[3] ? cr.w[3] : T(0);

    const int c0 = blockIdx.y * c_per_block;
    const int c1 = min(c0 + c_per_block, out.c);
    const T* s = img.plane(b, c0);
    T* d = out.plane(b, c0) + y * out.sh + x * out.sw;
#pragma unroll 4
    for (int c = c0; c < c1; ++c, s += img.sc, d += out.sc) {
        T v = T(0);
        v += __ldg(s + o[0]) * w0;
        v += __ldg(s + o[1]) * w1;
        v += __ldg(s + o[2]) * w2;
        v += __ldg(s + o[3]) * w3;
        st_stream(d, v);
    }
}

// ---- flow gradient, accumulate-then-weigh (gather_quad.cuh): M = sum_c grad_output[c] * (nw, ne, sw, se), then
// d out / d ix = -wy1 nw + wy1 ne - wy0 sw + wy0 se,  d out / d iy = -wx1 nw - wx0 ne + wx1 sw + wx0 se  once per pixel.
struct GwQuadPolicy {
    static constexpr int NW = 2;
    static constexpr bool PAD_ZERO = true;                   // zeros padding
    View<const float> img, flow, go;
    View<float> gflow;
    __host__ __device__ __forceinline__ const View<const float>& src() const { return img; }
    __host__ __device__ __forceinline__ const View<const float>& gout() const { return go; }
    // record: [1] x0 [2] y0 (bounded), [3] wx0 [4] wx1 [5] wy0 [6] wy1
    __device__ __forceinline__ int geometry(int b, int y, int x, int rx0, int ry0, float* rec) const {
        const float* f = flow.p + b * flow.sb + y * flow.sh + x * flow.sw;
        const float gx = __ldg(f), gy = __ldg(f + flow.sc);
        const float ix = ((gx + 1) * img.w - 1) / 2;
        const float iy = ((gy + 1) * img.h - 1) / 2;
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = f2i(fx0), y0 = f2i(fy0);
        int* ri = reinterpret_cast<int*>(rec);
        ri[1] = min(max(x0, -8), img.w + 8);
        ri[2] = min(max(y0, -8), img.h + 8);
        rec[3] = ix - float(x0);                 // wx0
        rec[4] = float(x0 + 1) - ix;             // wx1
        rec[5] = iy - float(y0);                 // wy0
        rec[6] = float(y0 + 1) - iy;             // wy1
        const bool near = fx0 >= float(rx0) && fx0 + 1.f <= float(rx0 + GQ_RW - 1) &&
                          fy0 >= float(ry0) && fy0 + 1.f <= float(ry0 + GQ_RH - 1);
        return near ? (y0 - ry0) * GQ_RW + (x0 - rx0) : -1;
    }
    __device__ __forceinline__ float far_tap(int iy, int ix, const float* plane) const {
        return ((unsigned)iy < (unsigned)img.h && (unsigned)ix < (unsigned)img.w) ? __ldg(plane + iy * img.sh + ix * img.sw) : 0.f;
    }
    __device__ __forceinline__ void finish(const float* M, const float* rec, int b, int y, int x) const {
        const float wx0 = rec[3], wx1 = rec[4], wy0 = rec[5], wy1 = rec[6];
        float gix = 0.f, giy = 0.f;
        gix -= M[0] * wy1; giy -= M[0] * wx1;
        gix += M[1] * wy1; giy -= M[1] * wx0;
        gix -= M[2] * wy0; giy += M[2] * wx1;
        gix += M[3] * wy0; giy += M[3] * wx0;
        float* o = gflow.p + b * gflow.sb + y * gflow.sh + x * gflow.sw;
        o[0] = (float(img.w) / 2) * gix;
        o[gflow.sc] = (float(img.h) / 2) * giy;
    }
};

template <typename T, int SL>
__global__ void __launch_bounds__(256)
grid_warp_bwd_kernel(View<const T> img, View<const T> flow, View<const T> gout,
                     View<T> gimg, View<T> gflow) {
    constexpr int PX = 256 / SL;
    __shared__ T red[SL > 1 ? SL : 1][2][PX];
    const int lane_px = threadIdx.x, slice = threadIdx.y;
    const int pix = blockIdx.x * PX + lane_px;
    const int b = blockIdx.z;
    const bool live = pix < gout.h * gout.w;
    const bool want_img = gimg.p != nullptr, want_flow = gflow.p != nullptr;

    T gix = T(0), giy = T(0);
    int y = 0, x = 0;
    if (live) {
        y = pix / gout.w;
        x = pix - y * gout.w;
        const T* f = flow.p + b * flow.sb + y * flow.sh + x * flow.sw;
        const Corner<T> cr = corners<T>(__ldg(f), __ldg(f + flow.sc), img.h, img.w);
        int so[4], dof[4];
        corner_offsets(cr, img.sh, img.sw, so);
        corner_offsets(cr, gimg.sh, gimg.sw, dof);
        // d out / d ix and d out / d iy coefficient of each corner value
        const T cx[4] = {-cr.wy1, cr.wy1, -cr.wy0, cr.wy0};
        const T cy[4] = {-cr.wx1, -cr.wx0, cr.wx1, cr.wx0};
        const int goff = y * gout.sh + x * gout.sw;
#pragma unroll 2
        for (int c = slice; c < gout.c; c += SL) {
            const T g = ld_stream(gout.plane(b, c) + goff);
            if (want_img) {
                T* d = gimg.plane(b, c);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (cr.v[k]) red_add(d + dof[k], cr.w[k] * g);
            }
            if (want_flow) {
                const T* s = img.plane(b, c);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (cr.v[k]) {
                        const T v = __ldg(s + so[k]);
                        gix += v * cx[k] * g;
                        giy += v * cy[k] * g;
                    }
            }
        }
    }

    if (!want_flow) return;
    if (SL > 1) {
        red[slice][0][lane_px] = gix;
        red[slice][1][lane_px] = giy;
        __syncthreads();
        if (slice != 0) return;
#pragma unroll
        for (int s = 1; s < SL; ++s) {
            gix += red[s][0][lane_px];
            giy += red[s][1][lane_px];
        }
    }
    if (!live) return;
    T* o = gflow.p + b * gflow.sb + y * gflow.sh + x * gflow.sw;
    o[0] = (T(img.w) / 2) * gix;
    o[gflow.sc] = (T(img.h) / 2) * giy;
}

template <typename T, int SL>
static void launch_bwd_sl(const View<const T>& img, const View<const T>& flow, const View<const T>& gout,
                          const View<T>& gi, const View<T>& gf, cudaStream_t st) {
    constexpr int PX = 256 / SL;
    dim3 grid(ceil_div((int64_t)gout.h * gout.w, PX), 1, gout.n), block(PX, SL);
    grid_warp_bwd_kernel<T, SL><<<grid, block, 0, st>>>(img, flow, gout, gi, gf);
}

template <typename T>
static int grid_warp_forward_t(const ffwm_tensor4* a, const ffwm_tensor4* b, const ffwm_tensor4* o, cudaStream_t st) {
    View<const T> img, flow;
    View<T> out;
    int rc;
    if ((rc = make_view<const T>(a, "images", &img))) return rc;
    if ((rc = make_view<const T>(b, "flow", &flow))) return rc;
    if ((rc = make_view<T>(o, "output", &out))) return rc;
    if (flow.c != 2) { set_error("grid_warp: flow needs 2 channels (x,y), got %d", flow.c); return FFWM_ERR_SHAPE; }
    if (out.n != flow.n || img.n != out.n || out.c != img.c || out.h != flow.h || out.w != flow.w) {
        set_error("grid_warp: output (%d,%d,%d,%d) inconsistent with images (%d,%d,..) / flow (%d,2,%d,%d)",
                  out.n, out.c, out.h, out.w, img.n, img.c, flow.n, flow.h, flow.w);
        return FFWM_ERR_SHAPE;
    }
    if ((int64_t)out.n * out.c * out.h * out.w == 0) return FFWM_OK;
    if (out.n > 65535) { set_error("grid_warp: batch %d > 65535", out.n); return FFWM_ERR_TOO_LARGE; }
    // Measured alternatives at the cfg5 point (0.37 ms direct): channel-lane gathers from a shared-memory slab —
    // tiled 0.68 ms, rolling strip 0.59 ms — and 128-bit corner-pair gathers 0.65 ms; with four taps the staging
    // costs more than the gather saves, so the forward pass is the direct kernel.
    const int pix_blocks = ceil_div((int64_t)out.h * out.w, 256);
    int64_t want = (int64_t)8 * sm_count();
    int chunks = int((want + (int64_t)pix_blocks * out.n - 1) / ((int64_t)pix_blocks * out.n));
    chunks = max(1, min(min(chunks, ceil_div(out.c, 4)), 65535));
    const int c_per_block = ceil_div(out.c, chunks);
    chunks = ceil_div(out.c, c_per_block);
    dim3 grid(pix_blocks, chunks, out.n);
    grid_warp_fwd_kernel<T><<<grid, 256, 0, st>>>(img, flow, out, c_per_block);
    return check_launch("grid_warp_forward");
}

template <typename T>
static int grid_warp_backward_t(const ffwm_tensor4* a, const ffwm_tensor4* b, const ffwm_tensor4* go,
                                const ffwm_tensor4* ga, const ffwm_tensor4* gb, cudaStream_t st) {
    View<const T> img, flow, gout;
    View<T> gi, gf;
    int rc;
    if ((rc = make_view<const T>(a, "images", &img))) return rc;
    if ((rc = make_view<const T>(b, "flow", &flow))) return rc;
    if ((rc = make_view<const T>(go, "grad_output", &gout))) return rc;
    if ((rc = make_view<T>(ga, "grad_images", &gi, true))) return rc;
    if ((rc = make_view<T>(gb, "grad_flow", &gf, true))) return rc;
    if (flow.c != 2) { set_error("grid_warp: flow needs 2 channels (x,y), got %d", flow.c); return FFWM_ERR_SHAPE; }
    if (gout.n != flow.n || img.n != gout.n || gout.c != img.c || gout.h != flow.h || gout.w != flow.w) {
        set_error("grid_warp_backward: grad_output (%d,%d,%d,%d) inconsistent with inputs", gout.n, gout.c, gout.h, gout.w);
        return FFWM_ERR_SHAPE;
    }
    if (gi.p && (gi.n != img.n || gi.c != img.c || gi.h != img.h || gi.w != img.w)) {
        set_error("grid_warp_backward: grad_images shape differs from images"); return FFWM_ERR_SHAPE;
    }
    if (gf.p && (gf.n != flow.n || gf.c != 2 || gf.h != flow.h || gf.w != flow.w)) {
        set_error("grid_warp_backward: grad_flow shape differs from flow"); return FFWM_ERR_SHAPE;
    }
    if ((!gi.p && !gf.p) || (int64_t)gout.n * gout.h * gout.w == 0) return FFWM_OK;
    if (gout.n > 65535) { set_error("grid_warp: batch %d > 65535", gout.n); return FFWM_ERR_TOO_LARGE; }
    if constexpr (sizeof(T) == 4) {
        // grad_images through the tiled scatter when the maps are large and of equal size (a flow
        // that is a perturbed identity then lands inside the tile's halo)
        if (gi.p && img.h == gout.h && img.w == gout.w && scatter_tiled_applicable(gout, gi)) {
            int rc2 = opt(OPT_SCATTER_TILED) ? launch_scatter_tiled(GridWarpScatterGeo{flow, img.h, img.w}, gout, gi, 7, st)
                                                   : launch_scatter_rows(GridWarpScatterGeo{flow, img.h, img.w}, gout, gi, st);
            if (rc2) return rc2;
            if ((rc2 = check_launch("grid_warp_backward(tiled scatter)"))) return rc2;
            if (!gf.p) return FFWM_OK;
            gi.p = nullptr;
        }
        if (!gi.p && gf.p && img.h == gout.h && img.w == gout.w && !opt(OPT_DISABLE_TILED_GFLOW) &&
            gather_quad_applicable(gout.n, gout.c, gout.h, gout.w, img)) {
            const int rc2 = launch_gather_quad(GwQuadPolicy{img, flow, gout, gf}, gout.n, gout.h, gout.w, st);
            if (rc2) return rc2;
            return check_launch("grid_warp_backward(quad flow gradient)");
        }
    }
    const int c = gout.c;
    if (c >= 8) launch_bwd_sl<T, 8>(img, flow, gout, gi, gf, st);
    else if (c >= 4) launch_bwd_sl<T, 4>(img, flow, gout, gi, gf, st);
    else if (c >= 2) launch_bwd_sl<T, 2>(img, flow, gout, gi, gf, st);
    else launch_bwd_sl<T, 1>(img, flow, gout, gi, gf, st);
    return check_launch("grid_warp_backward");
}

}  // namespace ffwm

extern "C" int ffwm_grid_warp_forward(const ffwm_tensor4* images, const ffwm_tensor4* flow,
                                      const ffwm_tensor4* output, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::grid_warp_forward_t<float>(images, flow, output, st);
    if (dtype == FFWM_F64) return ffwm::grid_warp_forward_t<double>(images, flow, output, st);
    ffwm::set_error("grid_warp_forward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}

extern "C" int ffwm_grid_warp_backward(const ffwm_tensor4* images, const ffwm_tensor4* flow,
                                       const ffwm_tensor4* grad_output, const ffwm_tensor4* grad_images,
                                       const ffwm_tensor4* grad_flow, int dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == FFWM_F32) return ffwm::grid_warp_backward_t<float>(images, flow, grad_output, grad_images, grad_flow, st);
    if (dtype == FFWM_F64) return ffwm::grid_warp_backward_t<double>(images, flow, grad_output, grad_images, grad_flow, st);
    ffwm::set_error("grid_warp_backward: unsupported dtype %d", dtype);
    return FFWM_ERR_ARG;
}
