// grid_warp on the rolling-strip gather (roll_gather.cuh): forward and flow gradient for fp32 maps of
// equal input/output size whose sampling grid is a perturbed identity (the generator's feature
// warps).  The ring is filled with the image ZERO-padded, so the four corners of a pixel whose 2x2
// window lies inside the ring need no validity test (zeros padding, align_corners=False).
// Included by grid_warp.cu after Corner / corners().
#pragma once
#include "roll_gather.cuh"

namespace ffwm {

constexpr int GWR_PW = 8;   // [0] off(y0,x0) | off(y1,x0) << 16   [1] slow flag   [4..7] weights / wx0 wx1 wy0 wy1

struct GwRollGeo {
    int offs;
    bool fast;
    float wx0, wx1, wy0, wy1;
};

// corners() arithmetic (ATen grid_sampler_2d), plus the ring test.
__device__ __forceinline__ GwRollGeo gw_roll_geometry(float gx, float gy, int hi, int wi, int rx0, int ystep) {
    GwRollGeo g;
    const float ix = ((gx + 1) * wi - 1) / 2;
    const float iy = ((gy + 1) * hi - 1) / 2;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = f2i(fx0), y0 = f2i(fy0);
    g.wx1 = float(x0 + 1) - ix; g.wx0 = ix - float(x0);
    g.wy1 = float(y0 + 1) - iy; g.wy0 = iy - float(y0);
    g.fast = fx0 >= float(rx0) && fx0 + 1.f <= float(rx0 + RG_RW - 1) &&
             fy0 >= float(ystep - RG_M) && fy0 + 1.f <= float(ystep + RG_SH + RG_M - 1);
    g.offs = 0;
    if (g.fast) {
        const int cb = x0 - rx0;
        g.offs = ((y0 & (RG_RING - 1)) * RG_RW + cb) | ((((y0 + 1) & (RG_RING - 1)) * RG_RW + cb) << 16);
    }
    return g;
}

// corner values of a pixel whose window left the ring: per-lane global loads, zeros outside the image
__device__ __noinline__ float4 gw_roll_slow_values(const View<const float>& img, const View<const float>& flow,
                                                   const float* plane_lane, int b, int y, int x, float4* wts) {
    const float* f = flow.p + b * flow.sb + y * flow.sh + x * flow.sw;
    const Corner<float> cr = corners<float>(__ldg(f), __ldg(f + flow.sc), img.h, img.w);
    int o[4];
    corner_offsets(cr, img.sh, img.sw, o);
    float4 v;
    v.x = cr.v[0] ? __ldg(plane_lane + o[0]) : 0.f;
    v.y = cr.v[1] ? __ldg(plane_lane + o[1]) : 0.f;
    v.z = cr.v[2] ? __ldg(plane_lane + o[2]) : 0.f;
    v.w = cr.v[3] ? __ldg(plane_lane + o[3]) : 0.f;
    *wts = make_float4(cr.wx0, cr.wx1, cr.wy0, cr.wy1);
    return v;
}

// MODE 0: forward (dst = output).  MODE 1: flow gradient (dst = grad_flow; CTA = strip x 128-row
// segment, walks the segment once per group of 32 channels).
template <int MODE>
__global__ void __launch_bounds__(RG_THREADS, 1)
grid_warp_roll_kernel(View<const float> img, View<const float> flow, View<const float> gout, View<float> dst, int seg_rows) {
    constexpr int SEG_STEPS = RG_SEG / RG_SH;
    extern __shared__ __align__(16) unsigned char rg_smem_raw[];
    float* slab = reinterpret_cast<float*>(rg_smem_raw);                       // [32][1025]
    float* prm_all = slab + 32 * RG_CHP;                                       // [16 warps][32 px][8]
    float* stage_all = prm_all + RG_WARPS * 32 * GWR_PW;                       // [16 warps][32][9]
    float* accs_all = stage_all + RG_WARPS * 32 * RG_SPITCH;                   // MODE 1: [16 warps][16 steps][8 px][2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int oh = MODE == 0 ? dst.h : gout.h, ow = MODE == 0 ? dst.w : gout.w, nc = MODE == 0 ? dst.c : gout.c;
    const int x0 = blockIdx.x * RG_SW, b = blockIdx.z;
    const int yseg = MODE == 0 ? 0 : blockIdx.y * seg_rows;
    const int seg_h = MODE == 0 ? oh : min(seg_rows, oh - yseg);
    const int nsteps = (seg_h + RG_SH - 1) / RG_SH;
    const int rx0 = x0 - RG_M;
    const int wrow = warp >> 1, xw0 = x0 + (warp & 1) * RG_PXW;
    float* prm = prm_all + warp * (32 * GWR_PW);
    float* stage = stage_all + warp * (32 * RG_SPITCH);
    float* acc = accs_all + warp * (SEG_STEPS * RG_PXW * 2);
    const float* slab_lane = slab + lane * RG_CHP;
    const int gsub = lane >> 3, gx = xw0 + (lane & 7);

    if (MODE == 1)
        for (int i = lane; i < SEG_STEPS * RG_PXW * 2; i += 32) acc[i] = 0.f;

    const int c_begin = MODE == 0 ? blockIdx.y * 32 : 0;
    const int c_end = MODE == 0 ? min(nc, c_begin + 32) : nc;
    for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        const int nch = min(32, nc - c0);
        const bool last_group = c0 + 32 >= c_end;
        const float* plane_lane = img.p + b * img.sb + (int64_t)(c0 + min(lane, nch - 1)) * img.sc;
        __syncthreads();                                   // the previous group's last step is done with the ring
        rg_fill_rows<true>(slab, img, b, c0, nch, rx0, yseg - RG_M, RG_M + 2 * RG_SH, warp, lane);

        float nfx = 0.f, nfy = 0.f;
        auto load_flow = [&](int s_base) {
            const int y = yseg + (s_base + gsub) * RG_SH + wrow;
            nfx = nfy = 0.f;
            if (y < oh && gx < ow) {
                const float* f = flow.p + b * flow.sb + y * flow.sh + gx * flow.sw;
                nfx = __ldg(f); nfy = __ldg(f + flow.sc);
            }
        };
        load_flow(0);
        float gl[8];
        if (MODE == 1) rg_load_row(gl, gout, b, c0, nch, yseg + wrow, xw0, lane);

        for (int s = 0; s < nsteps; ++s) {
            if ((s & (RG_BLK - 1)) == 0) {
                __syncwarp();
                const int ystep = yseg + (s + gsub) * RG_SH;
                const GwRollGeo g = gw_roll_geometry(nfx, nfy, img.h, img.w, rx0, ystep);
                float4* P4 = reinterpret_cast<float4*>(prm + lane * GWR_PW);
                P4[0] = make_float4(__int_as_float(g.offs), __int_as_float(g.fast ? 0 : 1), 0.f, 0.f);
                if (MODE == 0) P4[1] = make_float4(g.wx1 * g.wy1, g.wx0 * g.wy1, g.wx1 * g.wy0, g.wx0 * g.wy0);
                else P4[1] = make_float4(g.wx0, g.wx1, g.wy0, g.wy1);
                load_flow(s + RG_BLK);
                __syncwarp();
            }
            rg_cp_async_wait_all();
            __syncthreads();
            if (s + 1 < nsteps) rg_fill_rows<true>(slab, img, b, c0, nch, rx0, yseg + (s + 1) * RG_SH + RG_M, RG_SH, warp, lane);
            const int y = yseg + s * RG_SH + wrow;
            if (MODE == 1) {
                rg_stage_row(stage, gl, lane);
                __syncwarp();
                if (s + 1 < nsteps) rg_load_row(gl, gout, b, c0, nch, y + RG_SH, xw0, lane);
            }
            if (y < oh) {                                    // warp-uniform
                if (MODE == 0) {
#pragma unroll 4
                    for (int px = 0; px < RG_PXW; ++px) {
                        if (xw0 + px >= ow) break;           // warp-uniform
                        const float4* P4 = reinterpret_cast<const float4*>(prm + ((s & (RG_BLK - 1)) * RG_PXW + px) * GWR_PW);
                        const float4 h = P4[0];
                        float4 w4 = P4[1];
                        float4 v;
                        if (__float_as_int(h.y) == 0) {
                            const float* r0 = slab_lane + (__float_as_int(h.x) & 0xffff);
                            const float* r1 = slab_lane + (__float_as_int(h.x) >> 16);
                            v = make_float4(r0[0], r0[1], r1[0], r1[1]);
                        } else {
                            float4 wts;
                            v = gw_roll_slow_values(img, flow, plane_lane, b, y, xw0 + px, &wts);
                            w4 = make_float4(wts.y * wts.w, wts.x * wts.w, wts.y * wts.z, wts.x * wts.z);
                        }
                        float r = 0.f;
                        r += v.x * w4.x;
                        r += v.y * w4.y;
                        r += v.z * w4.z;
                        r += v.w * w4.w;
                        stage[lane * RG_SPITCH + px] = r;
                    }
                    __syncwarp();
                    rg_store_row(stage, dst, b, c0, nch, y, xw0, lane);
                } else {
#pragma unroll 1
                    for (int p4 = 0; p4 < RG_PXW / 4; ++p4) {
                        float red[8];
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const int px = p4 * 4 + kk;
                            red[2 * kk] = red[2 * kk + 1] = 0.f;
                            if (xw0 + px < ow) {             // warp-uniform
                                const float4* P4 = reinterpret_cast<const float4*>(prm + ((s & (RG_BLK - 1)) * RG_PXW + px) * GWR_PW);
                                const float4 h = P4[0];
                                float4 wts = P4[1];          // wx0 wx1 wy0 wy1
                                const float g = stage[lane * RG_SPITCH + px];
                                float4 v;
                                if (__float_as_int(h.y) == 0) {
                                    const float* r0 = slab_lane + (__float_as_int(h.x) & 0xffff);
                                    const float* r1 = slab_lane + (__float_as_int(h.x) >> 16);
                                    v = make_float4(r0[0], r0[1], r1[0], r1[1]);
                                } else {
                                    v = gw_roll_slow_values(img, flow, plane_lane, b, y, xw0 + px, &wts);
                                }
                                // d out / d ix = -wy1 nw + wy1 ne - wy0 sw + wy0 se ;  d out / d iy = -wx1 nw - wx0 ne + wx1 sw + wx0 se
                                float gix = 0.f, giy = 0.f;
                                gix -= v.x * wts.w; giy -= v.x * wts.y;
                                gix += v.y * wts.w; giy -= v.y * wts.x;
                                gix -= v.z * wts.z; giy += v.z * wts.y;
                                gix += v.w * wts.z; giy += v.w * wts.x;
                                red[2 * kk] = gix * g;
                                red[2 * kk + 1] = giy * g;
                            }
                        }
                        const float tot = gt_packed_reduce<8>(red, lane);        // lane l: value index l >> 2
                        if ((lane & 3) == 0) acc[(s * RG_PXW + p4 * 4) * 2 + (lane >> 2)] += tot;
                    }
                    if (last_group) {
                        __syncwarp();
                        const int x = xw0 + lane;
                        if (lane < RG_PXW && x < ow) {
                            float* o = dst.p + b * dst.sb + y * dst.sh + x * dst.sw;
                            o[0] = (float(img.w) / 2) * acc[(s * RG_PXW + lane) * 2];
                            o[dst.sc] = (float(img.h) / 2) * acc[(s * RG_PXW + lane) * 2 + 1];
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

template <int MODE>
static int launch_grid_warp_roll(const View<const float>& img, const View<const float>& flow,
                                 const View<const float>& gout, const View<float>& dst, int n, int c, int h, int w, cudaStream_t st) {
    const size_t smem = sizeof(float) * (32 * RG_CHP + RG_WARPS * 32 * GWR_PW + RG_WARPS * 32 * RG_SPITCH +
                                         (MODE == 1 ? RG_WARPS * (RG_SEG / RG_SH) * RG_PXW * 2 : 0));
    cudaError_t e = cudaFuncSetAttribute(grid_warp_roll_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("grid_warp_roll: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    const int seg = roll_segment_rows(n, h, w);
    dim3 grid(ceil_div(w, RG_SW), MODE == 0 ? ceil_div(c, 32) : ceil_div(h, seg), n);
    grid_warp_roll_kernel<MODE><<<grid, RG_THREADS, smem, st>>>(img, flow, gout, dst, seg);
    return FFWM_OK;
}

}  // namespace ffwm
