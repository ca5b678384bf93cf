// Flow-gradient gathers, third generation: accumulate FIRST, weigh LAST.
//
// The flow gradient of a warp op is, per output pixel p,  sum_c g[c,p] * F(window of src[c] at p)  where F
// is a fixed set of bilinear forms of the NW x NW window whose coefficients depend on the pixel only.
// gather_tiled.cuh / roll_gather.cuh evaluate F per (pixel, channel) — ~70 FFMA plus a cross-lane
// reduction per pixel and 32 channels, 240 warp-instructions per pixel and 32 channels (ncu) — and are
// instruction-bound.  But F is linear in the window, so
//
//       sum_c g[c,p] * F(V_c)  =  F( sum_c g[c,p] * V_c )        V_c = the NW x NW window of channel c
//
// i.e. ONE FFMA per tap and channel into NW*NW per-pixel accumulators M, and F — the Gaussian /
// bilinear weights, the normaliser, the reference's SAFE_DIVs — once per pixel at the very end.
//
// Execution: a CTA owns a 16x16 tile of the output grid (512 threads, one CTA per SM) and stages, 16 channels
// at a time and double-buffered (the copies of the next 16 channels fly while these are accumulated), the
// 31 x 32 halo region of the source — PADDED the way the op pads (edge
// replication / zeros), so windows inside the region need no per-tap clamping: tap (i,j) is
// `base + i*32 + j` — with 128-bit cp.async where the layout allows.  A warp owns one tile row (16
// pixels); its lanes are 8 channels x 4 pixels: lane (c8, p4) accumulates M for the pixels 4q + p4,
// q = 0..3, of the row (4 x NW*NW registers) over the channels = c8 (mod 8) of every group.  Channel
// pitch = 4 (mod 32) words: the 8 channels of a tap load fall on banks 4 apart, the 4 pixels collide
// only when their offsets agree mod 4 (measured: ~2 wavefronts per load).  After the last group M is
// summed over the 8 channel lanes (3 shuffles per value) and one lane per pixel applies F.
// Pixels whose window leaves the region (|displacement| > ~6 px) read their taps from global memory.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace ffwm {

constexpr int GQ_TW = 16, GQ_TH = 16, GQ_NPX = 256;
constexpr int GQ_THREADS = 512, GQ_WARPS = 16;
constexpr int GQ_RH = 31, GQ_RW = 32;                 // staged region: 31 rows x 32 columns ...
constexpr int GQ_MX = 8, GQ_MY = 7;                   // ... starting 8 columns left of / 7 rows above the tile
constexpr int GQ_CHP = GQ_RH * GQ_RW + 4;             // channel pitch 996 = 4 (mod 32), 16-byte aligned
constexpr int GQ_GP = GQ_NPX + 4;                     // grad_output pitch 260 = 4 (mod 32)
constexpr int GQ_REC = 16;                            // floats of per-pixel record

__device__ __forceinline__ unsigned gq_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// slab[c][31][32] <- src[b, c0+c, ry0.., rx0..] padded (PAD_ZERO: zeros outside the image, else edge replication).
// vec: unit column stride, rows 16-byte aligned and width % 4 == 0 -> a 4-column chunk is entirely inside or
// outside the image and goes with one 16-byte cp.async; lane -> (row lane/8 of four, chunk lane%8).
template <bool PAD_ZERO>
__device__ __forceinline__ void gq_fill_half(float* slab, const View<const float>& src, int b, int c0, int nch,
                                             int rx0, int ry0, int warp, int lane, bool vec) {
    // one HALF of a channel group: 16 channels, warp w copies channel c0 + w
    const int c = warp;
    const float* plane = src.p + b * src.sb + (int64_t)(c0 + min(c, max(nch, 1) - 1)) * src.sc;   // channels past nch repeat the last one
    float* sc = slab + c * GQ_CHP;
    if (vec) {
        const int rsub = lane >> 3, ch4 = lane & 7;
        const int gx4 = rx0 + 4 * ch4;
        const bool col_in = (unsigned)gx4 < (unsigned)src.w;
        const int gxb = gx4 < 0 ? 0 : src.w - 1;
#pragma unroll 2
        for (int r = rsub; r < GQ_RH; r += 4) {
            const int gy = ry0 + r;
            const bool row_in = (unsigned)gy < (unsigned)src.h;
            float* d = sc + r * GQ_RW + 4 * ch4;
            if (PAD_ZERO) {
                if (row_in && col_in)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gq_smem_u32(d)), "l"(plane + gy * src.sh + gx4) : "memory");
                else
                    *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                const float* rowp = plane + min(max(gy, 0), src.h - 1) * src.sh;
                if (col_in) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gq_smem_u32(d)), "l"(rowp + gx4) : "memory");
                } else {
                    const float v = __ldg(rowp + gxb);
                    *reinterpret_cast<float4*>(d) = make_float4(v, v, v, v);
                }
            }
        }
    } else {
        const int gx = rx0 + lane;
        const bool col_in = (unsigned)gx < (unsigned)src.w;
        const int gxc = min(max(gx, 0), src.w - 1);
#pragma unroll 2
        for (int r = 0; r < GQ_RH; ++r) {
            const int gy = ry0 + r;
            float* d = sc + r * GQ_RW + lane;
            if (PAD_ZERO) {
                if (col_in && (unsigned)gy < (unsigned)src.h)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(gq_smem_u32(d)), "l"(plane + gy * src.sh + gx * src.sw) : "memory");
                else
                    *d = 0.f;
            } else {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(gq_smem_u32(d)),
                             "l"(plane + min(max(gy, 0), src.h - 1) * src.sh + gxc * src.sw) : "memory");
            }
        }
    }
}

// G[c][260] <- t[b, c0+c, tile] for 16 channels; warp -> channel, lane -> (column lane%16, row parity lane/16); zeros outside.
__device__ __forceinline__ void gq_fill_tile_half(float* G, const View<const float>& t, int b, int c0, int nch,
                                                  int ty0, int tx0, int warp, int lane) {
    const int c = warp, x = lane & 15, r0 = lane >> 4;
    const bool ok = c < nch && tx0 + x < t.w;
    const float* gp = t.p + b * t.sb + (int64_t)(c0 + min(c, max(nch, 1) - 1)) * t.sc + (tx0 + x) * t.sw;
    float* d = G + c * GQ_GP + x;
#pragma unroll
    for (int k = 0; k < GQ_TH / 2; ++k) {
        const int r = r0 + 2 * k, y = ty0 + r;
        if (ok && y < t.h) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(gq_smem_u32(d + r * GQ_TW)), "l"(gp + y * t.sh) : "memory");
        else d[r * GQ_TW] = 0.f;
    }
}

// Policy P:
//   P::NW                      window size (2 or 4);  P::PAD_ZERO
//   P::geometry(b, y, x, rx0, ry0, rec[16]) -> region offset of the window's first tap (row*32 + col) or -1 when the
//                              window is not inside the region; fills the pixel's record (whatever finish() needs;
//                              rec[1], rec[2] = the UNCLAMPED integer column / row of the window's first tap)
//   P::far_tap(iy, ix, plane)  value of tap (iy, ix) (unclamped integer coordinates) read from global memory
//   P::finish(M[NW*NW], rec, b, y, x)   applies the bilinear forms and stores the pixel's gradient
//   src(), gout() views
template <class P>
__global__ void __launch_bounds__(GQ_THREADS, 1)
gather_quad_kernel(P pol, int vec) {
    constexpr int NW = P::NW, NT = NW * NW;
    extern __shared__ __align__(16) unsigned char gq_smem_raw[];
    float* slab = reinterpret_cast<float*>(gq_smem_raw);                 // [32][996]
    float* G = slab + 32 * GQ_CHP;                                       // [32][260]
    float* recs = G + 32 * GQ_GP;                                        // [256][16]
    const View<const float>& src = pol.src();
    const View<const float>& gout = pol.gout();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * GQ_TW, ty0 = blockIdx.y * GQ_TH, b = blockIdx.z;
    const int rx0 = tx0 - GQ_MX, ry0 = ty0 - GQ_MY;

    if (tid < GQ_NPX) {
        const int y = ty0 + tid / GQ_TW, x = tx0 + tid % GQ_TW;
        float* rec = recs + tid * GQ_REC;
        int off = -2;                                                    // -2: pixel outside the image
        if (y < gout.h && x < gout.w) off = pol.geometry(b, y, x, rx0, ry0, rec);
        reinterpret_cast<int*>(rec)[0] = off;
    }
    __syncthreads();

    // lane -> (channel c8 of eight, pixel slot p4 of four); the warp's tile row is processed as 4 quads of pixels.
    // The 4 pixels of a quad hit the same bank when their window offsets agree mod 4 (the 8 channels sit 4 banks
    // apart), so the row's 16 pixels are dealt into the quads by (rank inside their offset class, class): a quad
    // then holds one pixel of each class as long as the classes last (smooth flows: consecutive pixels already do).
    const int c8 = lane & 7, p4 = lane >> 3;
    const int y = ty0 + warp;
    int* perm = reinterpret_cast<int*>(slab + 32 * GQ_CHP + 32 * GQ_GP + GQ_NPX * GQ_REC) + warp * GQ_TW;   // [16 warps][16]
    {
        const int myoff = lane < GQ_TW ? reinterpret_cast<const int*>(recs + (warp * GQ_TW + lane) * GQ_REC)[0] : -2;
        const int cls = (myoff >= 0 ? myoff : lane) & 3;
        unsigned m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = __ballot_sync(0xffffffffu, lane < GQ_TW && cls == k);
        const unsigned below = (1u << lane) - 1u;
        const unsigned mine = cls == 0 ? m[0] : cls == 1 ? m[1] : cls == 2 ? m[2] : m[3];
        const int rank = __popc(mine & below);
        int pos = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) pos += min(__popc(m[k]), rank) + ((k < cls && __popc(m[k]) > rank) ? 1 : 0);
        if (lane < GQ_TW) perm[pos] = lane;
    }
    __syncwarp();
    int off[4], pix[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        pix[q] = perm[4 * q + p4];
        off[q] = reinterpret_cast<const int*>(recs + (warp * GQ_TW + pix[q]) * GQ_REC)[0];
    }
    float M[4][NT];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < NT; ++t) M[q][t] = 0.f;
    const bool all_near = __all_sync(0xffffffffu, off[0] >= 0 && off[1] >= 0 && off[2] >= 0 && off[3] >= 0);

    // Channel groups go through the shared memory in HALVES of 16 channels, double-buffered: while half h is
    // accumulated, the cp.asyncs of half h+1 are in flight and those of half h+2 are issued as soon as h is done.
    const int nhalf = (gout.c + 15) / 16;
    auto issue = [&](int h) {                                            // fill buffer h & 1 with channels [16h, 16h+16)
        if (h < nhalf) {
            const int c0 = 16 * h, nch = min(16, gout.c - c0);
            gq_fill_half<P::PAD_ZERO>(slab + (h & 1) * 16 * GQ_CHP, src, b, c0, nch, rx0, ry0, warp, lane, vec != 0);
            gq_fill_tile_half(G + (h & 1) * 16 * GQ_GP, gout, b, c0, nch, ty0, tx0, warp, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");             // one group per half, empty past the end
    };
    issue(0);
    issue(1);
    for (int h = 0; h < nhalf; ++h) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");             // this thread's copies of half h have landed
        __syncthreads();                                                 // ... and everybody else's
        const int c0 = 16 * h, nch = min(16, gout.c - c0);
        const float* slab_h = slab + (h & 1) * 16 * GQ_CHP;
        const float* G_h = G + (h & 1) * 16 * GQ_GP;
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            const int ch = s * 8 + c8;
            const float* sl = slab_h + ch * GQ_CHP;
            const float* gl = G_h + ch * GQ_GP + warp * GQ_TW;
            if (all_near) {                                              // warp-uniform: no per-pixel branches
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float g = gl[pix[q]];                          // zero for channels past nch
                    const float* w = sl + off[q];
#pragma unroll
                    for (int i = 0; i < NW; ++i)
#pragma unroll
                        for (int j = 0; j < NW; ++j) M[q][i * NW + j] = fmaf(g, w[i * GQ_RW + j], M[q][i * NW + j]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float g = gl[pix[q]];                          // zero for channels past nch / pixels outside
                    if (off[q] >= 0) {
                        const float* w = sl + off[q];
#pragma unroll
                        for (int i = 0; i < NW; ++i)
#pragma unroll
                            for (int j = 0; j < NW; ++j) M[q][i * NW + j] = fmaf(g, w[i * GQ_RW + j], M[q][i * NW + j]);
                    } else if (off[q] == -1) {
                        const int* rec = reinterpret_cast<const int*>(recs + (warp * GQ_TW + pix[q]) * GQ_REC);
                        const int ifx = rec[1], ify = rec[2];
                        const float* plane = src.p + b * src.sb + (int64_t)(c0 + min(ch, nch - 1)) * src.sc;
#pragma unroll
                        for (int i = 0; i < NW; ++i)                     // fully unrolled: M must stay in registers
#pragma unroll
                            for (int j = 0; j < NW; ++j)
                                M[q][i * NW + j] = fmaf(g, pol.far_tap(ify + i, ifx + j, plane), M[q][i * NW + j]);
                    }
                }
            }
        }
        __syncthreads();                                                 // buffer h & 1 is free again
        issue(h + 2);
    }
    // sum over the 8 channel lanes; lane c8 == 0 finishes pixel 4q + p4
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            float v = M[q][t];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            M[q][t] = v;
        }
    }
    // the 16 pixels of the row are finished by 16 lanes in parallel: totals go through the (now idle) slab
    __syncthreads();                                                     // every warp is done with the slab
    float* tot = slab + warp * (GQ_TW * (NT + 1));                       // [16 pixels][NT + 1]
    if (c8 == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int t = 0; t < NT; ++t) tot[pix[q] * (NT + 1) + t] = M[q][t];
    }
    __syncwarp();
    if (lane < GQ_TW && y < gout.h && tx0 + lane < gout.w) {
        float Mp[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) Mp[t] = tot[lane * (NT + 1) + t];
        pol.finish(Mp, recs + (warp * GQ_TW + lane) * GQ_REC, b, y, tx0 + lane);
    }
}

template <class P>
static int launch_gather_quad(const P& pol, int n, int h, int w, cudaStream_t st) {
    const size_t smem = sizeof(float) * (32 * GQ_CHP + 32 * GQ_GP + GQ_NPX * GQ_REC + GQ_WARPS * GQ_TW);
    cudaError_t e = cudaFuncSetAttribute(gather_quad_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gather_quad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    const View<const float>& src = pol.src();
    const int vec = src.sw == 1 && (src.w & 3) == 0 && (src.sh & 3) == 0 && (src.sc & 3) == 0 && (src.sb & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(src.p) & 15) == 0 && !opt(OPT_GQ_SCALAR_FILL);
    dim3 grid(ceil_div(w, GQ_TW), ceil_div(h, GQ_TH), n);
    gather_quad_kernel<P><<<grid, GQ_THREADS, smem, st>>>(pol, vec);
    return FFWM_OK;
}

// (A FORWARD kernel on the same staging — the NW*NW tap weights of a lane's four pixels in registers, one shared
// load + one FFMA per tap, outputs stored straight from the (8 channels x 4 pixels) lanes — was parity-green but
// slower than the rolling forward at the cfg5 point: 1.00 vs 0.89 ms for 16 taps, 0.62 vs 0.52 ms (direct) for 4;
// its 4-byte stores scatter over 8 channel planes.  Removed; a version that transposes the outputs through shared
// memory is round-2 work.)

inline bool gather_quad_applicable(int n, int c, int h, int w, const View<const float>& src) {
    if (opt(OPT_DISABLE_TILED) || opt(OPT_DISABLE_QUAD)) return false;
    if (src.sh < 0 || src.sw < 0 || c < 16 || n > 65535) return false;
    if ((int64_t)(src.h - 1) * src.sh + (int64_t)(src.w - 1) * src.sw >= (1 << 30)) return false;
    if (ceil_div(h, GQ_TH) > 65535) return false;
    if (opt(OPT_FORCE_TILED)) return true;          // tests: small ragged shapes through the tiled kernels
    const int64_t tiles = (int64_t)ceil_div(w, GQ_TW) * ceil_div(h, GQ_TH) * n;
    return tiles >= sm_count() / 2;
}

}  // namespace ffwm
