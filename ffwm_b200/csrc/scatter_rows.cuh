// Row-owner scatter-add for the warp backward passes (grad wrt the sampled tensor), second
// generation of scatter_tiled.cuh.
//
// scatter_tiled.cuh sorts every TAP by destination and walks the sorted stream with a "last entry
// of this destination" flag: ~13 instructions per tap and channel group, 300 warp-instructions per
// pixel and 32 channels for a 16-tap op (ncu, profiles/r01m_ncu_hot.txt).  The taps of one pixel
// are not independent, though: they form an NW x NW window of CONSECUTIVE destination columns and
// rows whose weights are separable (row weight x column weight).  So here
//
//   * the unit of work is a window ROW (pixel, i): NW taps on consecutive columns of one
//     destination row.  The rows of a tile's 31x31 destination region are OWNED by warps (warp w owns
//     rows w and w+16), and the window rows are bucketed by destination row once per tile
//     (counting sort of 256*NW items instead of 256*NW*NW entries);
//   * lanes are channels; the owner accumulates its destination row — 31 columns x 32 channels — in a
//     warp-private shared-memory buffer with plain read-modify-writes: the lanes of a warp never
//     collide (different channels) and no other warp touches the row, so no atomics and no "last
//     entry" bookkeeping are needed; per item: one broadcast item (pixel, column, row weight), the
//     pixel's column weights (one 128-bit broadcast) and the lane's grad_output.  (Keeping the row in
//     registers behind a jump table on the column was measured slower: the compare tree costs more
//     than the shared-memory round trips.)
//   * a finished row leaves the buffer as COALESCED red.adds (31 consecutive columns of one channel
//     per instruction) and the buffer is zeroed on the way out.  The region is in VIRTUAL coordinates:
//     columns/rows outside the image are clamped (resample2d: the reference clamps every tap index)
//     or dropped (grid_warp: zeros padding) when the row is flushed, so windows need no per-tap
//     clamping;
//   * grad_output of the NEXT channel group is prefetched into registers while this one is walked.
// Pixels whose window leaves the region (|displacement| > ~6 px) keep the per-tap far list of
// scatter_tiled.cuh (direct REDs, lanes are channels).
#pragma once
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "scatter_tiled.cuh"

namespace ffwm {

constexpr int SR_RW = 31;                       // destination region: 31 x 31 around the 16x16 tile
constexpr int SR_BPITCH = 33;                   // transpose buffer [32 channels][33]

template <int NW>
struct SrSmem {
    static constexpr size_t bytes() {
        return sizeof(float) * (32 * ST_GPITCH)                     // G: grad_output tile [32][257]
               + sizeof(float) * (ST_WARPS * 32 * SR_BPITCH)        // per-warp transpose buffers
               + sizeof(float) * (ST_NPX * NW)                      // per-pixel column weights
               + sizeof(int2) * (ST_NPX * NW)                       // items {pixel | column << 8, row weight}, bucketed by destination row
               + sizeof(StEntry) * (ST_NPX * NW * NW)               // far taps
               + sizeof(int) * (2 * 32 + 8);                        // counts, offsets, misc
    }
};

// Geo (see scatter_tiled.cuh) plus:
//   Geo::NW, Geo::CLAMP
//   Geo::window(b, y, x, rx0, ry0, &cb, &rb, wx[NW], wy[NW]) -> true when the pixel's NW x NW window lies inside
//   the region; cb/rb = region column/row of the window's first column/row, wx/wy = column / row weights by
//   window position (their product is the tap weight).
template <class Geo>
__global__ void __launch_bounds__(ST_THREADS, 1)
scatter_rows_kernel(Geo geo, View<const float> gout, View<float> gsrc, int ml) {
    constexpr int NW = Geo::NW, NT = NW * NW, RW = SR_RW;
    static_assert(Geo::NT == NT && Geo::RW == RW, "window and tap list must agree");
    extern __shared__ __align__(16) unsigned char sr_smem_raw[];
    float* G = reinterpret_cast<float*>(sr_smem_raw);                          // [32][257]
    float* bufs = G + 32 * ST_GPITCH;                                          // [16 warps][32][33]
    float* wts = bufs + ST_WARPS * 32 * SR_BPITCH;                             // [256][NW] column weights
    int2* items = reinterpret_cast<int2*>(wts + ST_NPX * NW);                  // [256*NW]
    StEntry* far = reinterpret_cast<StEntry*>(items + ST_NPX * NW);            // [256*NT]
    int* cnt = reinterpret_cast<int*>(far + ST_NPX * NT);                      // [32]
    int* off = cnt + 32;                                                       // [32]
    int* misc = off + 32;                                                      // [0] = far taps

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * ST_TW, ty0 = blockIdx.y * ST_TH, b = blockIdx.z;
    int rx0, ry0;
    geo.region_origin(tx0, ty0, ml, rx0, ry0);

    if (tid < 32) cnt[tid] = 0;
    if (tid == 0) misc[0] = 0;
    for (int i = tid; i < ST_WARPS * 32 * SR_BPITCH; i += ST_THREADS) bufs[i] = 0.f;
    __syncthreads();

    // ---- 1. geometry, once per tile: window rows bucketed by destination row -------------------
    int my_cb = 0, my_rb = 0, my_slot[NW];
    float my_wy[NW];
    bool near = false;
    if (tid < ST_NPX) {
        const int y = ty0 + tid / ST_TW, x = tx0 + tid % ST_TW;
        if (y < gout.h && x < gout.w) {
            float wx[NW], wy[NW];
            near = geo.window(b, y, x, rx0, ry0, my_cb, my_rb, wx, wy);
            if (near) {
#pragma unroll
                for (int i = 0; i < NW; ++i) {
                    wts[tid * NW + i] = wx[i];
                    my_wy[i] = wy[i];
                    my_slot[i] = atomicAdd(&cnt[my_rb + i], 1);
                }
            } else {
                int iy[NT], ix[NT];
                float w[NT];
                geo.taps(b, y, x, iy, ix, w);
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    if (iy[t] < 0) continue;
                    StEntry e;
                    e.w = w[t];
                    e.p = tid | ((iy[t] * gsrc.sh + ix[t] * gsrc.sw) << 8);
                    far[atomicAdd(&misc[0], 1)] = e;
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) {                                          // exclusive scan of the 31 row counts
        const int v = lane < RW ? cnt[lane] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += u;
        }
        off[lane] = inc - v;
    }
    __syncthreads();
    if (near) {
#pragma unroll
        for (int i = 0; i < NW; ++i) items[off[my_rb + i] + my_slot[i]] = make_int2(tid | (my_cb << 8), __float_as_int(my_wy[i]));
    }
    const int nfar = misc[0];

    // grad_output staging: thread -> (channel tid/16, column tid%16), 16 rows; next group in registers
    const int gc = tid / ST_TW, gxl = tid % ST_TW;
    float gnext[ST_TH];
    auto load_g = [&](int c0) {
        const int nch = min(32, gout.c - c0);
        const bool ok = gc < nch && tx0 + gxl < gout.w;
        const float* gp = gout.p + b * gout.sb + (int64_t)(c0 + gc) * gout.sc + (tx0 + gxl) * gout.sw;
#pragma unroll
        for (int r = 0; r < ST_TH; ++r) {
            const int y = ty0 + r;
            gnext[r] = (ok && y < gout.h) ? ld_stream(gp + y * gout.sh) : 0.f;
        }
    };
    load_g(0);

    float* buf = bufs + warp * (32 * SR_BPITCH);
    const float* Gl = G + lane * ST_GPITCH;
    const int gx = rx0 + lane;                                // flush: lane = region column
    const int gxc = Geo::CLAMP ? min(max(gx, 0), gsrc.w - 1) : gx;
    const bool col_ok = lane < RW && (Geo::CLAMP || (unsigned)gx < (unsigned)gsrc.w);

    // ---- 2. channel groups -------------------------------------------------------------------
    for (int c0 = 0; c0 < gout.c; c0 += 32) {
        const int nch = min(32, gout.c - c0);
        __syncthreads();                                      // items written / previous group's G consumed
#pragma unroll
        for (int r = 0; r < ST_TH; ++r) G[gc * ST_GPITCH + r * ST_TW + gxl] = gnext[r];
        __syncthreads();
        if (c0 + 32 < gout.c) load_g(c0 + 32);

        for (int row = warp; row < RW; row += ST_WARPS) {
            const int i_beg = off[row], i_end = i_beg + cnt[row];
            if (i_beg == i_end) continue;                     // warp-uniform
            // the owner accumulates its destination row in its private buffer: lanes are channels, so the
            // read-modify-writes of a warp never collide, and no other warp touches this row
            float* bl = buf + lane * SR_BPITCH;
#pragma unroll 2
            for (int it = i_beg; it < i_end; ++it) {
                const int2 item = items[it];
                const int pix = item.x & 255;
                float* d = bl + (item.x >> 8);
                const float gw = Gl[pix] * __int_as_float(item.y);
                if (NW == 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wts + pix * NW);
                    d[0] += gw * w4.x; d[1] += gw * w4.y; d[NW - 2] += gw * w4.z; d[NW - 1] += gw * w4.w;
                } else {
                    const float2 w2 = *reinterpret_cast<const float2*>(wts + pix * NW);
                    d[0] += gw * w2.x; d[1] += gw * w2.y;
                }
            }
            __syncwarp();
            // one coalesced RED per channel; the buffer is left zeroed for the next row
            const int gy = ry0 + row;
            const bool row_ok = Geo::CLAMP || (unsigned)gy < (unsigned)gsrc.h;
            const int gyc = Geo::CLAMP ? min(max(gy, 0), gsrc.h - 1) : gy;
            float* gp = gsrc.p + b * gsrc.sb + (int64_t)c0 * gsrc.sc + gyc * gsrc.sh + gxc * gsrc.sw;
            if (lane < RW) {
#pragma unroll 4
                for (int c = 0; c < 32; ++c) {
                    const float v = buf[c * SR_BPITCH + lane];
                    buf[c * SR_BPITCH + lane] = 0.f;
                    if (v != 0.f && c < nch && row_ok && col_ok) red_add(gp + (int64_t)c * gsrc.sc, v);
                }
            }
            __syncwarp();
        }
        // far taps: direct REDs, lanes are channels
        for (int k = warp; k < nfar; k += ST_WARPS) {
            const StEntry en = far[k];
            const int p = en.p & 255, go = en.p >> 8;
            if (lane < nch) red_add(gsrc.p + b * gsrc.sb + (int64_t)(c0 + lane) * gsrc.sc + go, en.w * Gl[p]);
        }
    }
}

template <class Geo>
static int launch_scatter_rows(const Geo& geo, const View<const float>& gout, const View<float>& gsrc, int ml, cudaStream_t st) {
    const size_t smem = SrSmem<Geo::NW>::bytes();
    cudaError_t e = cudaFuncSetAttribute(scatter_rows_kernel<Geo>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("scatter_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(ceil_div(gout.w, ST_TW), ceil_div(gout.h, ST_TH), gout.n);
    scatter_rows_kernel<Geo><<<grid, ST_THREADS, smem, st>>>(geo, gout, gsrc, ml);
    return FFWM_OK;
}

}  // namespace ffwm
