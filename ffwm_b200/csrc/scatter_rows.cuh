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
//     (counting sort of 256*NW items by (row, first column) instead of 256*NW*NW entries);
//   * lanes are channels, so the lanes of a warp never collide, and no other warp touches the row: no
//     atomics and no "last entry" bookkeeping.  The items of a row arrive sorted by first column, so
//     the NW columns under the current item are a sliding window held in REGISTERS: a tap is one
//     FFMA, and a column is stored to the warp's row buffer exactly once, when the window moves past
//     it (on average a move of 0.6 columns per item).  Per item: one broadcast item (pixel, column,
//     row weight), the pixel's column weights (one 128-bit broadcast) and the lane's grad_output.
//     (Measured alternatives: read-modify-write of the row in shared memory 1.33 ms, the whole row in
//     registers behind a jump table on the column 1.83 ms, scatter_tiled.cuh 1.73 ms — ks4, cfg5.)
//   * a finished row leaves the buffer as COALESCED red.adds (31 consecutive columns of one channel
//     per instruction, predicated on a non-zero value) and the buffer is zeroed on the way out.  The region is in VIRTUAL coordinates:
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
constexpr int SR_BPITCH = 33;                   // row buffer [32 channels][33]
constexpr int SR_KEYS = SR_RW * 32;             // sort key = destination row * 32 + first column

template <int NW>
struct SrSmem {
    static constexpr size_t bytes() {
        return sizeof(float) * (32 * ST_GPITCH)                     // G: grad_output tile [32][257]
               + sizeof(float) * (ST_WARPS * 32 * SR_BPITCH)        // per-warp transpose buffers
               + sizeof(float) * (ST_NPX * NW)                      // per-pixel column weights
               + sizeof(int2) * (ST_NPX * NW)                       // items {pixel | column << 8, row weight}, bucketed by destination row
               + sizeof(StEntry) * (ST_NPX * NW * NW)               // far taps
               + sizeof(int) * (2 * (SR_KEYS + 1) + 32);            // counts, offsets over (row, column) keys, misc
    }
};

// red.add of a non-zero value, predicated instead of branched
__device__ __forceinline__ void sr_red_nonzero(float* p, float v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.neu.f32 q, %1, 0f00000000;\n\t@q red.global.add.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v) : "memory");
}

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
    int* cnt = reinterpret_cast<int*>(far + ST_NPX * NT);                      // [SR_KEYS + 1]
    int* off = cnt + (SR_KEYS + 1);                                            // [SR_KEYS + 1]
    int* misc = off + (SR_KEYS + 1);                                           // [0] = far taps, [1..16] = scan partials

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * ST_TW, ty0 = blockIdx.y * ST_TH, b = blockIdx.z;
    int rx0, ry0;
    geo.region_origin(tx0, ty0, ml, rx0, ry0);

    for (int i = tid; i < SR_KEYS + 1; i += ST_THREADS) cnt[i] = 0;
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
                    my_slot[i] = atomicAdd(&cnt[(my_rb + i) * 32 + my_cb], 1);
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
    // exclusive scan of cnt[0..SR_KEYS] -> off (two keys per thread, warp scan, scan of the warp totals)
    {
        constexpr int PER = (SR_KEYS + 1 + ST_THREADS - 1) / ST_THREADS;
        int a[PER], sum = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = PER * tid + k;
            a[k] = i <= SR_KEYS ? cnt[i] : 0;
            sum += a[k];
        }
        int inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        if (lane == 31) misc[1 + warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int v = lane < ST_WARPS ? misc[1 + lane] : 0;
            int winc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += u;
            }
            if (lane < ST_WARPS) misc[1 + lane] = winc - v;
        }
        __syncthreads();
        int ex = misc[1 + warp] + inc - sum;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = PER * tid + k;
            if (i <= SR_KEYS) off[i] = ex;
            ex += a[k];
        }
    }
    __syncthreads();
    if (near) {
#pragma unroll
        for (int i = 0; i < NW; ++i)
            items[off[(my_rb + i) * 32 + my_cb] + my_slot[i]] = make_int2(tid | (my_cb << 8), __float_as_int(my_wy[i]));
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
        if (tid == 0) misc[20] = 0;                           // row counter of this group
#pragma unroll
        for (int r = 0; r < ST_TH; ++r) G[gc * ST_GPITCH + r * ST_TW + gxl] = gnext[r];
        __syncthreads();
        if (c0 + 32 < gout.c) load_g(c0 + 32);

        // rows are handed out dynamically, heaviest (middle of the region) first
        for (;;) {
            int k = 0;
            if (lane == 0) k = atomicAdd(&misc[20], 1);
            k = __shfl_sync(0xffffffffu, k, 0);
            if (k >= RW) break;
            const int row = (k & 1) ? (RW / 2) - ((k + 1) >> 1) : (RW / 2) + (k >> 1);
            const int i_beg = off[row * 32], i_end = off[row * 32 + 32];
            if (i_beg == i_end) continue;                     // warp-uniform
            // The row's items arrive sorted by first column, so the NW columns under the current item
            // live in registers; a column is stored to the row buffer once, when the window moves past it.
            // Software pipeline: item it+2 and the operands of item it+1 are in flight while item it is added.
            float* bl = buf + lane * SR_BPITCH;
            float a[NW];
#pragma unroll
            for (int j = 0; j < NW; ++j) a[j] = 0.f;
            int2 item1 = items[i_beg];
            int2 item2 = items[min(i_beg + 1, i_end - 1)];
            int wb = item1.x >> 8;                            // window base column
            float g1 = Gl[item1.x & 255];
            float w1[NW];
            auto load_w = [&](float (&w)[NW], int pix) {
                if (NW == 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wts + pix * NW);
                    w[0] = w4.x; w[1] = w4.y; w[NW - 2] = w4.z; w[NW - 1] = w4.w;
                } else {
                    const float2 w2 = *reinterpret_cast<const float2*>(wts + pix * NW);
                    w[0] = w2.x; w[1] = w2.y;
                }
            };
            load_w(w1, item1.x & 255);
#pragma unroll 2
            for (int it = i_beg; it < i_end; ++it) {
                const int2 item = item1;
                const float gw = g1 * __int_as_float(item.y);
                float w[NW];
#pragma unroll
                for (int j = 0; j < NW; ++j) w[j] = w1[j];
                const int cb = item.x >> 8;
                item1 = item2;
                item2 = items[min(it + 2, i_end - 1)];
                g1 = Gl[item1.x & 255];
                load_w(w1, item1.x & 255);
                const int sh = cb - wb;                       // warp-uniform, >= 0
                if (sh != 0) {
                    float* d = bl + wb;
                    if (sh >= NW) {
#pragma unroll
                        for (int j = 0; j < NW; ++j) { d[j] = a[j]; a[j] = 0.f; }
                    } else if (sh == 1) {
                        d[0] = a[0];
#pragma unroll
                        for (int j = 0; j + 1 < NW; ++j) a[j] = a[j + 1];
                        a[NW - 1] = 0.f;
                    } else if (NW == 4 && sh == 2) {
                        d[0] = a[0]; d[1] = a[1];
                        a[0] = a[NW - 2]; a[1] = a[NW - 1]; a[NW - 2] = 0.f; a[NW - 1] = 0.f;
                    } else if (NW == 4) {                     // sh == 3
                        d[0] = a[0]; d[1] = a[1]; d[NW - 2] = a[NW - 2];
                        a[0] = a[NW - 1]; a[1] = 0.f; a[NW - 2] = 0.f; a[NW - 1] = 0.f;
                    }
                    wb = cb;
                }
#pragma unroll
                for (int j = 0; j < NW; ++j) a[j] = fmaf(gw, w[j], a[j]);
            }
            {
                float* d = bl + wb;
#pragma unroll
                for (int j = 0; j < NW; ++j) d[j] = a[j];
            }
            __syncwarp();
            // one coalesced RED per channel; the buffer is left zeroed for the next row
            const int gy = ry0 + row;
            const bool row_ok = Geo::CLAMP || (unsigned)gy < (unsigned)gsrc.h;
            const int gyc = Geo::CLAMP ? min(max(gy, 0), gsrc.h - 1) : gy;
            if (lane < RW) {
                float* gp = gsrc.p + b * gsrc.sb + (int64_t)c0 * gsrc.sc + gyc * gsrc.sh + gxc * gsrc.sw;
                float* bp = buf + lane;
                const bool ok = row_ok && col_ok;
#pragma unroll 8
                for (int c = 0; c < 32; ++c, gp += gsrc.sc, bp += SR_BPITCH) {
                    const float v = *bp;
                    *bp = 0.f;
                    if (ok && c < nch) sr_red_nonzero(gp, v);
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
