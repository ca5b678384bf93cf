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
//     destination row.  The rows of a tile's destination region (31 rows x 32 columns around the
//     16x16 tile) are OWNED by one warp at a time, and the window rows are bucketed by (destination
//     row, first column) once per tile (counting sort of 256*NW items instead of 256*NW*NW entries);
//   * lanes are channels — TWO per lane (c and c+32), so every item's bookkeeping is paid once per 64
//     channels — hence the lanes of a warp never collide, and no other warp touches the row: no
//     atomics and no "last entry" bookkeeping.  The items of a row arrive sorted by first column, so
//     the NW columns under the current item are a sliding window held in REGISTERS: a tap is one
//     FFMA, and a column is stored to the warp's row buffer exactly once, when the window moves past
//     it.  Per item: one broadcast item (pixel, column, row weight), the pixel's column weights (one
//     128-bit broadcast) and the lane's two grad_output values; the item stream is software-pipelined;
//   * a finished row leaves the buffer as COALESCED red.adds — 128-bit vector REDs (REDG.ADD.F32x4,
//     sm_90+) covering whole 32-byte sectors when the layout allows — and the buffer is zeroed on
//     the way out.  The region is in VIRTUAL coordinates: columns/rows outside the image are clamped
//     (resample2d: the reference clamps every tap index) or dropped (grid_warp: zeros padding) when
//     the row is flushed, so windows need no per-tap clamping;
//   * rows are handed out dynamically (heaviest first); grad_output of the NEXT channel group is
//     prefetched into registers while this one is walked.
// Pixels whose window leaves the region (|displacement| > ~6 px) keep the per-tap far list of
// scatter_tiled.cuh (direct REDs, lanes are channels).
// Measured (ks4 grad_input1, cfg5 point): scatter_tiled 1.73 ms; rows in registers behind a jump
// table 1.83; read-modify-write rows in shared memory 1.33; sliding window, 32 channels per group and 16 warps
// 1.23; 64 channels per group: 8 warps 1.23 (latency-bound), 12 warps 1.01, 16 warps (far list cut to fit) 1.01.
#pragma once
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "scatter_tiled.cuh"

namespace ffwm {

constexpr int SR_RH = 31;                       // destination region: 31 rows ...
constexpr int SR_RW = 32;                       // ... x 32 columns (first column = tile - 8: 16-byte aligned)
constexpr int SR_MX = 8, SR_MY = 7;             // margins left / above the tile
constexpr int SR_BPITCH = 33;                   // row buffer [64 channels][33]
constexpr int SR_KEYS = SR_RH * 32;             // sort key = destination row * 32 + first column
constexpr int SR_THREADS = 384, SR_WARPS = 12;     // 256 threads form the geometry and stage grad_output, 12 warps walk the rows
constexpr int SR_CH = 64;                       // channels per group: lane l owns l and l + 32
constexpr int SR_GPITCH = ST_NPX + 1;           // G[c][257]

template <int NW>
struct SrSmem {
    static constexpr size_t bytes() {
        return sizeof(float) * (SR_CH * SR_GPITCH)                  // G: grad_output tile [64][257]
               + sizeof(float) * (SR_WARPS * SR_CH * SR_BPITCH)     // per-warp row buffers
               + sizeof(float) * (ST_NPX * NW)                      // per-pixel column weights
               + sizeof(int2) * (ST_NPX * NW + 4)                   // items {pixel | column << 8, row weight} (+ padding for the prefetch)
               + sizeof(StEntry) * (ST_NPX * NW * NW)               // far taps
               + sizeof(int) * (2 * (SR_KEYS + 1) + 32);            // counts, offsets over the keys, misc
    }
};

// red.add of a non-zero value, predicated instead of branched
__device__ __forceinline__ void sr_red_nonzero(float* p, float v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.neu.f32 q, %1, 0f00000000;\n\t@q red.global.add.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v) : "memory");
}
// 128-bit vector red.add (16-byte aligned address), skipped when all four values are zero
__device__ __forceinline__ void sr_red4_nonzero(float* p, float a, float b, float c, float d) {
    const unsigned any = (__float_as_uint(a) | __float_as_uint(b) | __float_as_uint(c) | __float_as_uint(d)) << 1;   // ignore the sign of zeros
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "r"(any) : "memory");
}

// Geo (see scatter_tiled.cuh) plus:
//   Geo::NW, Geo::CLAMP
//   Geo::window(b, y, x, rx0, ry0, rw, rh, &cb, &rb, wx[NW], wy[NW]) -> true when the pixel's NW x NW window lies
//   inside the rw x rh region at (rx0, ry0); cb/rb = region column/row of the window's first column/row, wx/wy =
//   column / row weights by window position (their product is the tap weight).
template <class Geo>
__global__ void __launch_bounds__(SR_THREADS, 1)
scatter_rows_kernel(Geo geo, View<const float> gout, View<float> gsrc, int vec_ok) {
    constexpr int NW = Geo::NW, NT = NW * NW;
    static_assert(Geo::NT == NT, "window and tap list must agree");
    extern __shared__ __align__(16) unsigned char sr_smem_raw[];
    float* G = reinterpret_cast<float*>(sr_smem_raw);                          // [64][257]
    float* bufs = G + SR_CH * SR_GPITCH;                                       // [12 warps][64][33]
    float* wts = bufs + SR_WARPS * SR_CH * SR_BPITCH;                          // [256][NW] column weights
    int2* items = reinterpret_cast<int2*>(wts + ST_NPX * NW);                  // [256*NW + 4]
    StEntry* far = reinterpret_cast<StEntry*>(items + ST_NPX * NW + 4);        // [256*NT]
    int* cnt = reinterpret_cast<int*>(far + ST_NPX * NT);                      // [SR_KEYS + 1]
    int* off = cnt + (SR_KEYS + 1);                                            // [SR_KEYS + 1]
    int* misc = off + (SR_KEYS + 1);                                           // [0] far taps, [1..12] scan partials, [20] row counter

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * ST_TW, ty0 = blockIdx.y * ST_TH, b = blockIdx.z;
    const int rx0 = tx0 - SR_MX, ry0 = ty0 - SR_MY;

    for (int i = tid; i < SR_KEYS + 1; i += SR_THREADS) cnt[i] = 0;
    if (tid == 0) misc[0] = 0;
    if (tid < 4) items[ST_NPX * NW + tid] = make_int2(0, 0);
    for (int i = tid; i < SR_WARPS * SR_CH * SR_BPITCH; i += SR_THREADS) bufs[i] = 0.f;
    __syncthreads();

    // ---- 1. geometry, once per tile (one pixel per thread): window rows bucketed by (row, column) ----
    int my_cb = 0, my_rb = 0, my_slot[NW];
    float my_wy[NW];
    bool near = false;
    if (tid < ST_NPX) {
        const int y = ty0 + tid / ST_TW, x = tx0 + tid % ST_TW;
        if (y < gout.h && x < gout.w) {
            float wx[NW], wy[NW];
            near = geo.window(b, y, x, rx0, ry0, SR_RW, SR_RH, my_cb, my_rb, wx, wy);
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
    // exclusive scan of cnt[0..SR_KEYS] -> off (four keys per thread, warp scan, scan of the warp totals)
    {
        constexpr int PER = (SR_KEYS + 1 + SR_THREADS - 1) / SR_THREADS;
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
            const int v = lane < SR_WARPS ? misc[1 + lane] : 0;
            int winc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += u;
            }
            if (lane < SR_WARPS) misc[1 + lane] = winc - v;
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

    // grad_output staging: thread -> (channel tid/16 + 16k, column tid%16), 16 rows; next group in registers
    const int gc = tid / ST_TW, gxl = tid % ST_TW;
    float gnext[SR_CH / 16][ST_TH];
    auto load_g = [&](int c0) {
        if (tid >= ST_NPX) return;
        const int nch = min(SR_CH, gout.c - c0);
        const bool colok = tx0 + gxl < gout.w;
#pragma unroll
        for (int k = 0; k < SR_CH / 16; ++k) {
            const int c = gc + 16 * k;
            const float* gp = gout.p + b * gout.sb + (int64_t)(c0 + c) * gout.sc + (tx0 + gxl) * gout.sw;
            const bool ok = colok && c < nch;
#pragma unroll
            for (int r = 0; r < ST_TH; ++r) {
                const int y = ty0 + r;
                gnext[k][r] = (ok && y < gout.h) ? ld_stream(gp + y * gout.sh) : 0.f;
            }
        }
    };
    load_g(0);

    float* buf = bufs + warp * (SR_CH * SR_BPITCH);
    float* bl = buf + lane * SR_BPITCH;                       // lane's first channel; the second one is 32 rows further
    const float* Gl = G + lane * SR_GPITCH;
    constexpr int G2 = 32 * SR_GPITCH, B2 = 32 * SR_BPITCH;   // offsets of the lane's second channel

    // ---- 2. channel groups -------------------------------------------------------------------
    for (int c0 = 0; c0 < gout.c; c0 += SR_CH) {
        const int nch = min(SR_CH, gout.c - c0);
        __syncthreads();                                      // items written / previous group's G consumed
        if (tid == 0) misc[20] = 0;                           // row counter of this group
        if (tid < ST_NPX) {
#pragma unroll
            for (int k = 0; k < SR_CH / 16; ++k)
#pragma unroll
                for (int r = 0; r < ST_TH; ++r) G[(gc + 16 * k) * SR_GPITCH + r * ST_TW + gxl] = gnext[k][r];
        }
        __syncthreads();
        if (c0 + SR_CH < gout.c) load_g(c0 + SR_CH);

        // rows are handed out dynamically, heaviest (middle of the region) first
        for (;;) {
            int k = 0;
            if (lane == 0) k = atomicAdd(&misc[20], 1);
            k = __shfl_sync(0xffffffffu, k, 0);
            if (k >= SR_RH) break;
            const int row = (k & 1) ? (SR_RH / 2) - ((k + 1) >> 1) : (SR_RH / 2) + (k >> 1);
            const int i_beg = off[row * 32], i_end = off[row * 32 + 32];
            if (i_beg == i_end) continue;                     // warp-uniform
            // sliding window of NW columns x 2 channels in registers; items two at a time (A, B), the
            // operands of the next one in flight while the current one is added
            float a0[NW], a1[NW];
#pragma unroll
            for (int j = 0; j < NW; ++j) a0[j] = a1[j] = 0.f;
            int wb = items[i_beg].x >> 8;                     // window base column
            auto advance = [&](int cb) {                      // move the window to column cb >= wb
                const int sh = cb - wb;                       // warp-uniform
                if (sh == 0) return;
                float* d = bl + wb;
                if (sh >= NW) {
#pragma unroll
                    for (int j = 0; j < NW; ++j) { d[j] = a0[j]; d[B2 + j] = a1[j]; a0[j] = 0.f; a1[j] = 0.f; }
                } else if (sh == 1) {
                    d[0] = a0[0]; d[B2] = a1[0];
#pragma unroll
                    for (int j = 0; j + 1 < NW; ++j) { a0[j] = a0[j + 1]; a1[j] = a1[j + 1]; }
                    a0[NW - 1] = 0.f; a1[NW - 1] = 0.f;
                } else if (NW == 4 && sh == 2) {
                    d[0] = a0[0]; d[1] = a0[1]; d[B2] = a1[0]; d[B2 + 1] = a1[1];
                    a0[0] = a0[NW - 2]; a0[1] = a0[NW - 1]; a0[NW - 2] = 0.f; a0[NW - 1] = 0.f;
                    a1[0] = a1[NW - 2]; a1[1] = a1[NW - 1]; a1[NW - 2] = 0.f; a1[NW - 1] = 0.f;
                } else if (NW == 4) {                         // sh == 3
                    d[0] = a0[0]; d[1] = a0[1]; d[NW - 2] = a0[NW - 2];
                    d[B2] = a1[0]; d[B2 + 1] = a1[1]; d[B2 + NW - 2] = a1[NW - 2];
                    a0[0] = a0[NW - 1]; a0[1] = 0.f; a0[NW - 2] = 0.f; a0[NW - 1] = 0.f;
                    a1[0] = a1[NW - 1]; a1[1] = 0.f; a1[NW - 2] = 0.f; a1[NW - 1] = 0.f;
                }
                wb = cb;
            };
            struct Opd { float g0, g1, w[NW]; };
            auto fetch = [&](Opd& o, int2 item) {
                const int pix = item.x & 255;
                o.g0 = Gl[pix];
                o.g1 = Gl[G2 + pix];
                if (NW == 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wts + pix * NW);
                    o.w[0] = w4.x; o.w[1] = w4.y; o.w[NW - 2] = w4.z; o.w[NW - 1] = w4.w;
                } else {
                    const float2 w2 = *reinterpret_cast<const float2*>(wts + pix * NW);
                    o.w[0] = w2.x; o.w[1] = w2.y;
                }
            };
            auto add = [&](const Opd& o, int2 item) {
                advance(item.x >> 8);
                const float wy = __int_as_float(item.y);
                const float gw0 = o.g0 * wy, gw1 = o.g1 * wy;
#pragma unroll
                for (int j = 0; j < NW; ++j) { a0[j] = fmaf(gw0, o.w[j], a0[j]); a1[j] = fmaf(gw1, o.w[j], a1[j]); }
            };
            int2 iA = items[i_beg], iB = items[i_beg + 1];    // reads past i_end stay inside the padded array
            Opd oA, oB;
            fetch(oA, iA);
            for (int it = i_beg; it < i_end; it += 2) {
                fetch(oB, iB);
                const int2 iA2 = items[it + 2], iB2 = items[it + 3];
                add(oA, iA);
                if (it + 1 < i_end) {
                    fetch(oA, iA2);
                    add(oB, iB);
                }
                iA = iA2;
                iB = iB2;
            }
            {
                float* d = bl + wb;
#pragma unroll
                for (int j = 0; j < NW; ++j) { d[j] = a0[j]; d[B2 + j] = a1[j]; }
            }
            __syncwarp();
            // flush: coalesced REDs, the buffer is left zeroed for the next row
            const int gy = ry0 + row;
            const bool row_ok = Geo::CLAMP || (unsigned)gy < (unsigned)gsrc.h;
            const int gyc = Geo::CLAMP ? min(max(gy, 0), gsrc.h - 1) : gy;
            float* grow = gsrc.p + b * gsrc.sb + (int64_t)c0 * gsrc.sc + gyc * gsrc.sh;
            if (vec_ok) {
                // lane -> (channel lane/8 of a group of four, 16-byte chunk lane%8): whole 32-byte sectors.
                // rx0 and the width are multiples of 4, so a chunk is entirely inside or outside the image.
                const int cq = lane >> 3, gx4 = rx0 + 4 * (lane & 7);
                const bool inside = (unsigned)gx4 < (unsigned)gsrc.w;
                const int gxb = gx4 < 0 ? 0 : gsrc.w - 1;    // border column the outside chunks fold onto (CLAMP)
                float* bp = buf + cq * SR_BPITCH + 4 * (lane & 7);
                float* gp = grow + (int64_t)cq * gsrc.sc;
#pragma unroll 4
                for (int c = cq; c < SR_CH; c += 4, bp += 4 * SR_BPITCH, gp += 4 * gsrc.sc) {
                    const float v0 = bp[0], v1 = bp[1], v2 = bp[2], v3 = bp[3];
                    bp[0] = 0.f; bp[1] = 0.f; bp[2] = 0.f; bp[3] = 0.f;
                    if (c < nch && row_ok) {
                        if (inside) sr_red4_nonzero(gp + gx4, v0, v1, v2, v3);
                        else if (Geo::CLAMP) sr_red_nonzero(gp + gxb, (v0 + v1) + (v2 + v3));
                    }
                }
            } else {
                const int gx = rx0 + lane;                    // lane = region column
                const int gxc = Geo::CLAMP ? min(max(gx, 0), gsrc.w - 1) : gx;
                const bool ok = row_ok && (Geo::CLAMP || (unsigned)gx < (unsigned)gsrc.w);
                float* gp = grow + gxc * gsrc.sw;
                float* bp = buf + lane;
#pragma unroll 8
                for (int c = 0; c < SR_CH; ++c, gp += gsrc.sc, bp += SR_BPITCH) {
                    const float v = *bp;
                    *bp = 0.f;
                    if (ok && c < nch) sr_red_nonzero(gp, v);
                }
            }
            __syncwarp();
        }
        // far taps: direct REDs, lanes are channels
        for (int k = warp; k < nfar; k += SR_WARPS) {
            const StEntry en = far[k];
            const int p = en.p & 255, go = en.p >> 8;
            float* gp = gsrc.p + b * gsrc.sb + (int64_t)(c0 + lane) * gsrc.sc + go;
            if (lane < nch) red_add(gp, en.w * Gl[p]);
            if (lane + 32 < nch) red_add(gp + 32 * gsrc.sc, en.w * Gl[G2 + p]);
        }
    }
}

template <class Geo>
static int launch_scatter_rows(const Geo& geo, const View<const float>& gout, const View<float>& gsrc, cudaStream_t st) {
    const size_t smem = SrSmem<Geo::NW>::bytes();
    cudaError_t e = cudaFuncSetAttribute(scatter_rows_kernel<Geo>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("scatter_rows: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    const int vec_ok = gsrc.sw == 1 && (gsrc.w & 3) == 0 && (gsrc.sh & 3) == 0 && (gsrc.sc & 3) == 0 && (gsrc.sb & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(gsrc.p) & 15) == 0 && !opt(OPT_SCATTER_SCALAR_FLUSH);
    dim3 grid(ceil_div(gout.w, ST_TW), ceil_div(gout.h, ST_TH), gout.n);
    scatter_rows_kernel<Geo><<<grid, SR_THREADS, smem, st>>>(geo, gout, gsrc, vec_ok);
    return FFWM_OK;
}

}  // namespace ffwm
