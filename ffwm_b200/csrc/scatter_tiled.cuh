// Tiled scatter-add for the warp backward passes (grad wrt the sampled tensor).
//
// Problem: every output-grid pixel p scatters NT weighted copies of grad_output[c,p] to NT
// positions of grad_source[c,.] — the SAME positions and weights for all C channels.  Done
// directly (the reference, and the *_bwd_kernel fallbacks here) that is NT global RED.ADDs per
// element; the L2 RED rate for scattered addresses (~220 G lane-REDs/s measured on B200) then
// bounds the pass at 2-10 % of the HBM roofline.
//
// Plan (one CTA = one 16x16 tile of the output grid, 512 threads, one CTA per SM):
//   1. geometry, once per tile: one thread per pixel forms its taps (policy `Geo`), and the taps
//      are bucketed BY DESTINATION inside the 31x31 halo region around the tile with a counting
//      sort in shared memory (native integer shared atomics for the counts, one block scan, one
//      fill).  Taps that land outside the region go to a "far" list at the back of the same array.
//   2. for every group of 32 channels: stage grad_output[32][256] in shared memory (coalesced
//      reads), then each warp OWNS a set of destinations and accumulates
//          acc(c = lane) = sum over the destination's entries  w_e * G[lane][p_e]
//      in a register: lanes are channels, so the tile reads are conflict-free and the entry
//      reads are broadcasts; no atomics of any kind.  Results are staged in a [32][961] region
//      buffer (conflict-free: 961 = 1 mod 32) and flushed row by row with COALESCED red.adds
//      (neighbouring tiles overlap in the halo, so the flush still has to accumulate) — 3.75
//      coalesced REDs per element instead of NT scattered ones, and zeros are skipped.
//      Far entries (rare for |flow| <= 6 px) are applied with direct REDs, one warp per entry.
// Geometry and sort are amortised over all C channels of the tile.
#pragma once
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace ffwm {

constexpr int ST_TW = 16, ST_TH = 16, ST_NPX = ST_TW * ST_TH;     // tile of the output grid
constexpr int ST_RW = 31, ST_RPX = ST_RW * ST_RW;                  // halo region of the source plane
constexpr int ST_THREADS = 512, ST_WARPS = ST_THREADS / 32;
constexpr int ST_GPITCH = ST_NPX + 1;                              // G[c][p], lanes across c: conflict free

struct StEntry {          // 8 bytes
    int p;                // near: tile-local pixel index; far: pixel index | (global element offset << 8)
    float w;
};

template <int NT>
struct StSmem {
    static constexpr int kEntries = ST_NPX * NT;
    static constexpr size_t bytes() {
        return sizeof(float) * (32 * ST_RPX)            // R: region accumulators [32][961]
               + sizeof(float) * (32 * ST_GPITCH)       // G: grad_output tile [32][257]
               + sizeof(StEntry) * kEntries             // entries (near from the front, far from the back)
               + sizeof(int) * (2 * (ST_RPX + 1) + 64); // cnt, off, misc
    }
};

// Geo::region_origin(tx0,ty0,ml,&rx0,&ry0)  where the tile's taps are expected to land
// Geo::NT                       taps per pixel
// Geo::taps(b, y, x, iy[], ix[], w[]) -> fills clamped/valid destination coords and weights;
//                                        w == 0 entries with iy < 0 are skipped (invalid taps)
template <class Geo>
__global__ void __launch_bounds__(ST_THREADS, 1)
scatter_tiled_kernel(Geo geo, View<const float> gout, View<float> gsrc, int ml) {
    constexpr int NT = Geo::NT;
    extern __shared__ __align__(16) unsigned char st_smem_raw[];
    float* R = reinterpret_cast<float*>(st_smem_raw);
    float* G = R + 32 * ST_RPX;
    StEntry* ent = reinterpret_cast<StEntry*>(G + 32 * ST_GPITCH);
    int* cnt = reinterpret_cast<int*>(ent + ST_NPX * NT);
    int* off = cnt + (ST_RPX + 1);
    int* misc = off + (ST_RPX + 1);        // [0] = number of far entries, [1..16] = warp totals of the scan

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * ST_TW, ty0 = blockIdx.y * ST_TH, b = blockIdx.z;
    int rx0, ry0;                            // origin of the 31x31 destination region in the source plane
    geo.region_origin(tx0, ty0, ml, rx0, ry0);

    for (int i = tid; i < ST_RPX + 1; i += ST_THREADS) cnt[i] = 0;
    if (tid == 0) misc[0] = 0;
    __syncthreads();

    // ---- 1a. geometry + counting (one thread per tile pixel) -----------------------------
    int dslot[NT];       // near: (dest << 12) | slot ; far: -1 - k ; skipped: INT_MIN
    float wgt[NT];
    int goff[NT];        // far only: global element offset inside one plane of gsrc
    if (tid < ST_NPX) {
        const int y = ty0 + tid / ST_TW, x = tx0 + tid % ST_TW;
        if (y < gout.h && x < gout.w) {
            int iy[NT], ix[NT];
            geo.taps(b, y, x, iy, ix, wgt);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                if (iy[t] < 0) { dslot[t] = INT_MIN; continue; }
                const int ly = iy[t] - ry0, lx = ix[t] - rx0;
                if ((unsigned)ly < (unsigned)ST_RW && (unsigned)lx < (unsigned)ST_RW) {
                    const int d = ly * ST_RW + lx;
                    dslot[t] = (d << 12) | atomicAdd(&cnt[d], 1);
                } else {
                    dslot[t] = -1 - atomicAdd(&misc[0], 1);
                    goff[t] = iy[t] * gsrc.sh + ix[t] * gsrc.sw;
                }
            }
        } else {
#pragma unroll
            for (int t = 0; t < NT; ++t) dslot[t] = INT_MIN;
        }
    }
    __syncthreads();

    // ---- 1b. exclusive scan of cnt[0..961] -> off ----------------------------------------
    {
        const int i0 = 2 * tid, i1 = 2 * tid + 1;
        const int a = i0 <= ST_RPX ? cnt[i0] : 0, c2 = i1 <= ST_RPX ? cnt[i1] : 0;
        const int s = a + c2;
        int inc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        if (lane == 31) misc[1 + warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int v = lane < ST_WARPS ? misc[1 + lane] : 0, winc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, winc, d);
                if (lane >= d) winc += u;
            }
            if (lane < ST_WARPS) misc[1 + lane] = winc - v;      // exclusive warp prefix
        }
        __syncthreads();
        const int ex = misc[1 + warp] + inc - s;
        if (i0 <= ST_RPX) off[i0] = ex;
        if (i1 <= ST_RPX) off[i1] = ex + a;
    }
    __syncthreads();

    // ---- 1c. fill the entry array ---------------------------------------------------------
    if (tid < ST_NPX) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int ds = dslot[t];
            if (ds == INT_MIN) continue;
            if (ds >= 0) {
                StEntry e; e.p = tid; e.w = wgt[t];
                ent[off[ds >> 12] + (ds & 4095)] = e;
            } else {
                StEntry e; e.p = tid | (goff[t] << 8); e.w = wgt[t];
                ent[ST_NPX * NT - 1 - (-1 - ds)] = e;
            }
        }
    }
    __syncthreads();
    const int nfar = misc[0];
    // split the destinations into ST_WARPS contiguous ranges with (nearly) equal entry counts
    int q_beg, q_end, e_beg, e_end;
    {
        const int total = off[ST_RPX];
        auto first_dest_at = [&](int target) {           // smallest q with off[q] >= target
            int lo = 0, hi = ST_RPX;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (off[mid] < target) lo = mid + 1; else hi = mid;
            }
            return lo;
        };
        q_beg = warp == 0 ? 0 : first_dest_at((int)(((long long)total * warp) / ST_WARPS));
        q_end = warp == ST_WARPS - 1 ? ST_RPX : first_dest_at((int)(((long long)total * (warp + 1)) / ST_WARPS));
        e_beg = off[q_beg];
        e_end = off[q_end];
    }

    // ---- 2. channel groups -------------------------------------------------------------------
    for (int c0 = 0; c0 < gout.c; c0 += 32) {
        const int nch = min(32, gout.c - c0);
        // stage grad_output[c0..c0+31][tile]: thread -> (channel tid/16, column tid%16), 16 rows
        {
            const int c = tid / ST_TW, x = tid % ST_TW;
            const bool ok = c < nch && tx0 + x < gout.w;
            const float* gp = gout.p + b * gout.sb + (int64_t)(c0 + c) * gout.sc + (tx0 + x) * gout.sw;
#pragma unroll 4
            for (int r = 0; r < ST_TH; ++r) {
                const int y = ty0 + r;
                G[c * ST_GPITCH + r * ST_TW + x] = (ok && y < gout.h) ? ld_stream(gp + y * gout.sh) : 0.f;
            }
        }
        __syncthreads();
        // destinations owned by this warp: a contiguous range, so its entries are one contiguous
        // stream that can be read ahead (4 entries in flight) instead of one dependent
        // entry -> tile-value -> fma chain per entry
        const float* Gl = G + lane * ST_GPITCH;
        float* Rl = R + lane * ST_RPX;
        {
            int q = q_beg, e = e_beg;
            int nb = off[q + 1];
            float acc = 0.f;
            for (; e < e_end; e += 4) {
                StEntry en[4];
                float gv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    en[k].p = 0; en[k].w = 0.f;
                    if (e + k < e_end) en[k] = ent[e + k];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) gv[k] = Gl[en[k].p];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (e + k < e_end) {
                        while (e + k >= nb) {            // warp-uniform: close finished (or empty) destinations
                            Rl[q] = acc;
                            acc = 0.f;
                            ++q;
                            nb = off[q + 1];
                        }
                        acc = fmaf(en[k].w, gv[k], acc);
                    }
                }
            }
            for (; q < q_end; ++q) {                     // the open destination, then trailing empty ones
                Rl[q] = acc;
                acc = 0.f;
            }
        }
        // far entries: direct REDs, lanes are channels
        for (int k = warp; k < nfar; k += ST_WARPS) {
            const StEntry en = ent[ST_NPX * NT - 1 - k];
            const int p = en.p & 255, go = en.p >> 8;
            if (lane < nch) red_add(gsrc.p + b * gsrc.sb + (int64_t)(c0 + lane) * gsrc.sc + go, en.w * Gl[p]);
        }
        __syncthreads();
        // flush the region: one warp per (channel, region row), coalesced along x
        for (int pr = warp; pr < 32 * ST_RW; pr += ST_WARPS) {
            const int c = pr / ST_RW, row = pr - c * ST_RW;
            const int gy = ry0 + row, gx = rx0 + lane;
            if (c < nch && lane < ST_RW && (unsigned)gy < (unsigned)gsrc.h && (unsigned)gx < (unsigned)gsrc.w) {
                const float v = R[c * ST_RPX + row * ST_RW + lane];
                if (v != 0.f) red_add(gsrc.p + b * gsrc.sb + (int64_t)(c0 + c) * gsrc.sc + gy * gsrc.sh + gx * gsrc.sw, v);
            }
        }
        __syncthreads();
    }
}

// The far encoding keeps the in-plane element offset in 23 bits.
inline bool scatter_tiled_applicable(const View<const float>& gout, const View<float>& gsrc) {
    if (getenv("FFWM_DISABLE_TILED")) return false;
    const int64_t span = (int64_t)(gsrc.h - 1) * (gsrc.sh < 0 ? -gsrc.sh : gsrc.sh) + (int64_t)(gsrc.w - 1) * (gsrc.sw < 0 ? -gsrc.sw : gsrc.sw);
    if (gsrc.sh < 0 || gsrc.sw < 0 || span >= (1 << 23)) return false;
    if (gout.c < 16 || gout.n > 65535) return false;
    const int64_t tiles = (int64_t)ceil_div(gout.w, ST_TW) * ceil_div(gout.h, ST_TH) * gout.n;
    return tiles >= sm_count() / 2 && ceil_div(gout.h, ST_TH) <= 65535;
}

template <class Geo>
static int launch_scatter_tiled(const Geo& geo, const View<const float>& gout, const View<float>& gsrc, int ml, cudaStream_t st) {
    const size_t smem = StSmem<Geo::NT>::bytes();
    cudaError_t e = cudaFuncSetAttribute(scatter_tiled_kernel<Geo>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("scatter_tiled: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(ceil_div(gout.w, ST_TW), ceil_div(gout.h, ST_TH), gout.n);
    scatter_tiled_kernel<Geo><<<grid, ST_THREADS, smem, st>>>(geo, gout, gsrc, ml);
    return FFWM_OK;
}

}  // namespace ffwm
