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
//      are bucketed BY DESTINATION inside the RW x RW halo region around the tile with a counting
//      sort in shared memory (native integer shared atomics for the counts, one block scan, one
//      fill).  Every entry carries its tile pixel, its destination and a "last entry of this
//      destination" flag.  Taps that land outside the region go to a "far" list at the back of
//      the same array.  The destinations are then cut into 16 contiguous ranges with equal
//      entry counts, one per warp.
//   2. for every group of 32 channels: stage grad_output[32][256] in shared memory (coalesced
//      reads) and walk the entry stream: lanes are channels, so
//          acc(c = lane) += w_e * G[lane][p_e]          (conflict-free LDS, broadcast entry)
//      and on a "last" flag the register is stored to the region buffer R[32][RW*RW] — no atomics
//      of any kind, no per-destination loop overhead.  R is flushed row by row with COALESCED
//      red.adds (neighbouring tiles overlap in the halo, so the flush has to accumulate); zeros
//      are skipped.  Far entries (rare for |flow| <= 6 px) are applied with direct REDs.
// Geometry and sort are amortised over all C channels of the tile.
#pragma once
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace ffwm {

constexpr int ST_TW = 16, ST_TH = 16, ST_NPX = ST_TW * ST_TH;     // tile of the output grid
constexpr int ST_THREADS = 512, ST_WARPS = ST_THREADS / 32;
constexpr int ST_GPITCH = ST_NPX + 1;                              // G[c][p], lanes across c: conflict free

struct StEntry {          // 8 bytes
    int p;                // near: pixel | dest << 8 | last << 31 ; far: pixel | (global element offset << 8)
    float w;
};

template <int NT, int RW>
struct StSmem {
    static constexpr int RPX = RW * RW;
    static constexpr size_t bytes() {
        return sizeof(float) * (32 * RPX)               // R: region accumulators [32][RW*RW]
               + sizeof(float) * (32 * ST_GPITCH)       // G: grad_output tile [32][257]
               + sizeof(StEntry) * (ST_NPX * NT)        // entries (near from the front, far from the back)
               + sizeof(int) * (2 * (RPX + 1) + 64);    // cnt, off, misc
    }
};

// Geo::NT, Geo::RW (odd, RW*RW = 1 mod 32: 15, 17, 31)
// Geo::region_origin(tx0,ty0,ml,&rx0,&ry0)  where the tile's taps are expected to land
// Geo::taps(b, y, x, iy[], ix[], w[]) -> clamped/valid destination coords and weights; iy < 0 = skip
template <class Geo>
__global__ void __launch_bounds__(ST_THREADS, 1)
scatter_tiled_kernel(Geo geo, View<const float> gout, View<float> gsrc, int ml) {
    constexpr int NT = Geo::NT, RW = Geo::RW, RPX = RW * RW;
    static_assert(RPX % 32 == 1 && (32 * RPX) % 4 == 0, "region pitch must be 1 mod 32");
    extern __shared__ __align__(16) unsigned char st_smem_raw[];
    float* R = reinterpret_cast<float*>(st_smem_raw);
    float* G = R + 32 * RPX;
    StEntry* ent = reinterpret_cast<StEntry*>(G + 32 * ST_GPITCH);
    int* cnt = reinterpret_cast<int*>(ent + ST_NPX * NT);
    int* off = cnt + (RPX + 1);
    int* misc = off + (RPX + 1);           // [0] = number of far entries, [1..16] = warp totals of the scan

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * ST_TW, ty0 = blockIdx.y * ST_TH, b = blockIdx.z;
    int rx0, ry0;                            // origin of the destination region in the source plane
    geo.region_origin(tx0, ty0, ml, rx0, ry0);

    for (int i = tid; i < RPX + 1; i += ST_THREADS) cnt[i] = 0;
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
                if ((unsigned)ly < (unsigned)RW && (unsigned)lx < (unsigned)RW) {
                    const int d = ly * RW + lx;
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

    // ---- 1b. exclusive scan of cnt[0..RPX] -> off -----------------------------------------
    {
        constexpr int PER = (RPX + 1 + ST_THREADS - 1) / ST_THREADS;     // 2 for RW=31, 1 for RW<=17
        int a[PER], s = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = PER * tid + k;
            a[k] = i <= RPX ? cnt[i] : 0;
            s += a[k];
        }
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
        int ex = misc[1 + warp] + inc - s;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int i = PER * tid + k;
            if (i <= RPX) off[i] = ex;
            ex += a[k];
        }
    }
    __syncthreads();

    // ---- 1c. fill the entry array ---------------------------------------------------------
    if (tid < ST_NPX) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int ds = dslot[t];
            if (ds == INT_MIN) continue;
            StEntry e;
            e.w = wgt[t];
            if (ds >= 0) {
                const int d = ds >> 12, slot = ds & 4095;
                e.p = tid | (d << 8) | (slot == cnt[d] - 1 ? INT_MIN : 0);
                ent[off[d] + slot] = e;
            } else {
                e.p = tid | (goff[t] << 8);
                ent[ST_NPX * NT - 1 - (-1 - ds)] = e;
            }
        }
    }
    __syncthreads();
    const int nfar = misc[0];
    // cut the destinations into ST_WARPS contiguous ranges with (nearly) equal entry counts
    int e_beg, e_end;
    {
        const int total = off[RPX];
        auto first_dest_at = [&](int target) {           // smallest q with off[q] >= target
            int lo = 0, hi = RPX;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (off[mid] < target) lo = mid + 1; else hi = mid;
            }
            return lo;
        };
        const int q_beg = warp == 0 ? 0 : first_dest_at((int)(((long long)total * warp) / ST_WARPS));
        const int q_end = warp == ST_WARPS - 1 ? RPX : first_dest_at((int)(((long long)total * (warp + 1)) / ST_WARPS));
        e_beg = off[q_beg];
        e_end = off[q_end];
    }
    const float* Gl = G + lane * ST_GPITCH;
    float* Rl = R + lane * RPX;
    const int r_lo = max(0, -ry0), r_hi = min(RW, gsrc.h - ry0);       // region rows inside the image
    const int gx = rx0 + lane;
    const bool col_ok = lane < RW && (unsigned)gx < (unsigned)gsrc.w;

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
        // clear the region accumulators (destinations without entries must read as zero)
        for (int i = tid; i < (32 * RPX) / 4; i += ST_THREADS) reinterpret_cast<float4*>(R)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        // this warp's slice of the entry stream
        {
            float acc = 0.f;
            int e = e_beg;
            for (; e + 4 <= e_end; e += 4) {
                StEntry en[4];
                float gv[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) en[k] = ent[e + k];
#pragma unroll
                for (int k = 0; k < 4; ++k) gv[k] = Gl[en[k].p & 255];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc = fmaf(en[k].w, gv[k], acc);
                    if (en[k].p < 0) {                   // warp-uniform: last entry of its destination
                        Rl[(en[k].p >> 8) & 0x3ff] = acc;
                        acc = 0.f;
                    }
                }
            }
            for (; e < e_end; ++e) {
                const StEntry en = ent[e];
                acc = fmaf(en.w, Gl[en.p & 255], acc);
                if (en.p < 0) {
                    Rl[(en.p >> 8) & 0x3ff] = acc;
                    acc = 0.f;
                }
            }
        }
        // far entries: direct REDs, lanes are channels
        for (int k = warp; k < nfar; k += ST_WARPS) {
            const StEntry en = ent[ST_NPX * NT - 1 - k];
            const int p = en.p & 255, go = en.p >> 8;
            if (lane < nch) red_add(gsrc.p + b * gsrc.sb + (int64_t)(c0 + lane) * gsrc.sc + go, en.w * Gl[p]);
        }
        __syncthreads();
        // flush: warp w owns channels 2w, 2w+1; one coalesced RED per region row, zeros skipped
        if (col_ok) {
#pragma unroll
            for (int cc = 0; cc < 32 / ST_WARPS; ++cc) {
                const int c = warp * (32 / ST_WARPS) + cc;
                if (c >= nch) break;
                float* gp = gsrc.p + b * gsrc.sb + (int64_t)(c0 + c) * gsrc.sc + (int64_t)(ry0 + r_lo) * gsrc.sh + gx * gsrc.sw;
                const float* rp = R + c * RPX + r_lo * RW + lane;
#pragma unroll 4
                for (int row = r_lo; row < r_hi; ++row, gp += gsrc.sh, rp += RW) {
                    const float v = *rp;
                    if (v != 0.f) red_add(gp, v);
                }
            }
        }
        __syncthreads();
    }
}

// The far encoding keeps the in-plane element offset in 23 bits.
inline bool scatter_tiled_applicable(const View<const float>& gout, const View<float>& gsrc) {
    if (opt(OPT_DISABLE_TILED)) return false;
    const int64_t span = (int64_t)(gsrc.h - 1) * (gsrc.sh < 0 ? -gsrc.sh : gsrc.sh) + (int64_t)(gsrc.w - 1) * (gsrc.sw < 0 ? -gsrc.sw : gsrc.sw);
    if (gsrc.sh < 0 || gsrc.sw < 0 || span >= (1 << 23)) return false;
    if (gout.c < 16 || gout.n > 65535) return false;
    if (ceil_div(gout.h, ST_TH) > 65535) return false;
    if (opt(OPT_FORCE_TILED)) return true;          // tests: small ragged shapes through the tiled kernels
    const int64_t tiles = (int64_t)ceil_div(gout.w, ST_TW) * ceil_div(gout.h, ST_TH) * gout.n;
    return tiles >= sm_count() / 2;
}

template <class Geo>
static int launch_scatter_tiled(const Geo& geo, const View<const float>& gout, const View<float>& gsrc, int ml, cudaStream_t st) {
    const size_t smem = StSmem<Geo::NT, Geo::RW>::bytes();
    cudaError_t e = cudaFuncSetAttribute(scatter_tiled_kernel<Geo>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("scatter_tiled: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(ceil_div(gout.w, ST_TW), ceil_div(gout.h, ST_TH), gout.n);
    scatter_tiled_kernel<Geo><<<grid, ST_THREADS, smem, st>>>(geo, gout, gsrc, ml);
    return FFWM_OK;
}

}  // namespace ffwm
