// Shared pieces of the tiled GATHER kernels (forward passes and the flow-gradient passes).
//
// Direct gathers (one thread per pixel, taps through L1) are L1-wavefront-bound for incoherent
// flows: neighbouring lanes hit different 128-byte lines, 6-9 wavefronts per load (ncu: l1tex
// 85-90 % at 12-30 % of HBM).  The tiled kernels turn the access around:
//
//   * a CTA owns a 16x16 tile of the output grid and stages the 31x31 halo region of the source,
//     32 channels at a time, in shared memory as slab[c][31*31] — 961 = 1 (mod 32), so
//   * a warp processes ONE pixel at a time with its 32 lanes = 32 CHANNELS: every tap is one
//     conflict-free shared-memory wavefront whatever the flow looks like, the per-pixel geometry
//     (formed once per tile by one thread per pixel and kept in shared memory) is read by
//     broadcast, and nothing diverges because all lanes share the pixel;
//   * results are transposed back through a small per-warp buffer so the NCHW stores are 64-byte
//     row segments.
// Taps outside the halo (|displacement| > ~6 px) fall back to a global load per lane.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace ffwm {

constexpr int GT_TW = 16, GT_TH = 16, GT_NPX = 256;
constexpr int GT_RW = 31, GT_RPX = GT_RW * GT_RW;
constexpr int GT_THREADS = 512, GT_WARPS = 16;
constexpr int GT_SPITCH = 17;                       // per-warp output staging [32][17]
constexpr int GT_GPITCH = GT_NPX + 1;

// Offset of a (clamped, valid) source coordinate as the kernels store it: >= 0 -> index into
// one channel's 31x31 slab; < 0 -> -(1 + element offset inside one global plane).
__device__ __forceinline__ int gt_tap_offset(int iy, int ix, int ry0, int rx0, int sh, int sw) {
    const int ly = iy - ry0, lx = ix - rx0;
    if ((unsigned)ly < (unsigned)GT_RW && (unsigned)lx < (unsigned)GT_RW) return ly * GT_RW + lx;
    return -(1 + iy * sh + ix * sw);
}

__device__ __forceinline__ float gt_load(const float* slab_lane, const float* plane_lane, int off) {
    return off >= 0 ? slab_lane[off] : __ldg(plane_lane - off - 1);
}

// slab[c][31*31] <- src[b, c0+c, ry0.., rx0..] for the part of the region inside the image.
// Warp w copies channels 2w and 2w+1: per region row, 31 lanes issue one 4-byte cp.async (LDGSTS)
// covering 124 contiguous bytes; all row copies of a warp are in flight at once and nothing is
// staged in registers.  The loop only advances one global and one shared pointer.
// Complete with gt_fill_wait() + __syncthreads().
__device__ __forceinline__ void gt_fill_slab(float* slab, const View<const float>& src, int b, int c0, int nch,
                                             int ry0, int rx0, int warp, int lane) {
    const int gx = rx0 + lane;
    if (lane >= GT_RW || (unsigned)gx >= (unsigned)src.w) return;
    const int r_lo = max(0, -ry0), r_hi = min(GT_RW, src.h - ry0);      // region rows inside the image
#pragma unroll
    for (int cc = 0; cc < 32 / GT_WARPS; ++cc) {
        const int c = warp * (32 / GT_WARPS) + cc;
        if (c >= nch) break;
        const float* gp = src.p + b * src.sb + (int64_t)(c0 + c) * src.sc + (int64_t)(ry0 + r_lo) * src.sh + gx * src.sw;
        unsigned sp = (unsigned)__cvta_generic_to_shared(slab) + 4u * (c * GT_RPX + r_lo * GT_RW + lane);
#pragma unroll 4
        for (int row = r_lo; row < r_hi; ++row, gp += src.sh, sp += 4u * GT_RW)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sp), "l"(gp) : "memory");
    }
}

__device__ __forceinline__ void gt_fill_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// G[c][256] <- t[b, c0+c, tile]; thread -> (channel tid/16, column tid%16), 16 rows.
__device__ __forceinline__ void gt_fill_tile(float* G, const View<const float>& t, int b, int c0, int nch,
                                             int ty0, int tx0, int tid) {
    const int c = tid / GT_TW, x = tid % GT_TW;
    const bool ok = c < nch && tx0 + x < t.w;
    const float* gp = t.p + b * t.sb + (int64_t)(c0 + c) * t.sc + (tx0 + x) * t.sw;
#pragma unroll 4
    for (int r = 0; r < GT_TH; ++r) {
        const int y = ty0 + r;
        G[c * GT_GPITCH + r * GT_TW + x] = (ok && y < t.h) ? ld_stream(gp + y * t.sh) : 0.f;
    }
}

// stage[32][17] (this warp's 16 pixels x 32 channels) -> out[b, c0.., y, tx0..tx0+15]
__device__ __forceinline__ void gt_store_row(const float* stage, const View<float>& out, int b, int c0, int nch,
                                             int y, int tx0, int lane) {
    const int x = lane & 15, half = lane >> 4;
    if (y >= out.h || tx0 + x >= out.w) return;
    float* op = out.p + b * out.sb + (int64_t)c0 * out.sc + y * out.sh + (tx0 + x) * out.sw;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
        const int c = it * 2 + half;
        if (c < nch) st_stream(op + (int64_t)c * out.sc, stage[c * GT_SPITCH + x]);
    }
}

// Sum N per-lane values over the 32 lanes of a warp with N + log-ish shuffles instead of 5*N:
// every butterfly step halves the number of live values (each lane keeps the half selected by
// one bit of its lane id).  On return v[0] of lane l holds the warp total of value index l >> (5 - log2 N)...
// for N = 16: index l >> 1; for N = 8: index l >> 2.  Lanes sharing an index hold the same total.
template <int N>
__device__ __forceinline__ float gt_packed_reduce(float (&v)[N], int lane) {
    static_assert(N == 8 || N == 16, "N must be 8 or 16");
    int offset = 16;
#pragma unroll
    for (int n = N / 2; n >= 1; n >>= 1, offset >>= 1) {
        const bool hi = (lane & offset) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = hi ? v[i] : v[i + n];
            const float keep = hi ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, offset);
        }
    }
    float r = v[0];
#pragma unroll
    for (; offset >= 1; offset >>= 1) r += __shfl_xor_sync(0xffffffffu, r, offset);
    return r;
}

inline bool gather_tiled_applicable(int n, int c, int h, int w, const View<const float>& src) {
    if (opt(OPT_DISABLE_TILED)) return false;
    if (src.sh < 0 || src.sw < 0 || c < 16 || n > 65535) return false;
    const int64_t tiles = (int64_t)ceil_div(w, GT_TW) * ceil_div(h, GT_TH) * n;
    return tiles >= sm_count() / 2 && ceil_div(h, GT_TH) <= 65535;
}

}  // namespace ffwm
