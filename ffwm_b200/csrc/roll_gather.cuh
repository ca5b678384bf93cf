// Rolling-strip channel-lane gathers (second generation of gather_tiled.cuh).
//
// gather_tiled.cuh stages a 31x31 halo region per 16x16 tile: 3.75x read amplification through
// L2, one CTA per SM whose fill -> barrier -> gather phases never overlap, and 16 per-tap offsets
// fetched per pixel.  Here a CTA owns a STRIP of the output grid, 16 columns wide, and walks it
// top to bottom in steps of 8 rows:
//
//   * the source rows the strip can touch live in a RING of 32 rows x 32 columns x 32 channels in
//     shared memory (slab[c][row & 31][col], channel pitch 1025 = 1 mod 32).  A step needs the
//     rows [8s-8, 8s+16); the 8 rows the NEXT step adds are requested with cp.async while this
//     step computes, so the fill latency hides behind the gather and every source row enters
//     shared memory once per strip (amplification 32/16 = 2x instead of 3.75x);
//   * the ring holds an image that is already PADDED the way the op pads — edge replication for
//     resample2d / block_extractor (the reference clamps every tap index), zeros for grid_warp —
//     so a pixel whose window lies inside the ring needs no per-tap clamping or validity test:
//     its taps are `row_offset[i] + j` with compile-time j, i.e. one base register per window row
//     and immediate offsets; the per-pixel record shrinks to 2 packed offset words + weights;
//   * lanes are CHANNELS as before (every tap is one conflict-free shared-memory wavefront for any
//     flow); a warp owns 8 consecutive pixels of one row of the step and forms the geometry of its
//     own pixels for four steps at a time (32 lanes = 32 pixels), in warp-private shared memory —
//     no block-wide geometry phase; the flow values of the next block are prefetched into
//     registers;
//   * one block barrier per step.
// Pixels whose window leaves the ring (|displacement| >= 7 px) take a per-lane global-load path.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace ffwm {

constexpr int RG_SW = 16;                         // strip width (output columns)
constexpr int RG_SH = 8;                          // output rows per step
constexpr int RG_M = 8;                           // margin on every side
constexpr int RG_RW = RG_SW + 2 * RG_M;           // ring row width: 32 columns
constexpr int RG_RING = 32;                       // ring rows (power of two)
constexpr int RG_CHP = RG_RING * RG_RW + 1;       // channel pitch: 1025 = 1 (mod 32)
constexpr int RG_THREADS = 512, RG_WARPS = 16;
constexpr int RG_PXW = 8;                         // pixels per warp and step
constexpr int RG_BLK = 4;                         // steps per geometry block (RG_BLK * RG_PXW = 32 lanes)
constexpr int RG_SPITCH = RG_PXW + 1;             // per-warp staging [32 channels][9]

__device__ __forceinline__ unsigned rg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rg_cp_async4(unsigned dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void rg_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ring rows [vr0, vr0+nrows) <- src[b, c0.., row, rx0..rx0+31] with the op's padding.
// Warp w copies channels 2w and 2w+1 (channels past nch repeat the last real one so that unused
// lanes read finite data); lane j copies column rx0+j: 128 contiguous bytes per row and warp.
template <bool ZERO_OOB>
__device__ __forceinline__ void rg_fill_rows(float* slab, const View<const float>& src, int b, int c0, int nch,
                                             int rx0, int vr0, int nrows, int warp, int lane) {
    const int gx = rx0 + lane;
    const bool col_in = (unsigned)gx < (unsigned)src.w;
    const int gxc = min(max(gx, 0), src.w - 1);
#pragma unroll
    for (int cc = 0; cc < 32 / RG_WARPS; ++cc) {
        const int c = warp * (32 / RG_WARPS) + cc;
        const float* plane = src.p + b * src.sb + (int64_t)(c0 + min(c, nch - 1)) * src.sc + gxc * src.sw;
        const unsigned sp0 = rg_smem_u32(slab) + 4u * (unsigned)(c * RG_CHP + lane);
#pragma unroll 4
        for (int r = vr0; r < vr0 + nrows; ++r) {
            const unsigned sp = sp0 + 4u * (unsigned)((r & (RG_RING - 1)) * RG_RW);
            if (ZERO_OOB) {
                if (col_in && (unsigned)r < (unsigned)src.h) rg_cp_async4(sp, plane + r * src.sh);
                else asm volatile("st.shared.f32 [%0], %1;" ::"r"(sp), "f"(0.f) : "memory");
            } else {
                rg_cp_async4(sp, plane + min(max(r, 0), src.h - 1) * src.sh);
            }
        }
    }
}

// stage[32][9] (this warp's 8 pixels x 32 channels) -> dst[b, c0.., y, xw0..xw0+7]: 32-byte row segments
__device__ __forceinline__ void rg_store_row(const float* stage, const View<float>& dst, int b, int c0, int nch,
                                             int y, int xw0, int lane) {
    const int px = lane & 7, csub = lane >> 3;
    if (xw0 + px >= dst.w) return;
    float* op = dst.p + b * dst.sb + (int64_t)c0 * dst.sc + y * dst.sh + (xw0 + px) * dst.sw;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int c = it * 4 + csub;
        if (c < nch) st_stream(op + (int64_t)c * dst.sc, stage[c * RG_SPITCH + px]);
    }
}

// Applicability of the rolling kernels: fp32, non-negative strides, enough strips to fill the GPU.
inline bool roll_applicable(int n, int c, int h, int w, const View<const float>& src, int ctas_per_strip) {
    if (opt(OPT_DISABLE_ROLL) || opt(OPT_DISABLE_TILED)) return false;
    if (src.sh < 0 || src.sw < 0 || c < 16 || n > 65535 || h < 32) return false;
    if ((int64_t)(src.h - 1) * src.sh + (int64_t)(src.w - 1) * src.sw >= (1 << 30)) return false;
    if (opt(OPT_FORCE_ROLL) || opt(OPT_FORCE_TILED)) return true;   // tests: small shapes through the rolling kernels
    const int64_t ctas = (int64_t)ceil_div(w, RG_SW) * n * ctas_per_strip;
    return ctas >= sm_count() / 2;
}

}  // namespace ffwm
