// corr_max: the correlation column-max of PerceptualCorrectness as a fused tcgen05 GEMM + running max.
//
// The reference (models/losses.py:341-353) cosine-normalises the VGG features of source and target per pixel, forms
//     correction = bmm(source_norm [b, N2, C], target_norm [b, C, N2])        [b, N2, N2]
// and keeps only max over the source axis: at relu1_1 (C = 64, N2 = 16384) that is a 1.07 GB intermediate and 34 GFLOP
// per sample written and re-read for one number per target pixel.  Here the product never leaves the SM:
//   * a prepass normalises both feature maps (x / (||x|| + eps), the reference's formula, fp32) and writes them as the
//     two bf16 parts of the 3xBF16 operand split (conv3x3_tc.cu), already in the shared-memory image of the tiles;
//   * one CTA owns 128 TARGET pixels (MMA M = TMEM lanes) and walks over all source pixels in tiles of NI = 256 (128
//     for C = 256) columns: the source tiles arrive by bulk async copies into a 2-3 stage ring, 12 x C/64 MMAs
//     (b1*b1 + b1*b2 + b2*b1, K = 16 channels each) fill one of two TMEM accumulators while the four epilogue warps
//     read the other one back (tcgen05.ld, 32 columns at a time) and fold it into a per-lane running maximum —
//     the max over the source axis is a max over COLUMNS, i.e. inside a thread, no cross-lane reduction;
//   * result: cmax[b, j] for the CTA's 128 target pixels, written once.
// No gradient is defined: the reference's inputs to this product are VGG features of the input images (no grad).
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace ffwm {

struct CmGeo {
    int b, c, n;           // batch, channels (multiple of 64, <= 256), pixels
    int ni;                // source pixels per tile (MMA N)
    int jt, nt, kc;        // target tiles (128 px), source tiles (ni px), 64-channel chunks
    int a_bytes, s_bytes;  // target tile image, one source stage
    int ns;                // ring depth
    float eps;
};

static bool cm_setup(CmGeo& g, int b, int c, int n, float eps) {
    if (b <= 0 || n <= 0 || c <= 0 || c % 64 != 0 || c > 256) return false;
    g.b = b, g.c = c, g.n = n, g.eps = eps;
    g.ni = c <= 128 ? 256 : 128;
    g.jt = (n + 127) / 128, g.nt = (n + g.ni - 1) / g.ni, g.kc = c / 64;
    g.a_bytes = 2 * (c / 8) * 2048;                    // [part][kg][128][16 B]
    g.s_bytes = 2 * 8 * g.ni * 16;                     // [part][kg 0..7][ni][16 B]
    g.ns = std::min(3, (227 * 1024 - 512 - g.a_bytes) / g.s_bytes);
    return g.ns >= 2;
}

// ---------------------------------------------------------------- prepass: normalise + split + tile images
// blockIdx.z = 0: target -> A images, 1: source -> B images.  One thread per (padded) pixel.
__global__ void corr_max_prep_kernel(const float* __restrict__ src, const float* __restrict__ tgt, int64_t s_b, int64_t s_c,
                                     unsigned char* __restrict__ img_a, unsigned char* __restrict__ img_b, CmGeo g) {
    const bool is_src = blockIdx.z == 1;
    const int tile_px = is_src ? g.ni : 128;
    const int npad = (is_src ? g.nt : g.jt) * tile_px;
    const int p = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (p >= npad) return;
    const float* x = (is_src ? src : tgt) + b * s_b + p;
    const bool valid = p < g.n;
    float ss = 0.f;
    if (valid)
        for (int ch = 0; ch < g.c; ++ch) { const float v = __ldg(x + ch * s_c); ss += v * v; }
    const float scale = 1.f / (sqrtf(ss) + g.eps);
    const int tile = p / tile_px, r = p - tile * tile_px;
    for (int kg = 0; kg < g.c / 8; ++kg) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = valid ? __ldg(x + (kg * 8 + j) * s_c) * scale : 0.f;
        unsigned char* d;
        int part;
        if (is_src) {
            part = 8 * g.ni * 16;
            d = img_b + (((int64_t)b * g.nt + tile) * g.kc + kg / 8) * g.s_bytes + (kg % 8) * g.ni * 16 + r * 16;
        } else {
            part = (g.c / 8) * 2048;
            d = img_a + ((int64_t)b * g.jt + tile) * g.a_bytes + kg * 2048 + r * 16;
        }
        split_store_bf(d, part, v);
    }
}

// ---------------------------------------------------------------- GEMM + running column max
// 192 threads: warps 0-3 epilogue (TMEM lane quarter = warp), warp 4 lane 0 issuer, warp 5 lane 0 bulk copies.
__global__ void __launch_bounds__(192, 1)
corr_max_kernel(const unsigned char* __restrict__ img_a, const unsigned char* __restrict__ img_b, float* __restrict__ out, CmGeo g) {
    extern __shared__ __align__(128) unsigned char cm_smem[];
    unsigned char* sA = cm_smem;
    unsigned char* sS = cm_smem + g.a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sS + g.ns * g.s_bytes);   // a_full[0] full[1..3] empty[4..6] acc_full[7,8] acc_empty[9,10]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int jt = blockIdx.x, b = blockIdx.y;
    const int tmem_cols = 2 * g.ni;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        for (int i = 0; i < 3; ++i) { mbar_init(&bars[1 + i], 1); mbar_init(&bars[4 + i], 1); }
        mbar_init(&bars[7], 1); mbar_init(&bars[8], 1);
        mbar_init(&bars[9], 128); mbar_init(&bars[10], 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const int nstage_total = g.nt * g.kc;

    if (warp == 5) {
        if (lane == 0) {
            // ================= bulk copies: the target tile once, then the source stages =================
            {
                const uint32_t bar = smem_u32(&bars[0]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(g.a_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sA)),
                             "l"(img_a + ((int64_t)b * g.jt + jt) * g.a_bytes), "r"(g.a_bytes), "r"(bar)
                             : "memory");
            }
            const unsigned char* sb = img_b + (int64_t)b * nstage_total * g.s_bytes;
            for (int s = 0; s < nstage_total; ++s) {
                const int slot = s % g.ns;
                if (s >= g.ns) mbar_wait(&bars[4 + slot], ((s / g.ns) - 1) & 1);
                const uint32_t bar = smem_u32(&bars[1 + slot]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(g.s_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(sS + slot * g.s_bytes)),
                             "l"(sb + (int64_t)s * g.s_bytes), "r"(g.s_bytes), "r"(bar)
                             : "memory");
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            // ================= issuer =================
            const uint32_t idesc = umma_idesc_bf16(128, g.ni);
            const uint32_t a_part = (uint32_t)(g.c / 8) * 2048, s_part = (uint32_t)8 * g.ni * 16;
            mbar_wait(&bars[0], 0);
            int s = 0;
            for (int t = 0; t < g.nt; ++t) {
                const int acc = t & 1;
                if (t >= 2) mbar_wait(&bars[9 + acc], ((t >> 1) - 1) & 1);          // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + acc * g.ni;
                for (int kc = 0; kc < g.kc; ++kc, ++s) {
                    const int slot = s % g.ns;
                    mbar_wait(&bars[1 + slot], (s / g.ns) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(sA) + (uint32_t)kc * 8 * 2048, sb = smem_u32(sS + slot * g.s_bytes);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t dA1 = umma_desc(sa + ks * 2 * 2048, 2048, 128), dA2 = dA1 + (uint64_t)(a_part >> 4);
                        const uint64_t dB1 = umma_desc(sb + ks * 2 * g.ni * 16, g.ni * 16, 128), dB2 = dB1 + (uint64_t)(s_part >> 4);
                        umma_bf16(d, dA1, dB1, idesc, kc > 0 || ks > 0);
                        umma_bf16(d, dA1, dB2, idesc, true);
                        umma_bf16(d, dA2, dB1, idesc, true);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[4 + slot])) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[7 + acc])) : "memory");
            }
        }
    } else {
        // ================= epilogue warps: running max over the columns of every source tile =================
        float best = -INFINITY;
        for (int t = 0; t < g.nt; ++t) {
            const int acc = t & 1;
            mbar_wait(&bars[7 + acc], (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int i0 = t * g.ni;
            for (int col0 = 0; col0 < g.ni; col0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * g.ni + col0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (i0 + col0 + 32 <= g.n) {                                    // warp-uniform
#pragma unroll
                    for (int k = 0; k < 32; ++k) best = fmaxf(best, __uint_as_float(v[k]));
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (i0 + col0 + k < g.n) best = fmaxf(best, __uint_as_float(v[k]));   // padding columns are not source pixels
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[9 + acc])) : "memory");
        }
        const int j = jt * 128 + warp * 32 + lane;
        if (j < g.n) out[(int64_t)b * g.n + j] = best;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
}

}  // namespace ffwm

// Bytes of device workspace ffwm_corr_max needs (0: unsupported shape — C must be a multiple of 64 up to 256).
extern "C" int64_t ffwm_corr_max_workspace_bytes(int b, int c, int n) {
    ffwm::CmGeo g;
    if (!ffwm::cm_setup(g, b, c, n, 0.f)) return 0;
    return (int64_t)b * ((int64_t)g.jt * g.a_bytes + (int64_t)g.nt * g.kc * g.s_bytes);
}

// cmax (B, N) = max over source pixels i of  <source[b,:,i] / (|source[b,:,i]| + eps), target[b,:,j] / (|target[b,:,j]| + eps)>
// source, target: (B, C, H, W) fp32 with contiguous (H, W) planes (N = H*W), equal shapes.  Replaces the bmm + max of
// models/losses.py:347-353.  No gradient (none is defined in the reference's use).
extern "C" int ffwm_corr_max(const ffwm_tensor4* source, const ffwm_tensor4* target, float eps, float* cmax, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    using namespace ffwm;
    View<const float> sv, tv;
    int rc;
    if ((rc = make_view<const float>(source, "source", &sv))) return rc;
    if ((rc = make_view<const float>(target, "target", &tv))) return rc;
    if (sv.n != tv.n || sv.c != tv.c || sv.h != tv.h || sv.w != tv.w) { set_error("corr_max: source and target shapes differ"); return FFWM_ERR_SHAPE; }
    if ((int64_t)sv.n * sv.h * sv.w == 0) return FFWM_OK;
    if (sv.sw != 1 || sv.sh != sv.w || tv.sw != 1 || tv.sh != tv.w || sv.sb != tv.sb || sv.sc != tv.sc) {
        set_error("corr_max: (H, W) planes must be contiguous and both tensors laid out alike");
        return FFWM_ERR_SHAPE;
    }
    CmGeo g;
    if (!cm_setup(g, sv.n, sv.c, sv.h * sv.w, eps)) { set_error("corr_max: C = %d must be a multiple of 64 up to 256", sv.c); return FFWM_ERR_ARG; }
    const int64_t need = ffwm_corr_max_workspace_bytes(sv.n, sv.c, sv.h * sv.w);
    if (!cmax || !workspace || workspace_bytes < need) { set_error("corr_max: workspace too small (%lld < %lld bytes) or null output", (long long)workspace_bytes, (long long)need); return FFWM_ERR_SHAPE; }
    if (g.b > 65535) { set_error("corr_max: batch too large"); return FFWM_ERR_TOO_LARGE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char* img_a = static_cast<unsigned char*>(workspace);
    unsigned char* img_b = img_a + (int64_t)g.b * g.jt * g.a_bytes;
    const int npad = std::max(g.jt * 128, g.nt * g.ni);
    corr_max_prep_kernel<<<dim3((npad + 127) / 128, g.b, 2), 128, 0, st>>>(sv.p, tv.p, sv.sb, sv.sc, img_a, img_b, g);
    if ((rc = check_launch("corr_max (prepass)"))) return rc;
    const int smem = g.a_bytes + g.ns * g.s_bytes + 256;
    cudaError_t e = cudaFuncSetAttribute(corr_max_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("corr_max: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    corr_max_kernel<<<dim3(g.jt, g.b), 192, smem, st>>>(img_a, img_b, cmax, g);
    return check_launch("corr_max");
}
