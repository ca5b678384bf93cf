// conv3x3_wgrad_tc: weight gradient of the 3x3 / stride 1 / pad 1 convolution on the tcgen05 tensor
// cores, fp32 in / fp32 out, 3xTF32 operand split (same accuracy contract as conv3x3_tc.cu).
//
// Validated on a B200 (round 2, call 1: 13 parity cases <= 1.6e-5 of max|dW|; cuDNN's strict-fp32 weight gradient
// measures 3-7e-5 on the same inputs).  Operand math is a template parameter like conv3x3_tc.cu's (the `math` argument):
// 3xTF32 (4 pixels per 16-byte chunk, K = 8 pixels per MMA) or 3xBF16 (8 pixels per chunk, K = 16 pixels per MMA:
// same shared-memory tiles, half the MMAs).  The text below describes the 3xTF32 geometry; PPC = pixels per chunk.
//
//   dW[co, ci, ky, kx] = sum over (b, y, x) of  gO[b, co, y, x] * X[b, ci, y + ky - 1, x + kx - 1]
//
// is a GEMM whose contraction runs over PIXELS, and NCHW keeps the pixels of a channel contiguous, so
// both operands are K-major as they lie in HBM (the forward kernel has to transpose its activations).
//
//   * M = 128 output channels (TMEM lanes), N = 3 x 48 = 144 = (ky, ci): the three vertical taps of 48
//     input channels are three groups of rows of ONE B operand (input rows y-1, y, y+1 staged one
//     after the other), K = 8 pixels per MMA (kind::tf32).
//   * the horizontal taps are descriptor start addresses.  A 16-byte K chunk does not hold four
//     neighbouring pixels but the four pixels {j, j+S, j+2S, j+3S}, S = W/4: then "one pixel to the
//     left" of chunk j is chunk j-1 for all four of its elements, i.e. the same staged B tile read
//     from a start address one chunk (LBO bytes) lower.  Two halo chunks (j0-1, j0+8) per stage hold
//     the neighbours of the stage's first / last chunk; elements that fall off the row are zero.
//     So X is staged ONCE for the nine taps (no shifted copies).
//   * three accumulators (kx = 0, 1, 2) of 128 x 144 fp32 in TMEM = 432 of the 512 columns.
//   * per MMA the tensor core reads A (128 rows x 32 B = 4 KB) and B (144 x 32 B = 4.5 KB) from shared
//     memory: ~67 clk at 128 B/clk against ~74 clk of tf32 math, which is why N packs the vertical taps —
//     nine separate N = 48 MMAs per K step would re-read A nine times and be shared-memory bound.
//   * split-K over (image, row, 32-pixel block) stages, one CTA per SM, at most WG_MAX_STAGES stages per CTA
//     (bounds the truncating fp32 accumulation chain); partial sums are reduced with fp32 REDs into the
//     caller's zero-filled dW (once per CTA, 128 x 432 values).
//   * warp-specialised 2-stage mbarrier pipeline as in conv3x3_tc.cu: 8 producer warps stage (split
//     hi/lo, 16-byte conflict-free stores), lane 0 of warp 8 issues 36 MMAs per stage and commits.
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace ffwm {

constexpr int WG_MT = 128;                              // output channels per CTA = MMA M
constexpr int WG_NCI = 48;                              // input channels per CTA
constexpr int WG_N = 3 * WG_NCI;                        // MMA N = (ky, ci)
constexpr int WG_CH = 8;                                // K chunks (of 4 pixels) per stage
constexpr int WG_PRODUCERS = 256;
constexpr int WG_A_LBO = WG_MT * 16;                    // bytes between consecutive chunks of A
constexpr int WG_A_PART = WG_CH * WG_A_LBO;             // hi or lo of A
constexpr int WG_B_LBO = WG_N * 16;
constexpr int WG_B_PART = (WG_CH + 2) * WG_B_LBO;       // chunks j0-1 .. j0+8
constexpr int WG_STAGE = 2 * WG_A_PART + 2 * WG_B_PART; // [A hi][A lo][B hi][B lo]
constexpr int WG_SMEM = 2 * WG_STAGE + 64;              // + 4 mbarriers + tmem address
constexpr int WG_TMEM_COLS = 512;
constexpr int WG_A_ITEMS = WG_MT / 8;                   // producer work items: 8 rows x all chunks of a stage
constexpr int WG_ITEMS = WG_A_ITEMS + WG_N / 8;
// Longest accumulation chain of one CTA, in stages (12 accumulator updates each).  The tensor core adds into its
// fp32 accumulators with truncation, so the error grows linearly with the number of updates: ~2e-5 of max|dW|
// after 768 updates (the forward kernel's measured 1-2e-5 at 675 updates; simulated: 1e-4 at 4096, 1.4e-3 at the
// 49152 a whole 8x128x128 batch would need).  Partial sums of the splits meet in fp32 REDs (round to nearest).
constexpr int WG_MAX_STAGES = 64;
constexpr int WG_EPI_PITCH = 9 * WG_NCI + 1;             // floats per row of the epilogue tile (odd: conflict-free columns)
constexpr int WG_ITEMS_PER_WARP = (WG_ITEMS + WG_PRODUCERS / 32 - 1) / (WG_PRODUCERS / 32);
static_assert(WG_A_LBO % 128 == 0 && WG_B_LBO % 128 == 0 && WG_STAGE % 128 == 0, "operand tiles stay 128-byte aligned");
static_assert(3 * WG_N <= WG_TMEM_COLS && WG_N % 16 == 0 && WG_N <= 256, "three accumulators of N columns");
static_assert(WG_SMEM <= 227 * 1024, "shared memory");
static_assert(WG_A_ITEMS == 2 * (WG_PRODUCERS / 32), "the bias-gradient partial sums assume two A items per producer warp");
static_assert(64 * WG_EPI_PITCH * 4 <= 2 * WG_STAGE, "the epilogue tile reuses the stage buffers");

struct WgGeo {
    int cout, cin, h, w;
    int s4;              // W / PPC: pixel stride between the elements of a chunk (PPC = 4 for 3xTF32, 8 for 3xBF16)
    int nch;             // chunks per stage = min(WG_CH, s4): even, divides s4
    int ncb;             // stages per image row = s4 / nch
    int n_ci_tiles;
    int stages_total;    // B * H * ncb
    int stages_per_split;
    int direct_epilogue; // 1: REDs straight from the TMEM registers (FFWM_WGRAD_DIRECT_EPILOGUE, A/B and fallback)
};

template <bool BF>
__global__ void __launch_bounds__(WG_PRODUCERS + 32, 1)
conv3x3_wgrad_tc_kernel(View<const float> x, View<const float> go, float* __restrict__ dw, int64_t s_co, int64_t s_ci,
                        int64_t s_ky, int64_t s_kx, float* __restrict__ dbias, WgGeo g) {
    constexpr int PPC = BF ? 8 : 4;                                          // pixels per 16-byte chunk
    extern __shared__ __align__(128) unsigned char wg_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(wg_smem + 2 * WG_STAGE);    // full[0,1] empty[2,3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wg_smem + 2 * WG_STAGE + 32);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cot = blockIdx.x / g.n_ci_tiles, cit = blockIdx.x % g.n_ci_tiles;
    const int st0 = blockIdx.y * g.stages_per_split;
    const int nst = min(g.stages_per_split, g.stages_total - st0);          // >= 1 (host)

    if (tid == 0) {
        mbar_init(&bars[0], WG_PRODUCERS);
        mbar_init(&bars[1], WG_PRODUCERS);
        mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(WG_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp < WG_PRODUCERS / 32) {
        // ================= producers =================
        // A work item is 8 operand rows (channels) x every chunk of the stage; lane -> (row r8, chunk phase jq):
        // the 8 lanes of a quarter-warp write 8 consecutive 16-byte rows (conflict-free), and a load
        // instruction covers 8 rows x 4 consecutive pixels.  Every slot is rewritten every stage (zeros where
        // the channel, the image row or the pixel does not exist), so nothing relies on an initial fill.
        const int r8 = lane & 7, jq = lane >> 3;
        // bias gradient = sum of grad_out over (b, y, x): the CTAs of input-channel tile 0 see every grad_out
        // element of their 128 output channels exactly once while staging A, so they add them up on the way
        // (items 0..15 are the A rows; a warp owns items warp and warp + 8, i.e. i = 0, 1)
        const bool do_bias = dbias != nullptr && cit == 0;
        float bsum[2] = {0.f, 0.f};
        for (int k = 0; k < nst; ++k) {
            const int buf = k & 1;
            if (k >= 2) mbar_wait(&bars[2 + buf], ((k >> 1) - 1) & 1);      // MMAs of stage k-2 done
            const int s = st0 + k;
            const int j0 = (s % g.ncb) * g.nch, rr = s / g.ncb, y = rr % g.h, b = rr / g.h;
            unsigned char* sbase = wg_smem + buf * WG_STAGE;
            // all loads of the stage in flight first (up to 5 items x 3 chunks x PPC pixels per lane), then split + store
            float v[WG_ITEMS_PER_WARP][3][PPC];
#pragma unroll
            for (int i = 0; i < WG_ITEMS_PER_WARP; ++i) {
                const int it = warp + i * (WG_PRODUCERS / 32);               // warp-uniform
                if (it >= WG_ITEMS) continue;
                const float* rowp;
                bool rok;
                int jbase, nchunk, sw;
                if (it < WG_A_ITEMS) {
                    const int co = cot * WG_MT + it * 8 + r8;
                    rok = co < g.cout;
                    rowp = go.p + b * go.sb + (int64_t)co * go.sc + (int64_t)y * go.sh;
                    jbase = j0, nchunk = g.nch, sw = go.sw;
                } else {
                    const int n = (it - WG_A_ITEMS) * 8 + r8, ky = n / WG_NCI, ci = cit * WG_NCI + n % WG_NCI;
                    const int yy = y + ky - 1;
                    rok = ci < g.cin && (unsigned)yy < (unsigned)g.h;
                    rowp = x.p + b * x.sb + (int64_t)ci * x.sc + (int64_t)yy * x.sh;
                    jbase = j0 - 1, nchunk = g.nch + 2, sw = x.sw;
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int cc = jq + 4 * q;
#pragma unroll
                    for (int m = 0; m < PPC; ++m) {
                        const int p = jbase + cc + m * g.s4;
                        const bool ok = rok && cc < nchunk && (unsigned)p < (unsigned)g.w;
                        v[i][q][m] = ok ? __ldg(rowp + (int64_t)p * sw) : 0.f;
                    }
                }
                if (i < 2 && do_bias) {                                      // A item (zeros where the row / chunk does not exist)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int m = 0; m < PPC; ++m) bsum[i] += v[i][q][m];
                }
            }
#pragma unroll
            for (int i = 0; i < WG_ITEMS_PER_WARP; ++i) {
                const int it = warp + i * (WG_PRODUCERS / 32);
                if (it >= WG_ITEMS) continue;
                const bool isA = it < WG_A_ITEMS;
                const int row = (isA ? it : it - WG_A_ITEMS) * 8 + r8;
                const int nchunk = isA ? g.nch : g.nch + 2, lbo = isA ? WG_A_LBO : WG_B_LBO, part = isA ? WG_A_PART : WG_B_PART;
                unsigned char* d0 = sbase + (isA ? 0 : 2 * WG_A_PART) + row * 16;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const int cc = jq + 4 * q;
                    if (cc < nchunk) split_store_m<BF>(d0 + cc * lbo, part, v[i][q]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // my stores -> visible to the tensor core
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[buf])) : "memory");
        }
        if (do_bias) {                                                       // warp-uniform
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float t = bsum[i];
                t += __shfl_xor_sync(0xffffffffu, t, 8);                     // the four chunk phases jq of a row
                t += __shfl_xor_sync(0xffffffffu, t, 16);
                const int co = cot * WG_MT + (warp + i * (WG_PRODUCERS / 32)) * 8 + r8;
                if (jq == 0 && co < g.cout) red_add(dbias + co, t);
            }
        }
    } else if (lane == 0) {
        // ================= issuer =================
        constexpr uint32_t IDESC = BF ? umma_idesc_bf16(WG_MT, WG_N) : umma_idesc_tf32(WG_MT, WG_N);
        for (int k = 0; k < nst; ++k) {
            const int buf = k & 1;
            mbar_wait(&bars[buf], (k >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sA = smem_u32(wg_smem + buf * WG_STAGE), sB = sA + 2 * WG_A_PART;
            const uint64_t dA0 = umma_desc(sA, WG_A_LBO, 128), dB0 = umma_desc(sB, WG_B_LBO, 128);
#pragma unroll
            for (int t = 0; t < WG_CH / 2; ++t) {                            // K step: chunks 2t, 2t+1
                if (2 * t >= g.nch) break;
                const uint64_t dA_hi = dA0 + (uint64_t)((2 * t * WG_A_LBO) >> 4), dA_lo = dA_hi + (WG_A_PART >> 4);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {                             // B chunk index = (j - j0 + 1) + (kx - 1)
                    const uint64_t dB_hi = dB0 + (uint64_t)(((2 * t + kx) * WG_B_LBO) >> 4), dB_lo = dB_hi + (WG_B_PART >> 4);
                    const uint32_t d = tmem + kx * WG_N;
                    umma_ss<BF>(d, dA_hi, dB_hi, IDESC, t > 0 || k > 0);
                    umma_ss<BF>(d, dA_hi, dB_lo, IDESC, true);
                    umma_ss<BF>(d, dA_lo, dB_hi, IDESC, true);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[2 + buf])) : "memory");
        }
    }

    // ---- epilogue (producer warps).  All MMAs have completed when the last commit arrives, so the stage buffers
    // are free.  Default: the accumulators go through shared memory in dW's own (ci, ky, kx) order, 64 output
    // channels at a time, and leave as REDs whose lanes are consecutive addresses of one dW row (a CTA's partial
    // result is 128 rows of up to 432 contiguous floats); g.direct_epilogue issues the REDs straight from the
    // TMEM registers instead (lanes = output channels: one 32-byte sector per lane and instruction).
    if (warp < WG_PRODUCERS / 32) {
        mbar_wait(&bars[2 + ((nst - 1) & 1)], ((nst - 1) >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, half = warp >> 2;
        float* tile = reinterpret_cast<float*>(wg_smem);                     // [64 rows][WG_EPI_PITCH]
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            // every warp walks its TMEM blocks in both passes (tcgen05.ld is warp-collective and cheap); a warp
            // whose lane quarter belongs to the other half of the output channels keeps nothing
            const bool mine = g.direct_epilogue ? pass == 0 : (q >> 1) == pass;
            if (g.direct_epilogue && pass == 1) break;
            const int co = cot * WG_MT + q * 32 + lane;
#pragma unroll 1
            for (int blk = half; blk < 3 * WG_N / 16; blk += 2) {
                if (!mine) continue;                                         // warp-uniform
                const int c0 = blk * 16;
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int kx = c0 / WG_N, n0 = c0 % WG_N, ky = n0 / WG_NCI, cil = n0 % WG_NCI;
                if (g.direct_epilogue) {
                    if (co < g.cout) {
                        float* dp = dw + (int64_t)co * s_co + ky * s_ky + kx * s_kx;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (cit * WG_NCI + cil + j < g.cin) red_add(dp + (int64_t)(cit * WG_NCI + cil + j) * s_ci, __uint_as_float(v[j]));
                    }
                } else {
                    float* tp = tile + ((q & 1) * 32 + lane) * WG_EPI_PITCH + ky * 3 + kx;   // odd pitch: lanes -> distinct banks
#pragma unroll
                    for (int j = 0; j < 16; ++j) tp[(cil + j) * 9] = __uint_as_float(v[j]);
                }
            }
            if (g.direct_epilogue) break;
            asm volatile("bar.sync 1, %0;" ::"n"(WG_PRODUCERS) : "memory");  // tile complete (producer warps only)
            const int nci = min(WG_NCI, g.cin - cit * WG_NCI);               // valid input channels of this tile
#pragma unroll 1
            for (int r = warp; r < 64; r += WG_PRODUCERS / 32) {
                const int co_r = cot * WG_MT + pass * 64 + r;
                if (co_r >= g.cout) break;                                   // warp-uniform; rows are ascending
                float* dp = dw + (int64_t)co_r * s_co + (int64_t)(cit * WG_NCI) * s_ci;
                const float* tr = tile + r * WG_EPI_PITCH;
                for (int e = lane; e < nci * 9; e += 32) {
                    const int ci_l = e / 9, t9 = e - ci_l * 9, ky = t9 / 3, kx = t9 - ky * 3;
                    red_add(dp + (int64_t)ci_l * s_ci + ky * s_ky + kx * s_kx, tr[e]);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(WG_PRODUCERS) : "memory");  // tile drained before the next pass refills it
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(WG_TMEM_COLS) : "memory");
}

}  // namespace ffwm

// grad_weight (Cout,Cin,3,3) += d/dw of conv2d(x, w, stride 1, padding 1) for grad_out; the caller zero-fills
// grad_weight (partial sums of the K splits arrive as REDs).  x (B,Cin,H,W), grad_out (B,Cout,H,W), W % 32 == 0.
// grad_bias (Cout floats, zero-filled by the caller, may be NULL) += sum of grad_out over (b, y, x).
extern "C" int ffwm_conv3x3_wgrad(const ffwm_tensor4* x, const ffwm_tensor4* grad_out, const ffwm_tensor4* grad_weight,
                                  float* grad_bias, int math, void* stream) {
    using namespace ffwm;
    View<const float> xv, gv;
    View<float> wv;
    int rc;
    if ((rc = make_view<const float>(x, "x", &xv))) return rc;
    if ((rc = make_view<const float>(grad_out, "grad_out", &gv))) return rc;
    if ((rc = make_view<float>(grad_weight, "grad_weight", &wv))) return rc;
    if (math != 0 && math != 1) { set_error("conv3x3_wgrad: math must be 0 (3xTF32) or 1 (3xBF16)"); return FFWM_ERR_ARG; }
    const bool bf = math != 0;
    if (xv.w % 32 != 0 || xv.w <= 0 || gv.w != xv.w || gv.h != xv.h || gv.n != xv.n || wv.n != gv.c || wv.c != xv.c || wv.h != 3 || wv.w != 3) {
        set_error("conv3x3_wgrad: needs W %% 32 == 0, equal N,H,W and grad_weight (Cout,Cin,3,3) (x %dx%dx%dx%d, grad_out %dx%dx%dx%d, grad_weight %dx%dx%dx%d)",
                  xv.n, xv.c, xv.h, xv.w, gv.n, gv.c, gv.h, gv.w, wv.n, wv.c, wv.h, wv.w);
        return FFWM_ERR_SHAPE;
    }
    if ((int64_t)wv.n * wv.c == 0 || (int64_t)xv.n * xv.h == 0) return FFWM_OK;
    WgGeo g;
    g.cout = gv.c, g.cin = xv.c, g.h = xv.h, g.w = xv.w;
    g.s4 = xv.w / (bf ? 8 : 4);                 // W % 32 == 0: s4 is a multiple of 4 (3xBF16) / 8 (3xTF32)
    g.nch = g.s4 % WG_CH == 0 ? WG_CH : 4;     // chunks per stage: even, divides s4
    g.ncb = g.s4 / g.nch;
    g.n_ci_tiles = ceil_div(g.cin, WG_NCI);
    const int64_t tiles = (int64_t)ceil_div(g.cout, WG_MT) * g.n_ci_tiles;
    const int64_t stages = (int64_t)xv.n * xv.h * g.ncb;
    if (stages > 0x7fffffffLL || tiles > 0x7fffffffLL) { set_error("conv3x3_wgrad: problem too large"); return FFWM_ERR_TOO_LARGE; }
    g.stages_total = (int)stages;
    // split the pixel range until tiles x splits fills the SMs, and further until no CTA accumulates more than
    // WG_MAX_STAGES stages (accuracy, see above; the extra CTAs run as further waves)
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(stages, sm_count() / tiles));
    g.stages_per_split = std::min(ceil_div(stages, splits), WG_MAX_STAGES);
    splits = ceil_div(stages, g.stages_per_split);                            // every split owns >= 1 stage
    g.direct_epilogue = opt(OPT_WGRAD_DIRECT_EPILOGUE) ? 1 : 0;
    if (splits > 65535) { set_error("conv3x3_wgrad: grid too large"); return FFWM_ERR_TOO_LARGE; }
    auto kern = bf ? conv3x3_wgrad_tc_kernel<true> : conv3x3_wgrad_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) { set_error("conv3x3_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid((unsigned)tiles, (unsigned)splits);
    kern<<<grid, WG_PRODUCERS + 32, WG_SMEM, static_cast<cudaStream_t>(stream)>>>(
        xv, gv, wv.p, wv.sb, wv.sc, (int64_t)wv.sh, (int64_t)wv.sw, grad_bias, g);
    return check_launch("conv3x3_wgrad");
}
