// conv3x3_tc: 3x3 / stride 1 / pad 1 convolution on the 5th-generation tensor cores (tcgen05),
// fp32 in, fp32 out, fp32-level accuracy through a 3xTF32 operand split.
//
// Where it is used: the dense 3x3 convolutions at 128x128 that dominate the FFWM generator
// (dres2: 195->195, att2: 128->128; SURVEY.md 8a a13 — 62 of netG's 105 GFLOP/image) and their
// data gradients (same kernel, weights packed transposed + flipped).  The reference runs them on
// cuDNN; under its fp32 parity target that means SIMT FFMA kernels (~35 TFLOP/s measured in the
// train step).  Here they are implicit GEMMs on tcgen05:
//
//   D[pixel, co] += A[pixel, ci] * B[co, ci]      for each of the 9 taps, ci in blocks of 8
//
//   * M = 128 pixels = one image row (W = 128), two (W = 64), four (W = 32) or eight (W = 16), so TMEM lanes are
//     consecutive x and the epilogue's global stores are coalesced rows of the NCHW output;
//     N = 64 output channels; a CTA owns 2-4 M tiles x 64 channels (128-256 TMEM columns).
//   * A (activations) is staged K-major without swizzle as [k-chunk][slot][4 channels]: 16 bytes
//     per pixel slot, slot = x + 1 with explicit zero halo slots, so the three horizontal taps
//     are the SAME shared-memory row read through descriptors whose start address differs by
//     16 bytes; the three vertical taps are neighbouring row buffers (6 input rows serve 4 output
//     rows).  For W < 128 an M tile spans several image rows, so three horizontally shifted
//     copies are staged instead and the vertical taps become start-address offsets (CvGeo).
//     The NCHW -> K-major transposition happens in the staging threads (4 coalesced loads ->
//     one 16-byte shared store per copy).
//   * B (weights) is pre-packed once per weight update into exactly the shared-memory image the
//     MMA wants, hi and lo parts already split, and copied in with 16-byte loads/stores.
//   * fp32 accuracy: every operand is split a = hi + lo with hi = a & 0xffffe000 (exactly
//     representable in TF32) and lo = a - hi; three MMAs hi*hi + hi*lo + lo*hi accumulate in
//     fp32 in TMEM (the dropped lo*lo term is < 2^-21 relative).
//   * warp-specialised 2-stage pipeline on mbarriers: 8 producer warps stage the activations of
//     K block k+1 while one elected thread of a 9th warp has the 108 MMAs of block k in flight
//     (4 rows x 9 taps x 3 split terms, committed to an mbarrier with tcgen05.commit) and pulls the
//     next packed weight tile with one bulk async copy (cp.async.bulk + complete_tx).
#include <stdint.h>

#include <algorithm>

#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace ffwm {

// Output channels per CTA = MMA N: 64, or 128 for W = 128 (NT is a template parameter).  With N = 64 an MMA reads
// 4 KB of A and 2 KB of B from shared memory for 33 clk of math (48 clk at 128 B/clk: shared-memory bound, ncu
// l1tex 72 % / tensor pipe 48 %); N = 128 reads 4 + 4 KB for 66 clk of math.  Callers choose per call (the nt argument
// of the entry points); ffwm_b200/conv.py takes N = 128 wherever W = 128 and Cout > 64 (FFWM_CONV_NT128=0 for the A/B;
// measured in profiles/r02a_conv_nt128.txt).
// Operand math (template parameter BF of everything below; chosen per call: the `math` argument of the entry points):
//   BF = false  3xTF32: a = hi + lo, hi = a & 0xffffe000 (exact in tf32); hi*hi + hi*lo + lo*hi; K block = 8 channels
//   BF = true   3xBF16: a = b1 + b2 + (dropped), b1 = bf16_rn(a), b2 = bf16_rn(a - b1); b1*b1 + b1*b2 + b2*b1;
//               K block = 16 channels.  A 16-byte slot holds 8 bf16 channels instead of 4 fp32, so every
//               shared-memory tile has the SAME size and layout, each MMA contracts twice as many channels at the
//               same cost (kind::f16 runs at twice the kind::tf32 rate), and the kernel — shared-memory bound, see
//               NT below — does half the MMAs.  Dropped terms: b1*b3, b2*b2, b3*b1 <= 3 * 2^-17 per product, random
//               sign: 4-6e-6 of max|out| over the step's K range (CPU simulation, profiles/README.md), the same order
//               as the tensor core's truncating fp32 accumulation (1-2e-5 measured) and inside the path's 1e-4.
template <bool BF> struct CvMath { static constexpr int KB = BF ? 16 : 8, CPS = BF ? 8 : 4; };   // channels per K block / per slot
constexpr int CV_PRODUCERS = 256;

// Geometry per image width WI (= 128, 64, 32 or 16).  The MMA M dimension is always 128 pixels:
//   WI = 128: an M tile is one image row; one copy of each input row with explicit zero halo
//             slots (slot = x + 1), horizontal taps = descriptor start + kx * 16 bytes.
//   WI < 128: an M tile is 128/WI consecutive image rows, stored back to back (slot = row*WI + x),
//             so a horizontal shift cannot be expressed by the start address across the row
//             seam; three copies are staged instead, copy kx holding in[.., x + kx - 1] with zeros
//             at the border; vertical taps = descriptor start + ky * WI * 16 bytes.
template <int WI, int NT>
struct CvGeo {
    static constexpr int RPT = 128 / WI;                       // image rows per M tile
    static constexpr int MT = (WI == 128 && NT == 64) ? 4 : 2; // M tiles (accumulators) per CTA
    static constexpr int ROWS = MT * RPT;                      // output rows per CTA
    static constexpr int IN_ROWS = ROWS + 2;
    static constexpr int NCOPY = WI == 128 ? 1 : 3;
    static constexpr int ROW_SLOTS = WI == 128 ? 136 : WI;     // 16-byte slots per input row and k-chunk
    static constexpr int A_CHUNK = IN_ROWS * ROW_SLOTS * 16;   // one k-chunk (4 channels), all rows, one copy
    static constexpr int A_COPY = 2 * A_CHUNK;                 // both k-chunks
    static constexpr int A_PART = NCOPY * A_COPY;              // hi or lo
    static constexpr int A_STAGE = 2 * A_PART;
    static constexpr int B_CHUNK = NT * 16;                    // one k-chunk of one tap: NT co x 4 ci
    static constexpr int B_TAP = 2 * B_CHUNK;                  // both k-chunks
    static constexpr int B_STAGE = 2 * 9 * B_TAP;              // layout: [tap][hl][kchunk][co][4]
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int SMEM = 2 * STAGE + 64;                // + 6 mbarriers + tmem address
    static constexpr int TMEM_COLS = 256;
    static_assert(MT * NT <= TMEM_COLS && SMEM <= 227 * 1024, "TMEM columns / shared memory");
};

// ---------------------------------------------------------------- weight packing
// w (Cout, Cin, 3, 3) -> packed[cob][kb][tap][hl][kchunk][co_local NT][4 ci]  (hl: 0 = hi, 1 = lo; NT = nt)
// dgrad = 1 packs the weights of the data-gradient convolution: roles of Cout/Cin swapped and the
// taps flipped, so that conv3x3(grad_output, packed) = grad_input.
template <int nt>
__global__ void conv3x3_pack_kernel(const float* __restrict__ w, float* __restrict__ packed, int cout, int cin, int dgrad,
                                    int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx) {
    const int n_out = dgrad ? cin : cout, n_in = dgrad ? cout : cin;     // of the convolution being packed
    const int ncob = (n_out + nt - 1) / nt, nkb = (n_in + 7) / 8;
    const int64_t total = (int64_t)ncob * nkb * 9 * 2 * nt * 4;       // (hi,lo) pairs are written together
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int j = r % 4; r /= 4;
        const int col = r % nt; r /= nt;
        const int kc = r % 2; r /= 2;
        const int tap = r % 9; r /= 9;
        const int kb = r % nkb; r /= nkb;
        const int cob = (int)r;
        const int o = cob * nt + col, c = kb * 8 + kc * 4 + j;
        float v = 0.f;
        if (o < n_out && c < n_in) {
            const int ky = tap / 3, kx = tap % 3;
            v = dgrad ? w[c * s_co + o * s_ci + (2 - ky) * s_ky + (2 - kx) * s_kx]
                      : w[o * s_co + c * s_ci + ky * s_ky + kx * s_kx];
        }
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const float lo = v - hi;
        const int64_t base = ((((int64_t)cob * nkb + kb) * 9 + tap) * 2) * (2 * nt * 4);
        const int64_t within = ((int64_t)kc * nt + col) * 4 + j;
        packed[base + within] = hi;
        packed[base + 2 * nt * 4 + within] = lo;
    }
}

// The 3xBF16 image: packed[cob][kb][tap][part][kchunk][co_local NT][8 ci] of bf16 (part: 0 = b1, 1 = b2), K block = 16.
template <int nt>
__global__ void conv3x3_pack_bf_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, int cout, int cin, int dgrad,
                                       int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx) {
    const int n_out = dgrad ? cin : cout, n_in = dgrad ? cout : cin;
    const int ncob = (n_out + nt - 1) / nt, nkb = (n_in + 15) / 16;
    const int64_t total = (int64_t)ncob * nkb * 9 * 2 * nt * 8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i;
        const int j = r % 8; r /= 8;
        const int col = r % nt; r /= nt;
        const int kc = r % 2; r /= 2;
        const int tap = r % 9; r /= 9;
        const int kb = r % nkb; r /= nkb;
        const int cob = (int)r;
        const int o = cob * nt + col, c = kb * 16 + kc * 8 + j;
        float v = 0.f;
        if (o < n_out && c < n_in) {
            const int ky = tap / 3, kx = tap % 3;
            v = dgrad ? w[c * s_co + o * s_ci + (2 - ky) * s_ky + (2 - kx) * s_kx]
                      : w[o * s_co + c * s_ci + ky * s_ky + kx * s_kx];
        }
        const __nv_bfloat16 b1 = __float2bfloat16_rn(v);
        const __nv_bfloat16 b2 = __float2bfloat16_rn(v - __bfloat162float(b1));
        const int64_t base = ((((int64_t)cob * nkb + kb) * 9 + tap) * 2) * (2 * nt * 8);
        const int64_t within = ((int64_t)kc * nt + col) * 8 + j;
        packed[base + within] = b1;
        packed[base + 2 * nt * 8 + within] = b2;
    }
}

// ---------------------------------------------------------------- the convolution
// Warp roles (288 threads): warps 0-7 stage activations (producers) and run the epilogue; warp 8
// lane 0 copies the packed weights with one bulk async copy per K block and issues the MMAs.
// Pipeline state lives in mbarriers: fullA[2] (256 producer arrivals), fullB[2] (bulk-copy
// transaction bytes), empty[2] (tcgen05.commit of the MMAs that read the buffer).
template <int WI, int NT, bool BF>
__global__ void __launch_bounds__(CV_PRODUCERS + 32, 1)
conv3x3_tc_kernel(View<const float> x, const float* __restrict__ packed, const float* __restrict__ bias,
                  View<float> out, int nkb) {
    using G = CvGeo<WI, NT>;
    constexpr int KB = CvMath<BF>::KB, CPS = CvMath<BF>::CPS;
    extern __shared__ __align__(128) unsigned char cv_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(cv_smem + 2 * G::STAGE);    // fullA[0,1] fullB[2,3] empty[4,5]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(cv_smem + 2 * G::STAGE + 48);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int y0 = blockIdx.x * G::ROWS, cob = blockIdx.y, b = blockIdx.z;

    // ---- one-time setup: zero both stages (halo / border slots and out-of-image rows stay zero), barriers, TMEM
    for (int i = tid; i < (2 * G::STAGE) / 16; i += CV_PRODUCERS + 32) reinterpret_cast<float4*>(cv_smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        mbar_init(&bars[0], CV_PRODUCERS);
        mbar_init(&bars[1], CV_PRODUCERS);
        mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 1);
        mbar_init(&bars[4], 1);
        mbar_init(&bars[5], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(G::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // the zero fill, before any async-proxy access
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const float* pk_base = packed + ((int64_t)cob * nkb) * (G::B_STAGE / 4);

    if (warp < CV_PRODUCERS / 32) {
        // ================= producers: activations of K block kb -> stage buffer kb & 1 =================
        // thread -> (pixel px, k-chunk kc, row phase ph); it loads 4 channels of every PH-th input row
        // (all loads in flight), splits hi/lo and writes 16-byte slots
        constexpr int PH = CV_PRODUCERS / (2 * WI);                // 1, 2 or 4 row phases
        constexpr int NR = (G::IN_ROWS + PH - 1) / PH;             // rows per thread
        const int px = tid % WI, kc = (tid / WI) & 1, ph = tid / (2 * WI);
        const float* gp0 = x.p + b * x.sb + (int64_t)(y0 - 1) * x.sh + px * x.sw;
        // The loads go to registers, so K block kb+1 is LOADED before K block kb is split and stored (software pipeline,
        // two register sets): L2 / HBM latency hides behind a whole K-block period.
        auto load_kb = [&](int kb, float (&v)[NR][CPS]) {
            const int c0 = kb * KB + kc * CPS;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int rr = ph + i * PH;
                const bool ok = rr < G::IN_ROWS && (unsigned)(y0 - 1 + rr) < (unsigned)x.h;
#pragma unroll
                for (int j = 0; j < CPS; ++j)
                    v[i][j] = (ok && c0 + j < x.c) ? __ldg(gp0 + (int64_t)(c0 + j) * x.sc + rr * x.sh) : 0.f;
            }
        };
        auto store_kb = [&](int kb, float (&v)[NR][CPS]) {
            const int buf = kb & 1;
            if (kb >= 2) mbar_wait(&bars[4 + buf], ((kb >> 1) - 1) & 1);          // MMAs of K block kb-2 done
            unsigned char* sA = cv_smem + buf * G::STAGE + kc * G::A_CHUNK;
#pragma unroll
            for (int i = 0; i < NR; ++i) {
                const int rr = ph + i * PH;
                if (rr >= G::IN_ROWS || (unsigned)(y0 - 1 + rr) >= (unsigned)x.h) continue;   // stays zero
                if (WI == 128) {
                    split_store_m<BF>(sA + (rr * G::ROW_SLOTS + px + 1) * 16, G::A_PART, v[i]);
                } else {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int xd = px - kx + 1;                              // copy kx holds in[x + kx - 1] at slot x
                        if ((unsigned)xd < (unsigned)WI) split_store_m<BF>(sA + kx * G::A_COPY + (rr * WI + xd) * 16, G::A_PART, v[i]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // my stores -> visible to the tensor core
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[buf])) : "memory");
        };
        float va[NR][CPS], vb[NR][CPS];
        load_kb(0, va);
        for (int kb = 0; kb < nkb; kb += 2) {
            if (kb + 1 < nkb) load_kb(kb + 1, vb);
            store_kb(kb, va);
            if (kb + 1 < nkb) {
                if (kb + 2 < nkb) load_kb(kb + 2, va);
                store_kb(kb + 1, vb);
            }
        }
    } else if (lane == 0) {
        // ================= issuer: weight copies + MMAs =================
        constexpr uint32_t IDESC = BF ? umma_idesc_bf16(128, NT) : umma_idesc_tf32(128, NT);
        auto copy_b = [&](int kb) {
            const int buf = kb & 1;
            const uint32_t bar = smem_u32(&bars[2 + buf]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(G::B_STAGE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(cv_smem + buf * G::STAGE + G::A_STAGE)),
                         "l"(pk_base + (int64_t)kb * (G::B_STAGE / 4)), "r"(G::B_STAGE), "r"(bar)
                         : "memory");
        };
        copy_b(0);
        for (int kb = 0; kb < nkb; ++kb) {
            const int buf = kb & 1;
            mbar_wait(&bars[buf], (kb >> 1) & 1);            // activations staged
            mbar_wait(&bars[2 + buf], (kb >> 1) & 1);        // weights landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sA = smem_u32(cv_smem + buf * G::STAGE), sB = sA + G::A_STAGE;
            const uint64_t dA0 = umma_desc(sA, G::A_CHUNK, 128), dB0 = umma_desc(sB, G::B_CHUNK, 128);
#pragma unroll
            for (int t = 0; t < G::MT; ++t) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap % 3;
                    const int a_off = WI == 128 ? ((t + ky) * G::ROW_SLOTS + kx) * 16
                                                : kx * G::A_COPY + (t * G::RPT + ky) * WI * 16;
                    const uint64_t dA_hi = dA0 + (uint64_t)(a_off >> 4), dA_lo = dA_hi + (G::A_PART >> 4);
                    const uint64_t dB_hi = dB0 + (uint64_t)((tap * 2 * G::B_TAP) >> 4), dB_lo = dB_hi + (G::B_TAP >> 4);
                    const uint32_t d = tmem + t * NT;
                    umma_ss<BF>(d, dA_hi, dB_hi, IDESC, tap > 0 || kb > 0);
                    umma_ss<BF>(d, dA_hi, dB_lo, IDESC, true);
                    umma_ss<BF>(d, dA_lo, dB_hi, IDESC, true);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[4 + buf])) : "memory");
            if (kb + 1 < nkb) {
                if (kb >= 1) mbar_wait(&bars[4 + (buf ^ 1)], (((kb + 1) >> 1) - 1) & 1);   // MMAs of K block kb-1 done
                copy_b(kb + 1);
            }
        }
    }

    // ---- epilogue (producer warps): TMEM -> registers -> NCHW global (lanes = consecutive x, coalesced per channel)
    if (warp < CV_PRODUCERS / 32) {
        mbar_wait(&bars[4 + ((nkb - 1) & 1)], ((nkb - 1) >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, half = warp >> 2;                      // TMEM lane quarter, column half
        const int m = q * 32 + lane;                                   // row of the M tile
        const int xo = m % WI, yo = m / WI;
#pragma unroll 1
        for (int t = 0; t < G::MT; ++t) {
            const int y = y0 + t * G::RPT + yo;
#pragma unroll
            for (int cc = 0; cc < NT / 32; ++cc) {
                const int col0 = half * (NT / 2) + cc * 16;
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NT + col0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (y < out.h) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int co = cob * NT + col0 + j;
                        if (co < out.c) {
                            float o = __uint_as_float(v[j]);
                            if (bias) o += __ldg(bias + co);
                            st_stream(out.p + b * out.sb + (int64_t)co * out.sc + y * out.sh + xo * out.sw, o);
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(G::TMEM_COLS) : "memory");
}

template <int WI, int NT, bool BF>
static int launch_conv3x3_m(const View<const float>& xv, const float* packed, const float* bias, const View<float>& ov, cudaStream_t st) {
    using G = CvGeo<WI, NT>;
    const int ncob = ceil_div(ov.c, NT), nkb = ceil_div(xv.c, CvMath<BF>::KB);
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<WI, NT, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) { set_error("conv3x3_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(ceil_div(ov.h, G::ROWS), ncob, ov.n);
    conv3x3_tc_kernel<WI, NT, BF><<<grid, CV_PRODUCERS + 32, G::SMEM, st>>>(xv, packed, bias, ov, nkb);
    return FFWM_OK;
}
template <int WI, int NT = 64>
static int launch_conv3x3(const View<const float>& xv, const float* packed, const float* bias, const View<float>& ov, int math, cudaStream_t st) {
    return math ? launch_conv3x3_m<WI, NT, true>(xv, packed, bias, ov, st) : launch_conv3x3_m<WI, NT, false>(xv, packed, bias, ov, st);
}

}  // namespace ffwm

static bool cv_nt_ok(int nt) { return nt == 64 || nt == 128; }
static bool cv_math_ok(int math) { return math == FFWM_MATH_TF32X3 || math == FFWM_MATH_BF16X3; }

extern "C" int64_t ffwm_conv3x3_packed_floats(int cout, int cin, int nt, int math) {
    if (cout <= 0 || cin <= 0 || !cv_nt_ok(nt) || !cv_math_ok(math)) return 0;
    const int kbs = math ? 16 : 8;                                            // channels per K block of the operand math
    const int64_t ncob = (cout + nt - 1) / nt, nkb = (cin + kbs - 1) / kbs;
    return ncob * nkb * (2 * 9 * 2 * nt * 4);                                 // CvGeo::B_STAGE / 4 floats per (cob, kb)
}

extern "C" int ffwm_conv3x3_pack_weights(const ffwm_tensor4* weight, int dgrad, float* packed, int64_t packed_floats, int nt, int math, void* stream) {
    using namespace ffwm;
    if (!weight || !weight->data || !packed) { set_error("conv3x3_pack_weights: null pointer"); return FFWM_ERR_NULL; }
    if (weight->size[2] != 3 || weight->size[3] != 3) { set_error("conv3x3_pack_weights: kernel must be 3x3"); return FFWM_ERR_SHAPE; }
    if (!cv_nt_ok(nt)) { set_error("conv3x3_pack_weights: nt must be 64 or 128 (got %d)", nt); return FFWM_ERR_ARG; }
    if (!cv_math_ok(math)) { set_error("conv3x3_pack_weights: math must be 0 (3xTF32) or 1 (3xBF16)"); return FFWM_ERR_ARG; }
    const int cout = (int)weight->size[0], cin = (int)weight->size[1];
    const int64_t need = dgrad ? ffwm_conv3x3_packed_floats(cin, cout, nt, math) : ffwm_conv3x3_packed_floats(cout, cin, nt, math);
    if (packed_floats < need) { set_error("conv3x3_pack_weights: packed buffer too small (%lld < %lld floats)", (long long)packed_floats, (long long)need); return FFWM_ERR_SHAPE; }
    const int64_t pairs = need / 2;
    const int blocks = (int)std::min<int64_t>((pairs + 255) / 256, 4096);
    const float* wp = static_cast<const float*>(weight->data);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t s0 = weight->stride[0], s1 = weight->stride[1], s2 = weight->stride[2], s3 = weight->stride[3];
    if (math) {
        __nv_bfloat16* pb = reinterpret_cast<__nv_bfloat16*>(packed);
        if (nt == 64) conv3x3_pack_bf_kernel<64><<<blocks, 256, 0, st>>>(wp, pb, cout, cin, dgrad, s0, s1, s2, s3);
        else conv3x3_pack_bf_kernel<128><<<blocks, 256, 0, st>>>(wp, pb, cout, cin, dgrad, s0, s1, s2, s3);
    } else {
        if (nt == 64) conv3x3_pack_kernel<64><<<blocks, 256, 0, st>>>(wp, packed, cout, cin, dgrad, s0, s1, s2, s3);
        else conv3x3_pack_kernel<128><<<blocks, 256, 0, st>>>(wp, packed, cout, cin, dgrad, s0, s1, s2, s3);
    }
    return check_launch("conv3x3_pack_weights");
}

// conv2d(x, w, bias, stride 1, padding 1) for 3x3 kernels with `packed` = pack_weights(w): x (B,Cin,H,W) fp32,
// out (B,Cout,H,W) fp32, W in {128, 64, 32, 16}.  Replaces the cuDNN call behind nn.Conv2d(…, 3, 1, 1) for those shapes.
extern "C" int ffwm_conv3x3_forward(const ffwm_tensor4* x, const float* packed, const float* bias, const ffwm_tensor4* out, int nt, int math, void* stream) {
    using namespace ffwm;
    View<const float> xv;
    View<float> ov;
    int rc;
    if ((rc = make_view<const float>(x, "x", &xv))) return rc;
    if ((rc = make_view<float>(out, "out", &ov))) return rc;
    if (!packed) { set_error("conv3x3_forward: null packed weights"); return FFWM_ERR_NULL; }
    if ((xv.w != 128 && xv.w != 64 && xv.w != 32 && xv.w != 16) || ov.w != xv.w || xv.h != ov.h || xv.n != ov.n) {
        set_error("conv3x3_forward: needs W in {128,64,32,16} and equal N,H,W (x %dx%dx%dx%d, out %dx%dx%dx%d)", xv.n, xv.c, xv.h, xv.w, ov.n, ov.c, ov.h, ov.w);
        return FFWM_ERR_SHAPE;
    }
    if (nt != 64 && !(nt == 128 && xv.w == 128)) { set_error("conv3x3_forward: nt must be 64, or 128 with W = 128 (got nt %d, W %d)", nt, xv.w); return FFWM_ERR_ARG; }
    if (!cv_math_ok(math)) { set_error("conv3x3_forward: math must be 0 (3xTF32) or 1 (3xBF16)"); return FFWM_ERR_ARG; }
    if ((int64_t)ov.n * ov.c * ov.h == 0) return FFWM_OK;
    if (ov.n > 65535 || ceil_div(ov.c, nt) > 65535) { set_error("conv3x3_forward: grid too large"); return FFWM_ERR_TOO_LARGE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = nt == 128    ? launch_conv3x3<128, 128>(xv, packed, bias, ov, math, st)
         : xv.w == 128 ? launch_conv3x3<128>(xv, packed, bias, ov, math, st)
         : xv.w == 64 ? launch_conv3x3<64>(xv, packed, bias, ov, math, st)
         : xv.w == 32 ? launch_conv3x3<32>(xv, packed, bias, ov, math, st)
                      : launch_conv3x3<16>(xv, packed, bias, ov, math, st);
    if (rc) return rc;
    return check_launch("conv3x3_forward");
}
