// conv_gen_tc: every OTHER dense convolution of the path on the tcgen05 tensor cores — any kernel size up to
// 7x7, stride 1 or 2, any padding, any map size, forward convolution and transposed convolution (= the data
// gradient of a strided convolution = nn.ConvTranspose2d's forward), fp32 in / fp32 out, operand math chosen per
// call like conv3x3_tc.cu's: 3xTF32 (forward passes: library-grade fp32 accuracy, K blocks of 8 channels) or 3xBF16
// (data gradients: twice the tensor rate, K blocks of 16 channels); the text below describes 3xBF16.
//
// Where it is used (everything conv3x3_tc.cu's 3x3 / stride-1 / W in {128,64,32} tiles do
// not cover): FlowNet's stride-2 3x3 encoder, its 3x3 layers on maps of 16x16 and below, its 4x4 stride-2
// transposed convolutions and flow upsamplers (models/base_networks.py:30-57,59-165); the generator's 7x7 stem,
// 4x4 stride-2 encoders and 1x1 residual inputs (:208-246,274-312); MSDiscriminator's stride-2 3x3 and 1x1 layers
// (:354-437); LightCNN's 5x5 stem and 1x1 / small-map 3x3 MFM layers (lightcnn/light_cnn.py:13-26); VGG19 blocks 4-5
// — and the data gradient of each.  The reference runs all of them on cuDNN.
//
// One implicit GEMM, D[pixel, co] += A[pixel, ci] * B[co, ci] per (tap, block of 16 input channels):
//   * M = 128 output pixels of ONE output class, linearised over (image, row, column).  A forward convolution has one
//     class (all pixels).  A transposed convolution with stride s has s*s classes — output pixels with the same
//     (y mod s, x mod s) see the same subset of taps, (ky, kx) = (y + pad, x + pad) mod s, and inside a class the
//     operation is a stride-1 correlation over the class grid — so no MMA ever multiplies by a structural zero.
//   * A is GATHERED: producer thread (pixel, k-chunk) loads its 8 channels at the tap's input position (zero outside
//     the image), splits them into two bf16 parts and writes one 16-byte slot per part: K-major / no-swizzle tiles
//     [part][k-chunk][128 pixels][8 channels], conflict-free 128-bit stores.  Loads of a warp are 32 consecutive
//     pixels of one channel plane (stride 1) or every other pixel (stride 2).
//   * B (weights) is packed once per weight update into the shared-memory image, [unit = (k-block, tap)][part][k-chunk]
//     [NT co][8 ci], and pulled in with ONE bulk async copy per stage (mbarrier complete_tx).
//   * N = NT output channels, a runtime multiple of 16 up to 256 chosen per layer (195 -> 208, 384 -> 2 x 192,
//     2 -> 16): the instruction descriptor, the TMEM allocation and the shared-memory carve-up are runtime values.
//   * a stage = U = 3 or 4 consecutive units of the flattened (k-block, tap) sequence, ring of 2-4 stages, producers
//     (16 warps) / issuer (1 thread) / epilogue (8 of the producer warps) synchronised by mbarriers as in conv3x3_tc.cu.
//   * layers whose M x N tiles do not fill the GPU split the K sequence over several CTAs (blockIdx.z); partial sums
//     meet as fp32 REDs in the (then zero-filled here) output, bias added by split 0.
#include <cuda_bf16.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace ffwm {

constexpr int GN_PRODUCERS = 512;                // 16 producer warps: thread = (pixel, k-chunk, unit parity); 4 warps per scheduler hide the gather latency
constexpr int GN_MAXTAPS = 49;
constexpr int GN_MAXU = 4;                       // units per stage
constexpr int GN_A_UNIT = 2 * 2 * 128 * 16;      // [part][k-chunk][128 pixels][16 B]
constexpr int GN_SMEM_LIMIT = 227 * 1024 - 256;

struct GnTap { int8_t dy, dx; uint8_t ky, kx; };
__device__ const float gn_zero[4] = {0.f, 0.f, 0.f, 0.f};   // where the loads of an out-of-image tap are pointed
struct GnClass {
    int oy0, ox0;         // first output row / column of the class
    int hc, wc;           // class grid
    int tap0, ntaps;      // its taps in GnGeo::taps
    int tiles;            // ceil(n * hc * wc / 128)
    int pad_;
    int64_t packed_off;   // byte offset of the class's weight image
};
struct GnGeo {
    int n, cin, cout, hi, wi, ho, wo;
    int is, os;           // input step per class-grid step (conv: stride), output step (transposed: stride)
    int nt, ncob, nkb;    // N tile, N tiles, K blocks of kbs input channels
    int kbs;              // 16 (3xBF16) or 8 (3xTF32)
    int upst, nstage;     // units per stage, ring depth
    int nclass, splits;
    int b_unit, stage_bytes, tmem_cols;
    GnClass cls[4];
    GnTap taps[GN_MAXTAPS];
};

__host__ __device__ inline int gn_pick_nt(int n_out) {
    const int ntiles = (n_out + 255) / 256;
    return (((n_out + ntiles - 1) / ntiles) + 15) / 16 * 16;
}

// Fills classes and taps.  transposed = 0: out = conv(x, stride, pad); 1: out = conv_transpose(x, stride, pad).
// Returns false if a class has no tap (kernel smaller than the stride): the caller rejects the shape.
static bool gn_build(GnGeo& g, int kh, int kw, int stride, int pad, int transposed, int math) {
    int nt = 0;
    if (!transposed) {
        g.is = stride, g.os = 1, g.nclass = 1;
        GnClass& c = g.cls[0];
        c.oy0 = c.ox0 = 0, c.hc = g.ho, c.wc = g.wo, c.tap0 = 0;
        for (int ky = 0; ky < kh; ++ky)
            for (int kx = 0; kx < kw; ++kx) g.taps[nt++] = GnTap{(int8_t)(ky - pad), (int8_t)(kx - pad), (uint8_t)ky, (uint8_t)kx};
        c.ntaps = nt;
    } else {
        g.is = 1, g.os = stride, g.nclass = stride * stride;
        for (int ry = 0; ry < stride; ++ry)
            for (int rx = 0; rx < stride; ++rx) {
                GnClass& c = g.cls[ry * stride + rx];
                c.oy0 = ry, c.ox0 = rx;
                c.hc = ry < g.ho ? (g.ho - ry + stride - 1) / stride : 0;
                c.wc = rx < g.wo ? (g.wo - rx + stride - 1) / stride : 0;
                c.tap0 = nt;
                for (int ky = (ry + pad) % stride; ky < kh; ky += stride)
                    for (int kx = (rx + pad) % stride; kx < kw; kx += stride)
                        g.taps[nt++] = GnTap{(int8_t)((ry + pad - ky) / stride), (int8_t)((rx + pad - kx) / stride), (uint8_t)ky, (uint8_t)kx};
                c.ntaps = nt - c.tap0;
                if (c.ntaps == 0) return false;
            }
    }
    g.nt = gn_pick_nt(g.cout);
    g.ncob = (g.cout + g.nt - 1) / g.nt;
    g.kbs = math ? 16 : 8;
    g.nkb = (g.cin + g.kbs - 1) / g.kbs;
    g.b_unit = g.nt * 64;                                            // [part][k-chunk][nt][16 B]
    g.upst = g.nt <= 128 ? 4 : 3;
    g.stage_bytes = g.upst * (GN_A_UNIT + g.b_unit);
    g.nstage = std::max(2, std::min(4, GN_SMEM_LIMIT / g.stage_bytes));
    g.tmem_cols = 32;
    while (g.tmem_cols < g.nt) g.tmem_cols *= 2;
    int64_t off = 0;
    for (int i = 0; i < g.nclass; ++i) {
        GnClass& c = g.cls[i];
        c.tiles = (int)(((int64_t)g.n * c.hc * c.wc + 127) / 128);
        c.packed_off = off;
        off += (int64_t)g.ncob * g.nkb * c.ntaps * g.b_unit;
    }
    return true;
}

// ---------------------------------------------------------------- weight packing
// packed[class][cob][unit = kb * ntaps + t][part][kchunk][co_local nt][8 ci] (bf16), B[o][c] = w[o*s_o + c*s_c + ky*s_ky + kx*s_kx]
// (3xTF32: the same with 4 fp32 channels per 16-byte slot and hi / lo parts)
// One CTA packs an (o_tile output channels) x (one K block of input channels) x (all kh*kw taps) tile: it reads the tile with
// the weight's contiguous axis innermost (coalesced whichever of the first two dimensions is the input channel), keeps it in
// shared memory, and writes 16-byte slots in runs of o_tile * 16 contiguous bytes per (class, tap, part, k-chunk).
template <bool BF>
__global__ void __launch_bounds__(256) conv_gen_pack_kernel(const float* __restrict__ w, unsigned char* __restrict__ packed, int64_t s_o,
                                                            int64_t s_c, int64_t s_ky, int64_t s_kx, const __grid_constant__ GnGeo g, int o_shift,
                                                            int kh, int kw) {
    constexpr int CPS = BF ? 8 : 4, KB = 2 * CPS, KB_SHIFT = BF ? 4 : 3;   // channels per slot, per K block
    extern __shared__ float gn_pack_smem[];                            // [o_tile][KB * T + 1] floats, then T tap offsets
    const int T = kh * kw, RS = KB * T + 1, o_tile = 1 << o_shift;
    int* toff = reinterpret_cast<int*>(gn_pack_smem + o_tile * RS);
    const int ogroups = g.nt >> o_shift;
    const int og = blockIdx.x % ogroups, kb = (blockIdx.x / ogroups) % g.nkb, cob = blockIdx.x / (ogroups * g.nkb);
    const int o0 = cob * g.nt + og * o_tile, ch0 = kb * KB;
    for (int t = threadIdx.x; t < T; t += blockDim.x) toff[t] = (int)((t / kw) * s_ky + (t % kw) * s_kx);
    __syncthreads();
    // (row r, tap) pairs in memory order, walked without divisions: the index math, not the memory system, bounded the
    // first version of this kernel (profiles/r02v_pack_reduce.txt)
    const bool o_outer = s_c <= s_o;                                   // the input channels of one output channel lie closer together
    const int step_tap = (int)blockDim.x % T, step_r = (int)blockDim.x / T, rows = o_tile * KB;
    int tap = (int)threadIdx.x % T, r = (int)threadIdx.x / T;
#pragma unroll 4
    for (; r < rows;) {
        const int ch_l = o_outer ? (r & (KB - 1)) : (r >> o_shift), o_l = o_outer ? (r >> KB_SHIFT) : (r & (o_tile - 1));
        const int o = o0 + o_l, ch = ch0 + ch_l;
        float v = 0.f;
        if (o < g.cout && ch < g.cin) v = __ldg(w + (o * s_o + ch * s_c) + toff[tap]);
        gn_pack_smem[o_l * RS + ch_l * T + tap] = v;
        tap += step_tap;
        r += step_r;
        if (tap >= T) tap -= T, ++r;
    }
    __syncthreads();
    for (int ci = 0; ci < g.nclass; ++ci) {
        const GnClass& c = g.cls[ci];
        unsigned char* base = packed + c.packed_off + ((int64_t)cob * g.nkb + kb) * c.ntaps * (int64_t)g.b_unit + (og * o_tile) * 16;
#pragma unroll 2
        for (int i = threadIdx.x; i < ((c.ntaps * 4) << o_shift); i += blockDim.x) {
            const int o_l = i & (o_tile - 1), kc = (i >> o_shift) & 1, part = (i >> (o_shift + 1)) & 1, t = i >> (o_shift + 2);
            const GnTap tp = g.taps[c.tap0 + t];
            const float* src = gn_pack_smem + o_l * RS + kc * CPS * T + tp.ky * kw + tp.kx;
            float v[CPS];
#pragma unroll
            for (int j = 0; j < CPS; ++j) v[j] = src[j * T];
            uint32_t h[4];
            if (BF) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat16 e0 = __float2bfloat16_rn(v[2 * j]), e1 = __float2bfloat16_rn(v[2 * j + 1]);
                    if (part) {
                        e0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(e0));
                        e1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(e1));
                    }
                    h[j] = (uint32_t)__bfloat16_as_ushort(e0) | ((uint32_t)__bfloat16_as_ushort(e1) << 16);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float hi = __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u);
                    h[j] = __float_as_uint(part ? v[j] - hi : hi);
                }
            }
            *reinterpret_cast<uint4*>(base + t * g.b_unit + part * (2 * g.nt * 16) + (kc * g.nt + o_l) * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        }
    }
}

// ---------------------------------------------------------------- the convolution
// OCC = CTAs per SM the kernel is compiled for: 2 caps the registers at 56 and is launched with <= 113 KB of shared memory
// and <= 256 TMEM columns, so that two CTAs share an SM and one's load -> split -> store -> MMA latency chain hides behind
// the other's (the small-map layers are latency-bound: ncu shows 62 % of cycles without an eligible warp at one CTA per SM).
template <bool BF, int OCC>
__global__ void __launch_bounds__(GN_PRODUCERS + 32, OCC)
conv_gen_tc_kernel(View<const float> x, const unsigned char* __restrict__ packed, const float* __restrict__ bias, View<float> out,
                   const __grid_constant__ GnGeo g) {
    constexpr int CPS = BF ? 8 : 4;                                    // channels per 16-byte slot
    extern __shared__ __align__(128) unsigned char gn_smem[];
    const int cls_i = blockIdx.z % g.nclass, split = blockIdx.z / g.nclass;
    const GnClass& c = g.cls[cls_i];
    if ((int)blockIdx.x >= c.tiles) return;                            // CTA-uniform, before any barrier / allocation
    const int units = g.nkb * c.ntaps;
    const int stages_total = (units + g.upst - 1) / g.upst;
    const int per = (stages_total + g.splits - 1) / g.splits;
    const int s0 = split * per, nst = min(per, stages_total - s0);
    if (nst <= 0) return;

    uint64_t* bars = reinterpret_cast<uint64_t*>(gn_smem + g.nstage * g.stage_bytes);   // full[0..3] empty[4..7]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cob = blockIdx.y;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[i], GN_PRODUCERS);
            mbar_init(&bars[4 + i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(g.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    const int64_t npix = (int64_t)g.n * c.hc * c.wc;
    const int a_stage = g.upst * GN_A_UNIT;

    if (warp < GN_PRODUCERS / 32) {
        // ================= producers: gather the activations of the stage's units =================
        const int m = tid & 127, kc = (tid >> 7) & 1, uh = tid >> 8;       // this thread stages the units u with (u & 1) == uh
        const int64_t p = (int64_t)blockIdx.x * 128 + m;
        const bool pv = p < npix;
        int a = 0, b = 0, img = 0;
        if (pv) { b = (int)(p % c.wc); const int64_t q = p / c.wc; a = (int)(q % c.hc); img = (int)(q / c.hc); }
        const int ya = a * g.is, xb = b * g.is;
        const float* base = x.p + img * x.sb + (int64_t)ya * x.sh + (int64_t)xb * x.sw;
        const unsigned char* pk = packed + c.packed_off + (int64_t)cob * units * g.b_unit;
        // The gathers go to REGISTERS, so they do not depend on a free shared-memory slot: the loads of stage k+1 are
        // issued before stage k is split and stored (software pipeline, two register sets) and their L2 / HBM latency
        // hides behind a whole stage period instead of stalling every stage.
        // A load is `base + j * stride` with an immediate j; an out-of-image tap (or a pixel past the end) turns base into a
        // zero word and the stride into 0 instead of predicating every load; only the last K block of a tensor whose
        // channel count is not a multiple of the slot width takes the predicated path (warp-uniform).
        auto load_stage = [&](int k, float (&v)[GN_MAXU / 2][CPS]) {
            const int u0 = (s0 + k) * g.upst, nu = min(g.upst, units - u0);
            int kb = (u0 + uh) / c.ntaps, t = (u0 + uh) - kb * c.ntaps;
#pragma unroll
            for (int uu = 0; uu < GN_MAXU / 2; ++uu) {
                if (2 * uu + uh < nu) {
                    const GnTap tp = g.taps[c.tap0 + t];
                    const bool ok = pv && (unsigned)(ya + tp.dy) < (unsigned)g.hi && (unsigned)(xb + tp.dx) < (unsigned)g.wi;
                    const int c0 = kb * (2 * CPS) + kc * CPS;
                    const float* gp = ok ? base + (int64_t)tp.dy * x.sh + (int64_t)tp.dx * x.sw + (int64_t)c0 * x.sc : gn_zero;
                    const int64_t str = ok ? x.sc : 0;
                    if (c0 + CPS <= g.cin) {
#pragma unroll
                        for (int j = 0; j < CPS; ++j) v[uu][j] = __ldg(gp + j * str);
                    } else {
#pragma unroll
                        for (int j = 0; j < CPS; ++j) v[uu][j] = c0 + j < g.cin ? __ldg(gp + j * str) : 0.f;
                    }
                    t += 2;                                            // two units further (ntaps may be 1: wrap repeatedly)
                    while (t >= c.ntaps) { t -= c.ntaps; ++kb; }
                }
            }
        };
        auto store_stage = [&](int k, float (&v)[GN_MAXU / 2][CPS]) {
            const int slot = k % g.nstage;
            if (k >= g.nstage) mbar_wait(&bars[4 + slot], ((k / g.nstage) - 1) & 1);    // MMAs that read this slot are done
            const int u0 = (s0 + k) * g.upst, nu = min(g.upst, units - u0);
            unsigned char* st = gn_smem + slot * g.stage_bytes;
            if (tid == 0) {                                            // weights of the stage: one bulk copy
                const uint32_t bar = smem_u32(&bars[slot]), bytes = (uint32_t)(nu * g.b_unit);
                asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(st + a_stage)),
                             "l"(pk + (int64_t)u0 * g.b_unit), "r"(bytes), "r"(bar)
                             : "memory");
            }
#pragma unroll
            for (int uu = 0; uu < GN_MAXU / 2; ++uu)
                if (2 * uu + uh < nu) split_store_m<BF>(st + (2 * uu + uh) * GN_A_UNIT + kc * 2048 + m * 16, 4096, v[uu]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[slot])) : "memory");
        };
        float va[GN_MAXU / 2][CPS], vb[GN_MAXU / 2][CPS];
        load_stage(0, va);
        for (int k = 0; k < nst; k += 2) {
            if (k + 1 < nst) load_stage(k + 1, vb);
            store_stage(k, va);
            if (k + 1 < nst) {
                if (k + 2 < nst) load_stage(k + 2, va);
                store_stage(k + 1, vb);
            }
        }
    } else if (lane == 0) {
        // ================= issuer =================
        const uint32_t idesc = BF ? umma_idesc_bf16(128, g.nt) : umma_idesc_tf32(128, g.nt);
        for (int k = 0; k < nst; ++k) {
            const int slot = k % g.nstage;
            mbar_wait(&bars[slot], (k / g.nstage) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int u0 = (s0 + k) * g.upst, nu = min(g.upst, units - u0);
            const uint32_t sA = smem_u32(gn_smem + slot * g.stage_bytes), sB = sA + a_stage;
            for (int u = 0; u < nu; ++u) {
                const uint64_t dA1 = umma_desc(sA + u * GN_A_UNIT, 2048, 128), dA2 = dA1 + (4096 >> 4);
                const uint64_t dB1 = umma_desc(sB + u * g.b_unit, g.nt * 16, 128), dB2 = dB1 + (uint64_t)((g.nt * 32) >> 4);
                umma_ss<BF>(tmem, dA1, dB1, idesc, k > 0 || u > 0);
                umma_ss<BF>(tmem, dA1, dB2, idesc, true);
                umma_ss<BF>(tmem, dA2, dB1, idesc, true);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[4 + slot])) : "memory");
        }
    }

    // ---- epilogue (8 of the producer warps): TMEM -> registers -> NCHW global (lanes = consecutive pixels of the class)
    if (warp < 8) {
        const int last = nst - 1;
        mbar_wait(&bars[4 + last % g.nstage], (last / g.nstage) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, half = warp >> 2;
        const int64_t p = (int64_t)blockIdx.x * 128 + q * 32 + lane;
        const bool pv = p < npix;
        float* op = out.p;
        if (pv) {
            const int b = (int)(p % c.wc);
            const int64_t r = p / c.wc;
            const int a = (int)(r % c.hc), img = (int)(r / c.hc);
            op += img * out.sb + (int64_t)(c.oy0 + a * g.os) * out.sh + (int64_t)(c.ox0 + b * g.os) * out.sw;
        }
        const bool add_bias = bias != nullptr && split == 0;
        for (int col0 = half * 16; col0 < g.nt; col0 += 32) {          // warp-uniform bounds
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (pv) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int co = cob * g.nt + col0 + j;
                    if (co < g.cout) {
                        float o = __uint_as_float(v[j]);
                        if (add_bias) o += __ldg(bias + co);
                        if (g.splits > 1) red_add(op + (int64_t)co * out.sc, o);
                        else op[(int64_t)co * out.sc] = o;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(g.tmem_cols) : "memory");
}

static int gn_setup(GnGeo& g, const char* what, int n, int cin, int cout, int hi, int wi, int ho, int wo, int kh, int kw, int stride,
                    int pad, int transposed, int math) {
    if (math != 0 && math != 1) { set_error("%s: math must be 0 (3xTF32) or 1 (3xBF16)", what); return FFWM_ERR_ARG; }
    if (kh < 1 || kw < 1 || kh > 7 || kw > 7 || kh * kw > GN_MAXTAPS || (stride != 1 && stride != 2) || pad < 0 || pad > 7) {
        set_error("%s: kernel %dx%d stride %d padding %d is outside 1..7 / {1,2} / 0..7", what, kh, kw, stride, pad);
        return FFWM_ERR_ARG;
    }
    const int eh = transposed ? (hi - 1) * stride - 2 * pad + kh : (hi + 2 * pad - kh) / stride + 1;
    const int ew = transposed ? (wi - 1) * stride - 2 * pad + kw : (wi + 2 * pad - kw) / stride + 1;
    if (ho < 1 || wo < 1 || (transposed ? (ho < eh || ho >= eh + stride || wo < ew || wo >= ew + stride) : (ho != eh || wo != ew))) {
        set_error("%s: output %dx%d does not match input %dx%d, kernel %dx%d, stride %d, padding %d (%s)", what, ho, wo, hi, wi, kh, kw,
                  stride, pad, transposed ? "transposed" : "convolution");
        return FFWM_ERR_SHAPE;
    }
    g.n = n, g.cin = cin, g.cout = cout, g.hi = hi, g.wi = wi, g.ho = ho, g.wo = wo;
    if (!gn_build(g, kh, kw, stride, pad, transposed, math)) {
        set_error("%s: a %dx%d kernel with stride %d leaves output classes without any tap", what, kh, kw, stride);
        return FFWM_ERR_ARG;
    }
    return FFWM_OK;
}

}  // namespace ffwm

// Bytes of the packed image of a weight with n_out output and n_in input channels and kh x kw taps.
extern "C" int64_t ffwm_conv_packed_bytes(int n_out, int n_in, int kh, int kw, int math) {
    if (n_out <= 0 || n_in <= 0 || kh < 1 || kw < 1 || kh * kw > ffwm::GN_MAXTAPS || (math != 0 && math != 1)) return 0;
    const int nt = ffwm::gn_pick_nt(n_out), kbs = math ? 16 : 8;
    return (int64_t)((n_out + nt - 1) / nt) * ((n_in + kbs - 1) / kbs) * kh * kw * nt * 64;
}

// weight -> packed.  The weight is described by its four dimensions as they are in memory; `in_major` says which of
// the first two is the INPUT channel of the operation being packed:
//   in_major = 0  weight[n_out][n_in][kh][kw]   (nn.Conv2d forward; the data gradient of nn.ConvTranspose2d)
//   in_major = 1  weight[n_in][n_out][kh][kw]   (nn.ConvTranspose2d forward; the data gradient of nn.Conv2d)
// stride / pad / transposed must be those of the ffwm_conv_forward call that will consume the image.
extern "C" int ffwm_conv_pack_weights(const ffwm_tensor4* weight, int in_major, int stride, int pad, int transposed, int math,
                                      void* packed, int64_t packed_bytes, void* stream) {
    using namespace ffwm;
    if (!weight || !weight->data || !packed) { set_error("conv_pack_weights: null pointer"); return FFWM_ERR_NULL; }
    const int kh = (int)weight->size[2], kw = (int)weight->size[3];
    const int n_out = (int)weight->size[in_major ? 1 : 0], n_in = (int)weight->size[in_major ? 0 : 1];
    GnGeo g;
    // the class / tap structure does not depend on the map size; any consistent size works for packing
    const int hi = 8, ho = transposed ? (hi - 1) * stride - 2 * pad + kh : (hi + 2 * pad - kh) / stride + 1;
    if (ho < 1) { set_error("conv_pack_weights: bad geometry"); return FFWM_ERR_ARG; }
    int rc = gn_setup(g, "conv_pack_weights", 1, n_in, n_out, hi, hi, ho, ho, kh, kw, stride, pad, transposed, math);
    if (rc) return rc;
    const int64_t need = ffwm_conv_packed_bytes(n_out, n_in, kh, kw, math);
    if (packed_bytes < need) { set_error("conv_pack_weights: packed buffer too small (%lld < %lld bytes)", (long long)packed_bytes, (long long)need); return FFWM_ERR_SHAPE; }
    const int64_t s_o = weight->stride[in_major ? 1 : 0], s_c = weight->stride[in_major ? 0 : 1];
    const int T = kh * kw, kbs = math ? 16 : 8;
    int o_tile = g.nt % 32 == 0 ? 32 : 16;
    while (o_tile > 8 && (int64_t)o_tile * (kbs * T + 1) * 4 > 40 * 1024) o_tile /= 2;
    while (o_tile > 8 && (int64_t)g.ncob * g.nkb * (g.nt / o_tile) < 2 * sm_count()) o_tile /= 2;   // small weights: more, shorter CTAs
    const int o_shift = o_tile == 32 ? 5 : o_tile == 16 ? 4 : 3;
    const int64_t ctas = (int64_t)g.ncob * g.nkb * (g.nt / o_tile);
    if (ctas > 0x7fffffffLL) { set_error("conv_pack_weights: weight too large"); return FFWM_ERR_TOO_LARGE; }
    auto pk = math ? conv_gen_pack_kernel<true> : conv_gen_pack_kernel<false>;
    pk<<<(unsigned)ctas, 256, ((size_t)o_tile * (kbs * T + 1) + T) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float*>(weight->data), static_cast<unsigned char*>(packed), s_o, s_c, weight->stride[2], weight->stride[3], g, o_shift, kh, kw);
    return check_launch("conv_pack_weights");
}

// transposed = 0: out (B,Cout,Ho,Wo) = conv2d(x (B,Cin,Hi,Wi), W, bias, stride, pad)           [replaces cuDNN behind nn.Conv2d]
// transposed = 1: out = conv_transpose2d(x, W, bias, stride, pad), Ho in [(Hi-1)s-2p+kh, +s)   [nn.ConvTranspose2d, and the
//                 data gradient of a strided nn.Conv2d: x = grad_output, out = grad_input of its size]
// fp32, any strides; `packed` from ffwm_conv_pack_weights with the same kh, kw, stride, pad, transposed, math.  bias may be NULL.
// math: 0 = 3xTF32 operand split, 1 = 3xBF16 operand split.
extern "C" int ffwm_conv_forward(const ffwm_tensor4* x, const void* packed, const float* bias, const ffwm_tensor4* out, int kh, int kw,
                                 int stride, int pad, int transposed, int math, void* stream) {
    using namespace ffwm;
    View<const float> xv;
    View<float> ov;
    int rc;
    if ((rc = make_view<const float>(x, "x", &xv))) return rc;
    if ((rc = make_view<float>(out, "out", &ov))) return rc;
    if (!packed) { set_error("conv_forward: null packed weights"); return FFWM_ERR_NULL; }
    if (xv.n != ov.n) { set_error("conv_forward: batch mismatch (%d vs %d)", xv.n, ov.n); return FFWM_ERR_SHAPE; }
    if ((int64_t)ov.n * ov.c * ov.h * ov.w == 0) return FFWM_OK;
    if (xv.c == 0) { set_error("conv_forward: no input channels"); return FFWM_ERR_SHAPE; }
    GnGeo g;
    if ((rc = gn_setup(g, "conv_forward", xv.n, xv.c, ov.c, xv.h, xv.w, ov.h, ov.w, kh, kw, stride, pad, transposed, math))) return rc;
    int64_t ctas = 0;
    int max_tiles = 0, min_units = 1 << 30;
    for (int i = 0; i < g.nclass; ++i) {
        ctas += (int64_t)g.cls[i].tiles * g.ncob;
        max_tiles = std::max(max_tiles, g.cls[i].tiles);
        min_units = std::min(min_units, g.nkb * g.cls[i].ntaps);
    }
    if (max_tiles == 0) return FFWM_OK;
    // two CTAs per SM (see the kernel), measured per shape (profiles/r02y_conv_gen_occ.txt): a gain where many tiles queue per
    // SM (one CTA's epilogue overlaps the other's main loop: 1x1 195->195 @128 0.136 -> 0.095 ms) and where one or two tiles
    // stream a large weight (1024->1024 @2x2: 0.052 -> 0.033 ms); 5 % slower in between (shorter stages, two-deep ring)
    const int occ_opt = opt(OPT_CONV_OCC2);                            // 0: automatic, 1: never, 2: always
    const bool occ2 = occ_opt == 2 || (occ_opt == 0 && (ctas >= 2 * (int64_t)sm_count() || max_tiles <= 2));
    if (occ2) {
        g.upst = std::max(1, std::min(GN_MAXU, (56 * 1024) / (GN_A_UNIT + g.b_unit)));
        g.stage_bytes = g.upst * (GN_A_UNIT + g.b_unit);
        g.nstage = 2;
    }
    // split the K sequence while the tiles alone leave SMs idle; every split keeps >= 2 stages of the shortest class
    const int min_stages = (min_units + g.upst - 1) / g.upst;
    const int64_t slots = (int64_t)sm_count() * (occ2 ? 2 : 1);
    g.splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(slots / std::max<int64_t>(ctas, 1), min_stages / 2), 64));
    if ((int64_t)g.nclass * g.splits > 65535 || g.ncob > 65535) { set_error("conv_forward: grid too large"); return FFWM_ERR_TOO_LARGE; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g.splits > 1) {
        // partial sums arrive as REDs: the output starts from zero (dense tensors only: checked through the strides)
        const bool dense = ov.sw == 1 && ov.sh == ov.w && ov.sc == (int64_t)ov.h * ov.w && ov.sb == ov.sc * ov.c;
        if (!dense) g.splits = 1;
        else {
            cudaError_t e = cudaMemsetAsync(ov.p, 0, sizeof(float) * (size_t)ov.n * ov.c * ov.h * ov.w, st);
            if (e != cudaSuccess) { set_error("conv_forward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return int(e); }
        }
    }
    const int smem = g.nstage * g.stage_bytes + 128;
    auto kern = occ2 ? (math ? conv_gen_tc_kernel<true, 2> : conv_gen_tc_kernel<false, 2>) : (math ? conv_gen_tc_kernel<true, 1> : conv_gen_tc_kernel<false, 1>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_forward: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    dim3 grid(max_tiles, g.ncob, g.nclass * g.splits);
    kern<<<grid, GN_PRODUCERS + 32, smem, st>>>(xv, static_cast<const unsigned char*>(packed), bias, ov, g);
    return check_launch("conv_forward");
}
