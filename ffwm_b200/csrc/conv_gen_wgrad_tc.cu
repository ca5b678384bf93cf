// conv_gen_wgrad_tc: weight gradient of ANY dense convolution of the path on the tcgen05 tensor cores (kernels up
// to 7x7, stride 1 or 2, any padding / map size, nn.Conv2d and nn.ConvTranspose2d), fp32 in / fp32 out, 3xBF16 split.
//
//   dW[a][b][ky][kx] = sum over (n, y, x) of  S[n, a, y, x] * L[n, b, y*s - p + ky, x*s - p + kx]        (zero padding)
//
// S is the tensor on the SMALL side of the strided operation and L the one on the large side: for nn.Conv2d S = grad_out,
// L = the input (a = Cout, b = Cin); for nn.ConvTranspose2d S = the input, L = grad_out (a = Cin, b = Cout) — in both cases
// (a, b) are the first two dimensions of the weight as it lies in memory.  Replaces aten::convolution_backward's
// grad_weight (cuDNN wgrad engines) behind models/base_networks.py:30-57,208-246,274-312,354-437, lightcnn/light_cnn.py:13-26.
//
// The contraction runs over PIXELS.  Instead of transposing anything, both operands are staged exactly like the
// activations of conv_gen_tc.cu — thread = (pixel, group of 8 channels): 8 coalesced loads (lanes = 32 consecutive
// pixels of a channel plane), one 16-byte slot of 8 channels per bf16 part — and the tensor core reads those tiles as
// MN-MAJOR operands (instruction-descriptor bits a_major = b_major = 1): in the no-swizzle canonical layout a 16-byte
// slot is 8 consecutive M (or N) indices of one K index, 8 consecutive K indices are 128 contiguous bytes (LBO = 128 B
// between K groups of 8), and the next 8 M/N indices follow SBO = 16 B x (pixels per stage) later.  K = 16 pixels per MMA.
//   * M = 128 channels `a` of S;  N = TG x NB = (tap group) x (channels b of L, a multiple of 8), tap-major, up to 256:
//     rows of the B tile are L gathered at the tap's shifted position, so L is read once per tap through L1 / L2.
//   * a stage = 64 pixels of the linearised (n, y, x) range of S (4 K steps x 3 MMAs of the split), ring of 2-3 stages,
//     16 producer warps / issuer / epilogue (8 of the producer warps) on mbarriers as in the other tcgen05 kernels.
//   * split K: the pixel range is cut over blockIdx.y (at most 64 stages per CTA: bounds the truncating fp32 accumulation
//     chain to 768 updates).  Partial tiles go to a caller-provided workspace with plain 64-byte-per-lane stores and a
//     second kernel sums the splits into dW in a fixed order: deterministic, no atomics, no zero-fill contract.
#include <cuda_bf16.h>
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace ffwm {

constexpr int WGG_PRODUCERS = 512;                // 16 producer warps (4 per scheduler: the gathers are latency- and issue-bound)
constexpr int WGG_GQ = WGG_PRODUCERS / 64;         // group quarters: thread -> (pixel of the stage, first group)
__device__ const float wgg_zero[4] = {0.f, 0.f, 0.f, 0.f};   // where the loads of an out-of-image tap are pointed
constexpr int WGG_KP = 64;                         // pixels per stage
constexpr int WGG_GROUP = WGG_KP * 16;             // bytes of one 8-channel group of a tile part: [64 pixels][16 B]
constexpr int WGG_A_PART = 16 * WGG_GROUP;         // 128 channels
constexpr int WGG_MAX_STAGES = 64;                 // per CTA: the tensor core adds into its fp32 accumulators with truncation, so the error grows with the chain (768 updates: ~1.5e-5 of max|dW| measured)

struct WggGeo {
    int n, hs, ws, hl, wl;       // S (n, ca, hs, ws), L (n, cb, hl, wl)
    int ca, cb, kh, kw, stride, pad;
    int tg, nb, nn;              // taps per tile, channels b per tile, N = tg * nb
    int nat, nbt, ntg, tiles;    // tiles over a, b, tap groups
    int64_t npix;
    int stages_total, stages_per_split, splits;
    int b_part, stage_bytes, nstage, tmem_cols;
    // row mode (stride 1): one B tile per kernel ROW serves all kw taps of that row through descriptor offsets
    int rows_mode;               // 0: every (tap, channel) column is gathered on its own
    int seg, rps, bpr, bp;       // pixels per row segment, rows per stage, B pixels per row (seg + kw - 1), per stage (rps * bpr)
    int kyn;                     // kernel rows per tile (tg = kyn * kw taps)
};

// MN-major, no-swizzle shared-memory descriptor: same fields as umma_desc (umma.cuh); for this layout the "leading"
// offset steps to the next group of 8 K indices and the "stride" offset to the next group of 8 M/N indices.
// (validated on a B200: with the two offsets exchanged every parity case fails)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t k_group_bytes, uint32_t mn_group_bytes) {
    return umma_desc(saddr, k_group_bytes, mn_group_bytes);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int m, int n) {
    return umma_idesc_bf16(m, n) | (1u << 15) | (1u << 16);
}

template <bool ROWS>
__global__ void __launch_bounds__(WGG_PRODUCERS + 32, 1)
conv_gen_wgrad_tc_kernel(View<const float> S, View<const float> L, float* __restrict__ ws, const __grid_constant__ WggGeo g) {
    extern __shared__ __align__(128) unsigned char wg_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(wg_smem + g.nstage * g.stage_bytes);     // full[0..2] empty[4..6]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int tgi = tile % g.ntg, bt = (tile / g.ntg) % g.nbt, at = tile / (g.ntg * g.nbt);
    const int st0 = split * g.stages_per_split;
    const int nst = min(g.stages_per_split, g.stages_total - st0);                        // >= 1 (host)

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&bars[i], WGG_PRODUCERS);
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

    if (warp < WGG_PRODUCERS / 32) {
        // ================= producers =================
        // Everything that does not depend on the stage is formed once: the thread's 2 groups of S channels and its up to
        // 4 groups of the gathered L tile (tap, first channel, valid channel count) — the stage loop is loads, splits and
        // stores only; the pixel coordinates advance incrementally (no 64-bit division per stage).  A load is
        // `base + e * stride` with an immediate e; an out-of-image tap (or a pixel past the end) turns base into a zero
        // word and the stride into 0 instead of predicating eight loads.
        constexpr int NA = 16 / WGG_GQ, NB = 32 / WGG_GQ;                   // groups per thread
        const int px = tid & (WGG_KP - 1), g0 = tid / WGG_KP;              // pixel of the stage, first group (0..7)
        const int nbg = g.nb / 8;                                          // channel groups per tap
        const int ngB = g.nn / 8;                                          // groups of the B tile
        const int T = g.kh * g.kw;
        int64_t a_off[NA];
        int a_cnt[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int a0 = at * 128 + (g0 + WGG_GQ * i) * 8;
            a_off[i] = (int64_t)a0 * S.sc;
            a_cnt[i] = max(0, min(8, g.ca - a0));
        }
        // B items of this thread.  Gather mode: NB groups (tap, 8 channels) at the stage pixel px.  Row mode: up to NBR items
        // (kernel row ky, 8 channels, pixel of the padded row segment): the tile of a kernel row is loaded once, with kw - 1
        // halo pixels, and the kw taps of that row read it at descriptor offsets of one pixel (16 bytes).
        constexpr int NBR = 4;
        int64_t b_off[NB];
        int b_kyx[NB], b_cnt[NB];                                          // ky | kx << 8; valid channels (0: group unused)
        int r_dst[NBR], r_dy[NBR], r_dx[NBR];                              // row mode: smem byte offset, row / column offset from the stage origin
        if constexpr (!ROWS) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int gg = g0 + WGG_GQ * j;
                const int tl = gg / nbg, b0 = bt * g.nb + (gg - tl * nbg) * 8;
                const int tap = tgi * g.tg + tl, ky = tap / g.kw, kx = tap - ky * g.kw;
                const bool ok = gg < ngB && tap < T;
                b_kyx[j] = ky | (kx << 8);
                b_cnt[j] = ok ? max(0, min(8, g.cb - b0)) : 0;
                b_off[j] = (int64_t)b0 * L.sc + (int64_t)ky * L.sh + (int64_t)kx * L.sw;
            }
        } else {
            const int items = g.kyn * nbg * g.bp, ky0 = tgi * g.kyn;
#pragma unroll
            for (int j = 0; j < NBR; ++j) {
                const int it = tid + WGG_PRODUCERS * j;
                const int pxh = it % g.bp, gr = (it / g.bp) % nbg, kyi = it / (g.bp * nbg);
                const int r = pxh / g.bpr, xh = pxh - r * g.bpr;
                const int b0 = bt * g.nb + gr * 8;
                b_cnt[j] = it < items ? max(0, min(8, g.cb - b0)) : -1;    // -1: no item
                b_off[j] = (int64_t)b0 * L.sc;
                r_dy[j] = r + ky0 + kyi - g.pad;
                r_dx[j] = xh - g.pad;
                r_dst[j] = kyi * 2 * g.b_part + gr * (g.bp * 16) + pxh * 16;
            }
        }
        // 8 channels at `q`, `str` elements apart; cnt < 8 only for the last group of a tensor (warp-uniform)
        auto load8 = [](const float* q, int64_t str, int cnt, float (&v)[8]) {
            if (cnt == 8) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __ldg(q + e * str);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = e < cnt ? __ldg(q + e * str) : 0.f;
            }
        };
        // pixel of this thread in the first stage; later stages advance by WGG_KP pixels
        int64_t p = (int64_t)st0 * WGG_KP + px;
        int xs = (int)(p % g.ws), ys = (int)((p / g.ws) % g.hs), img = (int)(p / ((int64_t)g.ws * g.hs));
        for (int k = 0; k < nst; ++k) {
            const bool pv = p < g.npix;
            const float* sp = pv ? S.p + img * S.sb + (int64_t)ys * S.sh + (int64_t)xs * S.sw : wgg_zero;
            const int64_t s_str = pv ? S.sc : 0;
            const int yl0 = ys * g.stride - g.pad, xl0 = xs * g.stride - g.pad;
            const float* lp = L.p + img * L.sb + (int64_t)yl0 * L.sh + (int64_t)xl0 * L.sw;
            float va[NA][8], vb[NB][8];
#pragma unroll
            for (int i = 0; i < NA; ++i) load8(pv ? sp + a_off[i] : wgg_zero, s_str, a_cnt[i], va[i]);
            if constexpr (!ROWS) {
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    if (g0 + WGG_GQ * j < ngB) {                           // uniform over the 64 threads of a group quarter
                        const int ky = b_kyx[j] & 255, kx = b_kyx[j] >> 8;
                        const bool ok = pv && (unsigned)(yl0 + ky) < (unsigned)g.hl && (unsigned)(xl0 + kx) < (unsigned)g.wl;
                        load8(ok ? lp + b_off[j] : wgg_zero, ok ? L.sc : 0, b_cnt[j], vb[j]);
                    }
                }
            } else {
                // stage origin (first pixel of the stage) from this thread's own pixel: stages never straddle a row segment
                const int oy = ys - px / g.seg, ox = xs - px % g.seg;
                const float* lo = L.p + img * L.sb;
#pragma unroll
                for (int j = 0; j < NBR; ++j) {
                    if (b_cnt[j] >= 0) {
                        const int yl = oy + r_dy[j], xl = ox + r_dx[j];
                        const bool ok = (unsigned)yl < (unsigned)g.hl && (unsigned)xl < (unsigned)g.wl;
                        load8(ok ? lo + (int64_t)yl * L.sh + (int64_t)xl * L.sw + b_off[j] : wgg_zero, ok ? L.sc : 0, b_cnt[j], vb[j]);
                    }
                }
            }
            const int slot = k % g.nstage;
            if (k >= g.nstage) mbar_wait(&bars[4 + slot], ((k / g.nstage) - 1) & 1);        // MMAs that read this slot are done
            unsigned char* sA = wg_smem + slot * g.stage_bytes;
            unsigned char* sB = sA + 2 * WGG_A_PART;
#pragma unroll
            for (int i = 0; i < NA; ++i) split_store_bf(sA + (g0 + WGG_GQ * i) * WGG_GROUP + px * 16, WGG_A_PART, va[i]);
            if constexpr (!ROWS) {
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if (g0 + WGG_GQ * j < ngB) split_store_bf(sB + (g0 + WGG_GQ * j) * WGG_GROUP + px * 16, g.b_part, vb[j]);
            } else {
#pragma unroll
                for (int j = 0; j < NBR; ++j)
                    if (b_cnt[j] >= 0) split_store_bf(sB + r_dst[j], g.b_part, vb[j]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[slot])) : "memory");
            // next stage: WGG_KP pixels further along (n, y, x)
            p += WGG_KP;
            xs += WGG_KP;
            if (xs >= g.ws) {
                const int q = xs / g.ws;
                xs -= q * g.ws;
                ys += q;
                if (ys >= g.hs) { const int r = ys / g.hs; ys -= r * g.hs; img += r; }
            }
        }
    } else if (lane == 0) {
        // ================= issuer =================
        const uint32_t idesc = umma_idesc_bf16_mn(128, ROWS ? g.nb : g.nn);
        for (int k = 0; k < nst; ++k) {
            const int slot = k % g.nstage;
            mbar_wait(&bars[slot], (k / g.nstage) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sA = smem_u32(wg_smem + slot * g.stage_bytes), sB = sA + 2 * WGG_A_PART;
#pragma unroll
            for (int t = 0; t < WGG_KP / 16; ++t) {                        // K step: pixels 16t .. 16t+15 = 256 bytes into every group
                const uint64_t dA1 = umma_desc_mn(sA + t * 256, 128, WGG_GROUP), dA2 = dA1 + (uint64_t)(WGG_A_PART >> 4);
                if constexpr (!ROWS) {
                    const uint64_t dB1 = umma_desc_mn(sB + t * 256, 128, WGG_GROUP), dB2 = dB1 + (uint64_t)(g.b_part >> 4);
                    umma_bf16(tmem, dA1, dB1, idesc, k > 0 || t > 0);
                    umma_bf16(tmem, dA1, dB2, idesc, true);
                    umma_bf16(tmem, dA2, dB1, idesc, true);
                } else {
                    // pixel 16t of the stage sits at (row r, column x) of the segment; tap (ky, kx) reads the row tile of ky
                    // kx pixels further right (the tile starts `pad` pixels left of the segment).  One MMA reads 4 KB of A and
                    // 32 * nb bytes of B from shared memory: with all kh * kw taps in one CTA (nb <= 56) the re-reads of A
                    // bound the kernel (measured: 0.81 vs 0.62 ms on dres2), so a tile holds ONE kernel row and nb up to 160
                    const int r = (16 * t) / g.seg, x = (16 * t) - r * g.seg;
                    const uint32_t row0 = sB + (uint32_t)(r * g.bpr + x) * 16;
                    for (int kyi = 0; kyi < g.kyn; ++kyi)
                        for (int kx = 0; kx < g.kw; ++kx) {
                            const uint64_t dB1 = umma_desc_mn(row0 + kyi * 2 * g.b_part + kx * 16, 128, g.bp * 16), dB2 = dB1 + (uint64_t)(g.b_part >> 4);
                            const uint32_t d = tmem + (uint32_t)((kyi * g.kw + kx) * g.nb);
                            umma_bf16(d, dA1, dB1, idesc, k > 0 || t > 0);
                            umma_bf16(d, dA1, dB2, idesc, true);
                            umma_bf16(d, dA2, dB1, idesc, true);
                        }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[4 + slot])) : "memory");
        }
    }

    // ---- epilogue: the partial tile [128 a][nn] of this split -> workspace (each lane: runs of 16 consecutive floats)
    if (warp < 8) {
        const int last = nst - 1;
        mbar_wait(&bars[4 + last % g.nstage], (last / g.nstage) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, half = warp >> 2;
        float* wp = ws + (((int64_t)split * g.tiles + tile) * 128 + q * 32 + lane) * g.nn;
        for (int col0 = half * 16; col0 < g.nn; col0 += 32) {
            uint32_t v[16];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<uint4*>(wp + col0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(g.tmem_cols) : "memory");
}

// dW[a][b][ky][kx] = sum over splits of the partial tiles, in split order (deterministic); overwrites dW.
// One CTA per (output row a, chunk of WGG_RED_BC channels b): it adds the splits of that row segment in every tap-group tile
// (128-byte pieces of workspace rows), lays the sums out as [b][tap] in shared memory and writes them as one contiguous run
// of dW (a tile holds [tap][b]: writing straight from the workspace order would scatter 4-byte words `taps` apart).
constexpr int WGG_RED_BC = 32;
__global__ void __launch_bounds__(128) conv_gen_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int64_t s_a, int64_t s_b,
                                                                    int64_t s_ky, int64_t s_kx, const __grid_constant__ WggGeo g) {
    __shared__ float sm[WGG_RED_BC * 50];                              // [32 b][T | 1], T <= 49
    const int T = g.kh * g.kw, TP = T | 1;
    const int al = blockIdx.x % 128, bt = (blockIdx.x / 128) % g.nbt, at = blockIdx.x / (128 * g.nbt);
    const int a = at * 128 + al, b0 = blockIdx.y * WGG_RED_BC;         // b0: first channel of this chunk inside the b tile
    if (a >= g.ca || bt * g.nb + b0 >= g.cb) return;
    const int64_t per_split = (int64_t)g.tiles * 128 * g.nn;
    const int lane = threadIdx.x & 31;
    const bool live = b0 + lane < g.nb;
    for (int tap = threadIdx.x >> 5; tap < T; tap += 4) {              // one warp per tap: 32 consecutive b of one workspace row
        const int tgi = tap / g.tg, tl = tap - tgi * g.tg;
        const int tile = (at * g.nbt + bt) * g.ntg + tgi;
        const float* p = ws + ((int64_t)tile * 128 + al) * g.nn + tl * g.nb + b0 + lane;
        float acc = 0.f;
        if (live) {
#pragma unroll 8
            for (int sp = 0; sp < g.splits; ++sp) acc += __ldg(p + sp * per_split);
        }
        sm[lane * TP + tap] = acc;
    }
    __syncthreads();
    const int nb_here = min(WGG_RED_BC, min(g.nb - b0, g.cb - (bt * g.nb + b0)));
    float* row = dw + a * s_a + (int64_t)(bt * g.nb + b0) * s_b;
    const bool dense = s_kx == 1 && s_ky == g.kw && s_b == T;
    for (int q = threadIdx.x; q < nb_here * T; q += blockDim.x) {
        const int bl = q / T, tap = q - bl * T;
        const float v = sm[bl * TP + tap];
        if (dense) row[q] = v;
        else row[bl * s_b + (tap / g.kw) * s_ky + (tap % g.kw) * s_kx] = v;
    }
}

static int wgg_setup(WggGeo& g, int n, int ca, int cb, int hs, int ws_, int hl, int wl, int kh, int kw, int stride, int pad) {
    if (kh < 1 || kw < 1 || kh > 7 || kw > 7 || (stride != 1 && stride != 2) || pad < 0 || pad > 7) {
        set_error("conv_wgrad: kernel %dx%d stride %d padding %d is outside 1..7 / {1,2} / 0..7", kh, kw, stride, pad);
        return FFWM_ERR_ARG;
    }
    const int eh = (hs - 1) * stride - 2 * pad + kh, ew = (ws_ - 1) * stride - 2 * pad + kw;   // extent of L the taps reach
    if (hs < 1 || ws_ < 1 || hl < eh || hl >= eh + stride || wl < ew || wl >= ew + stride) {
        set_error("conv_wgrad: small side %dx%d and large side %dx%d do not match kernel %dx%d, stride %d, padding %d", hs, ws_, hl, wl, kh,
                  kw, stride, pad);
        return FFWM_ERR_SHAPE;
    }
    g.n = n, g.hs = hs, g.ws = ws_, g.hl = hl, g.wl = wl, g.ca = ca, g.cb = cb, g.kh = kh, g.kw = kw, g.stride = stride, g.pad = pad;
    const int T = kh * kw;
    // N = tg * nb <= 256, a multiple of 16, nb a multiple of 8: fewest staged groups over all tiles wins
    int64_t best = -1;
    for (int tg = T; tg >= 1; --tg) {
        if (T % tg) continue;
        const int gran = (tg % 2) ? 16 : 8;
        int nb = (256 / tg) / gran * gran;
        nb = std::min(nb, (cb + gran - 1) / gran * gran);
        if (nb < 8) continue;
        const int64_t tiles_bt = (int64_t)((cb + nb - 1) / nb) * (T / tg);
        const int64_t cost = tiles_bt * (16 + tg * nb / 8);
        if (best < 0 || cost < best) best = cost, g.tg = tg, g.nb = nb;
    }
    if (best < 0) { set_error("conv_wgrad: no tiling for %d taps", T); return FFWM_ERR_ARG; }
    // row mode (stride 1, rows of S in whole 64-pixel stages): a tile is one KERNEL ROW: one [nb channels][segment + kw - 1
    // pixels] tile of L feeds the kw taps of that row through descriptor offsets, kw * nb <= 512 TMEM columns: L is loaded
    // once per kernel row instead of once per tap, S once per (b tile, kernel row) with nb up to 160 instead of 200 / tap.
    // Cost in (8-channel group, pixel) loads per stage, same unit as `best` * 64.
    g.rows_mode = 0, g.seg = g.rps = g.bpr = g.bp = 0, g.kyn = 1;
    const int seg = std::min(ws_, 64);
    if (stride == 1 && kw >= 2 && !opt(OPT_WGRAD_NO_ROWS) && ws_ % 16 == 0 && 64 % seg == 0 && ws_ % seg == 0 && hs % (64 / seg) == 0) {
        const int rps = 64 / seg, bpr = seg + kw - 1, bp = rps * bpr, kyn = 1;
        int64_t best_rows = best * 64;
        // nb >= 96: an MMA re-reads its 4 KB of A for 32 * nb bytes of B; with narrow tiles those re-reads cost more shared-
        // memory bandwidth than the shared loads save (measured: 64 -> 64 channels 0.135 ms in row mode, 0.117 ms gathered)
        for (int nb = 96; nb <= 256 && kyn * kw * nb <= 512; nb += 16) {
            if (nb - 16 >= cb) break;                                      // wider than the channel count rounded up to 16
            const int nbg = nb / 8;
            if ((int64_t)kyn * nbg * bp > 4 * WGG_PRODUCERS) continue;     // at most 4 items per producer thread
            if (2 * (2 * WGG_A_PART + (int64_t)kyn * 2 * nbg * bp * 16) > 227 * 1024 - 256) continue;   // two stages must fit
            const int64_t nbt = (cb + nb - 1) / nb;
            const int64_t cost = (kh / kyn) * nbt * (16 * 64 + (int64_t)kyn * nbg * bp);
            if (cost < best_rows) best_rows = cost, g.rows_mode = 1, g.tg = kyn * kw, g.nb = nb, g.seg = seg, g.rps = rps, g.bpr = bpr, g.bp = bp, g.kyn = kyn;
        }
    }
    g.nn = g.tg * g.nb;
    g.nat = (ca + 127) / 128, g.nbt = (cb + g.nb - 1) / g.nb, g.ntg = T / g.tg;
    g.tiles = g.nat * g.nbt * g.ntg;
    g.npix = (int64_t)n * hs * ws_;
    const int64_t stages = (g.npix + WGG_KP - 1) / WGG_KP;
    if (stages > 0x7fffffffLL) { set_error("conv_wgrad: problem too large"); return FFWM_ERR_TOO_LARGE; }
    g.stages_total = (int)stages;
    // split K over CTAs: one CTA per SM is resident (shared memory), so the kernel takes ceil(tiles * splits / SMs) waves of
    // (stages per split + ~6 stages of prologue / epilogue); the reduction pass reads `splits` partials per weight.  Pick the
    // chain length that minimises that, no longer than the accuracy cap.
    const int cap = std::max(1, std::min(opt(OPT_WGRAD_CHAIN) > 0 ? opt(OPT_WGRAD_CHAIN) : WGG_MAX_STAGES, 1024));
    const int sms = sm_count();
    double best_cost = -1.0;
    for (int sps = (int)std::min<int64_t>(cap, stages); sps >= 1; --sps) {
        const int64_t sp = (stages + sps - 1) / sps;
        if (sp > 1 && sps < std::min<int64_t>(cap, 8)) break;             // never shorter than 8 stages unless the whole problem is
        const int64_t waves = (g.tiles * sp + sms - 1) / sms;
        const double cost = (double)waves * (sps + 6) + 0.5 * (double)sp;
        if (best_cost < 0 || cost < best_cost) best_cost = cost, g.stages_per_split = sps, g.splits = (int)sp;
    }
    g.b_part = g.rows_mode ? (g.nb / 8) * g.bp * 16 : (g.nn / 8) * WGG_GROUP;
    g.stage_bytes = 2 * WGG_A_PART + (g.rows_mode ? g.kyn : 1) * 2 * g.b_part;
    g.nstage = std::max(2, std::min(3, (227 * 1024 - 256) / g.stage_bytes));
    g.tmem_cols = 32;
    while (g.tmem_cols < g.nn) g.tmem_cols *= 2;
    return FFWM_OK;
}

}  // namespace ffwm

// Bytes of workspace ffwm_conv_wgrad needs for these shapes (0: unsupported arguments, see ffwm_last_error).
extern "C" int64_t ffwm_conv_wgrad_workspace_bytes(int n, int ca, int cb, int hs, int ws, int hl, int wl, int kh, int kw, int stride, int pad) {
    ffwm::WggGeo g;
    if (n <= 0 || ca <= 0 || cb <= 0) return 0;
    if (ffwm::wgg_setup(g, n, ca, cb, hs, ws, hl, wl, kh, kw, stride, pad)) return 0;
    return (int64_t)g.splits * g.tiles * 128 * g.nn * (int64_t)sizeof(float);
}

// grad_weight (Ca, Cb, kh, kw; any strides; OVERWRITTEN) = weight gradient as defined at the top of this file.
// small (B, Ca, Hs, Ws) and large (B, Cb, Hl, Wl): nn.Conv2d: small = grad_out, large = input; nn.ConvTranspose2d:
// small = input, large = grad_out.  workspace: ffwm_conv_wgrad_workspace_bytes(...) bytes of device memory.
extern "C" int ffwm_conv_wgrad(const ffwm_tensor4* small, const ffwm_tensor4* large, const ffwm_tensor4* grad_weight, int stride, int pad,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace ffwm;
    View<const float> sv, lv;
    View<float> wv;
    int rc;
    if ((rc = make_view<const float>(small, "small", &sv))) return rc;
    if ((rc = make_view<const float>(large, "large", &lv))) return rc;
    if ((rc = make_view<float>(grad_weight, "grad_weight", &wv))) return rc;
    if (sv.n != lv.n || wv.n != sv.c || wv.c != lv.c) {
        set_error("conv_wgrad: grad_weight %dx%dx%dx%d does not match small %dx%d / large %dx%d channels", wv.n, wv.c, wv.h, wv.w, sv.n, sv.c, lv.n, lv.c);
        return FFWM_ERR_SHAPE;
    }
    if ((int64_t)wv.n * wv.c * wv.h * wv.w == 0) return FFWM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((int64_t)sv.n * sv.h * sv.w == 0) {      // empty batch: the gradient is zero
        for (int a = 0; a < wv.n; ++a)
            for (int b = 0; b < wv.c; ++b)
                cudaMemset2DAsync(wv.p + a * wv.sb + b * wv.sc, sizeof(float) * wv.sh, 0, sizeof(float) * wv.w, wv.h, st);
        return FFWM_OK;
    }
    WggGeo g;
    if ((rc = wgg_setup(g, sv.n, sv.c, lv.c, sv.h, sv.w, lv.h, lv.w, wv.h, wv.w, stride, pad))) return rc;
    const int64_t need = (int64_t)g.splits * g.tiles * 128 * g.nn * (int64_t)sizeof(float);
    if (!workspace || workspace_bytes < need) { set_error("conv_wgrad: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes, (long long)need); return FFWM_ERR_SHAPE; }
    if (g.splits > 65535) { set_error("conv_wgrad: grid too large"); return FFWM_ERR_TOO_LARGE; }
    auto kern = g.rows_mode ? conv_gen_wgrad_tc_kernel<true> : conv_gen_wgrad_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    kern<<<dim3(g.tiles, g.splits), WGG_PRODUCERS + 32, g.nstage * g.stage_bytes + 128, st>>>(sv, lv, static_cast<float*>(workspace), g);
    if ((rc = check_launch("conv_wgrad"))) return rc;
    conv_gen_wgrad_reduce_kernel<<<dim3(g.nat * g.nbt * 128, (g.nb + WGG_RED_BC - 1) / WGG_RED_BC), 128, 0, st>>>(
        static_cast<const float*>(workspace), wv.p, wv.sb, wv.sc, (int64_t)wv.sh, (int64_t)wv.sw, g);
    return check_launch("conv_wgrad (reduce)");
}
