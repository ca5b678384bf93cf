// block_extractor forward on the rolling-strip gather (roll_gather.cuh), k = 2 or 3, fp32.
// Included by block_extractor.cu after block_tap().
//
// The k*k samples of a flow pixel share a (k+1)x(k+1) window of the source (same fractional part,
// integer offsets), and the ring holds the source edge-replicated (the reference clamps every tap
// index), so a pixel whose window lies inside the ring reads `row_offset[n] + m` with immediate m.
// Lanes are channels; a warp owns 8 flow pixels of one flow row.  The output is k*k times larger
// than the source — the kernel is store-bound — so every output row segment (8k floats per
// channel) is transposed through a warp-private staging buffer (pitch 36: conflict-free both ways)
// and leaves as 128-bit stores that cover whole 32-byte sectors.
#pragma once
#include "roll_gather.cuh"

namespace ffwm {

constexpr int BER_PW = 8;        // [0] off0|off1<<16  [1] off2|off3<<16|slow<<31  [2] yBP0 [3] yBP1  [4..6] xRP  [7] yBP2
constexpr int BER_PITCH = 36;    // staging [8k columns][36]: lanes = channels on the way in, (channel, 32-byte pair) on the way out

template <int K>
__global__ void __launch_bounds__(RG_THREADS, 1)
block_extractor_fwd_roll_kernel(View<const float> src, View<const float> flow, View<float> out, int vec_ok) {
    constexpr int NC = RG_PXW * K;                                             // output columns per warp and row
    extern __shared__ __align__(16) unsigned char rg_smem_raw[];
    float* slab = reinterpret_cast<float*>(rg_smem_raw);                       // [32][1025]
    float* prm_all = slab + 32 * RG_CHP;                                       // [16 warps][32 px][8]
    float* stage_all = prm_all + RG_WARPS * 32 * BER_PW;                       // [16 warps][8k][36]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = blockIdx.x * RG_SW, c0 = blockIdx.y * 32, b = blockIdx.z;
    const int nch = min(32, out.c - c0);
    const int rx0 = x0 - RG_M;
    const int nsteps = (flow.h + RG_SH - 1) / RG_SH;
    const int wrow = warp >> 1, xw0 = x0 + (warp & 1) * RG_PXW;
    float* prm = prm_all + warp * (32 * BER_PW);
    float* stage = stage_all + warp * (NC * BER_PITCH);
    const float* slab_lane = slab + lane * RG_CHP;
    const float* plane_lane = src.p + b * src.sb + (int64_t)(c0 + min(lane, nch - 1)) * src.sc;
    const bool full = xw0 + RG_PXW <= flow.w;                                  // this warp's 8 flow pixels all exist

    rg_fill_rows<false>(slab, src, b, c0, nch, rx0, -RG_M, RG_M + 2 * RG_SH, warp, lane);   // rows [-8, 16)

    const int gsub = lane >> 3, gx = xw0 + (lane & 7);
    float nfx = 0.f, nfy = 0.f;
    auto load_flow = [&](int s_base) {
        const int y = (s_base + gsub) * RG_SH + wrow;
        nfx = nfy = 0.f;
        if (y < flow.h && gx < flow.w) {
            const float* f = flow.p + b * flow.sb + y * flow.sh + gx * flow.sw;
            nfx = __ldg(f); nfy = __ldg(f + flow.sc);
        }
    };
    load_flow(0);

    for (int s = 0; s < nsteps; ++s) {
        if ((s & (RG_BLK - 1)) == 0) {
            __syncwarp();
            const int ystep = (s + gsub) * RG_SH, yf = ystep + wrow;
            float xRP[K], yBP[K], flx[K], fly[K];
            bool fast = true;
#pragma unroll
            for (int j = 0; j < K; ++j) {                    // block_extractor_kernel.cu:57-76, same expression order
                const float flow_x = nfx + (j - K / 2), flow_y = nfy + (j - K / 2);
                const float dx = flow_x + float(gx), dy = flow_y + float(yf);
                flx[j] = floorf(dx); fly[j] = floorf(dy);
                xRP[j] = dx - flx[j]; yBP[j] = dy - fly[j];
                fast = fast && flx[j] == flx[0] + float(j) && fly[j] == fly[0] + float(j);
            }
            fast = fast && flx[0] >= float(rx0) && flx[0] + float(K) <= float(rx0 + RG_RW - 1) &&
                   fly[0] >= float(ystep - RG_M) && fly[0] + float(K) <= float(ystep + RG_SH + RG_M - 1);
            int off[4] = {0, 0, 0, 0};
            if (fast) {
                const int cb = int(flx[0]) - rx0, r0 = int(fly[0]);
#pragma unroll
                for (int n = 0; n <= K; ++n) off[n] = ((r0 + n) & (RG_RING - 1)) * RG_RW + cb;
            }
            float4* P4 = reinterpret_cast<float4*>(prm + lane * BER_PW);
            P4[0] = make_float4(__int_as_float(off[0] | (off[1] << 16)),
                                __int_as_float(off[2] | (off[3] << 16) | (fast ? 0 : int(0x80000000u))), yBP[0], yBP[1]);
            P4[1] = make_float4(xRP[0], xRP[1], K > 2 ? xRP[K - 1] : 0.f, K > 2 ? yBP[K - 1] : 0.f);
            load_flow(s + RG_BLK);
            __syncwarp();
        }
        rg_cp_async_wait_all();
        __syncthreads();
        if (s + 1 < nsteps) rg_fill_rows<false>(slab, src, b, c0, nch, rx0, (s + 1) * RG_SH + RG_M, RG_SH, warp, lane);
        const int yf = s * RG_SH + wrow;
        if (yf < flow.h) {                                   // warp-uniform
#pragma unroll 1
            for (int i = 0; i < K; ++i) {                    // output row yf*K + i
#pragma unroll 4
                for (int px = 0; px < RG_PXW; ++px) {
                    if (xw0 + px >= flow.w) break;           // warp-uniform
                    const float4* P4 = reinterpret_cast<const float4*>(prm + ((s & (RG_BLK - 1)) * RG_PXW + px) * BER_PW);
                    const float4 h = P4[0], h2 = P4[1];
                    const float xRPv[3] = {h2.x, h2.y, h2.z};
                    const float yB = i == 0 ? h.z : (i == 1 ? h.w : h2.w), yT = 1 - yB;
                    float* sp = stage + (px * K) * BER_PITCH + lane;
                    const int o23 = __float_as_int(h.y);
                    if (o23 >= 0) {
                        const int o01 = __float_as_int(h.x);
                        const int ot = i == 0 ? (o01 & 0xffff) : (i == 1 ? (o01 >> 16) : (o23 & 0xffff));
                        const int ob = i == 0 ? (o01 >> 16) : (i == 1 ? (o23 & 0xffff) : (o23 >> 16));
                        const float* rt = slab_lane + ot;
                        const float* rb = slab_lane + ob;
                        float top[K + 1], bot[K + 1];
#pragma unroll
                        for (int m = 0; m <= K; ++m) { top[m] = rt[m]; bot[m] = rb[m]; }
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            const float xR = xRPv[j], xL = 1 - xR;
                            float sample = 0.f;
                            sample += xL * yT * top[j];
                            sample += xR * yT * top[j + 1];
                            sample += xL * yB * bot[j];
                            sample += xR * yB * bot[j + 1];
                            sp[j * BER_PITCH] = sample;
                        }
                    } else {
                        // window not inside the ring (or its taps are not consecutive): the reference's
                        // per-tap arithmetic with global loads
                        const float* f = flow.p + b * flow.sb + yf * flow.sh + (xw0 + px) * flow.sw;
                        const float fx_raw = __ldg(f), fy_raw = __ldg(f + flow.sc);
#pragma unroll 1
                        for (int j = 0; j < K; ++j) {
                            const Bilin<float> t = block_tap<float>(fx_raw, fy_raw, xw0 + px, yf, j - K / 2, i - K / 2, src.h, src.w);
                            float sample = 0.f;
                            sample += t.xL_P * t.yT_P * __ldg(plane_lane + t.yT * src.sh + t.xL * src.sw);
                            sample += t.xR_P * t.yT_P * __ldg(plane_lane + t.yT * src.sh + t.xR * src.sw);
                            sample += t.xL_P * t.yB_P * __ldg(plane_lane + t.yB * src.sh + t.xL * src.sw);
                            sample += t.xR_P * t.yB_P * __ldg(plane_lane + t.yB * src.sh + t.xR * src.sw);
                            sp[j * BER_PITCH] = sample;
                        }
                    }
                }
                __syncwarp();
                float* orow = out.p + b * out.sb + (int64_t)c0 * out.sc + (int64_t)(yf * K + i) * out.sh;
                if (vec_ok && full) {
                    // lane -> (channel lane/2 of a half of the group, which 16 bytes of a 32-byte pair)
                    const int csub = lane >> 1, qh = lane & 1;
#pragma unroll
                    for (int it = 0; it < 2 * K; ++it) {
                        const int c = (it & 1) * 16 + csub, col = 4 * (2 * (it >> 1) + qh);
                        const float* sq = stage + col * BER_PITCH + c;
                        const float4 v = make_float4(sq[0], sq[BER_PITCH], sq[2 * BER_PITCH], sq[3 * BER_PITCH]);
                        if (c < nch) __stcs(reinterpret_cast<float4*>(orow + (int64_t)c * out.sc + xw0 * K + col), v);
                    }
                } else {
                    const int ncol = min(RG_PXW, flow.w - xw0) * K;
                    for (int idx = lane; idx < 32 * NC; idx += 32) {
                        const int c = idx / NC, col = idx - c * NC;
                        if (c < nch && col < ncol) st_stream(orow + (int64_t)c * out.sc + (xw0 * K + col) * out.sw, stage[col * BER_PITCH + c]);
                    }
                }
                __syncwarp();
            }
        }
    }
}

template <int K>
static int launch_be_fwd_roll(const View<const float>& src, const View<const float>& flow, const View<float>& out, cudaStream_t st) {
    const size_t smem = sizeof(float) * (32 * RG_CHP + RG_WARPS * 32 * BER_PW + RG_WARPS * RG_PXW * K * BER_PITCH);
    cudaError_t e = cudaFuncSetAttribute(block_extractor_fwd_roll_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("block_extractor_fwd_roll: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
    const int vec_ok = out.sw == 1 && (out.sh & 3) == 0 && (out.sc & 3) == 0 && (out.sb & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(out.p) & 15) == 0;
    dim3 grid(ceil_div(flow.w, RG_SW), ceil_div(out.c, 32), out.n);
    block_extractor_fwd_roll_kernel<K><<<grid, RG_THREADS, smem, st>>>(src, flow, out, vec_ok);
    return FFWM_OK;
}

}  // namespace ffwm
