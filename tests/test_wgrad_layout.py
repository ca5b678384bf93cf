"""CPU emulation of the staging / descriptor / accumulator indexing of the tcgen05 weight-gradient kernel
(ffwm_b200/csrc/conv3x3_wgrad_tc.cu).  The kernel was written without GPU access; what CAN be checked
here is its arithmetic plan: the strided K chunks {j, j+S, j+2S, j+3S}, the halo chunks, horizontal taps
as descriptor start offsets, vertical taps as row groups of one B operand, the hi/lo split, the
split-K stage decomposition and the TMEM column -> (kx, ky, ci) mapping of the epilogue.  The emulation
stages bytes into a flat "shared memory" with the kernel's offsets and evaluates each MMA through the
K-major / no-swizzle operand addressing the forward kernel already relies on
(element (row r, k) at start + (r/8)*SBO + (r%8)*16 + (k/4)*LBO + (k%4)*4), then compares the result
with torch's weight gradient.  Constants are parsed from the .cu so the two cannot drift apart."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

CU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ffwm_b200", "csrc", "conv3x3_wgrad_tc.cu")


def constants():
    src = open(CU).read()
    c = {}
    for name in ("WG_MT", "WG_NCI", "WG_CH", "WG_PRODUCERS", "WG_MAX_STAGES"):
        c[name] = int(re.search(r"constexpr int %s = (\d+);" % name, src).group(1))
    c["WG_N"] = 3 * c["WG_NCI"]
    c["A_LBO"] = c["WG_MT"] * 16
    c["A_PART"] = c["WG_CH"] * c["A_LBO"]
    c["B_LBO"] = c["WG_N"] * 16
    c["B_PART"] = (c["WG_CH"] + 2) * c["B_LBO"]
    c["STAGE"] = 2 * c["A_PART"] + 2 * c["B_PART"]
    # the relations the emulation hard-codes must be the ones in the source
    assert "constexpr int WG_N = 3 * WG_NCI;" in src and "constexpr int WG_A_LBO = WG_MT * 16;" in src
    assert "constexpr int WG_B_PART = (WG_CH + 2) * WG_B_LBO;" in src
    assert "constexpr int WG_STAGE = 2 * WG_A_PART + 2 * WG_B_PART;" in src
    return c


def split(v):
    hi = (v.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, (v - hi).astype(np.float32)


def operand(smem, start, rows, lbo, sbo=128):
    """rows x 8 tf32 operand of one MMA read through a K-major no-swizzle descriptor (byte offsets)."""
    r = np.arange(rows)[:, None]
    k = np.arange(8)[None, :]
    off = start + (r // 8) * sbo + (r % 8) * 16 + (k // 4) * lbo + (k % 4) * 4
    return smem[off // 4]


def emulate(x, go, splits_hint):
    c = constants()
    MT, NCI, CH, N = c["WG_MT"], c["WG_NCI"], c["WG_CH"], c["WG_N"]
    B, cin, H, W = x.shape
    cout = go.shape[1]
    S = W // 4
    ncb = S // CH
    n_ci_tiles = -(-cin // NCI)
    tiles = -(-cout // MT) * n_ci_tiles
    stages_total = B * H * ncb
    splits = max(1, min(stages_total, splits_hint // tiles))
    per = min(-(-stages_total // splits), c["WG_MAX_STAGES"])
    splits = -(-stages_total // per)
    dw = np.zeros((cout, cin, 3, 3), np.float64)
    dbias = np.zeros(cout, np.float64)
    for tile in range(tiles):
        cot, cit = divmod(tile, n_ci_tiles)
        for sp in range(splits):
            st0 = sp * per
            nst = min(per, stages_total - st0)
            assert nst >= 1
            acc = np.zeros((MT, 3 * N), np.float64)                       # TMEM: lanes x columns
            for k in range(nst):
                s = st0 + k
                j0 = (s % ncb) * CH
                rr = s // ncb
                y, b = rr % H, rr // H
                smem = np.full(c["STAGE"] // 4, np.nan, np.float32)       # every slot must be (re)written
                # ---- producers: item = 8 rows, lane -> (r8, jq), chunks cc = jq + 4q
                for it in range(MT // 8 + N // 8):
                    is_a = it < MT // 8
                    for r8 in range(8):
                        row = (it if is_a else it - MT // 8) * 8 + r8
                        if is_a:
                            co = cot * MT + row
                            rok = co < cout
                            src = go[b, co, y] if rok else None
                            jbase, nchunk, lbo, part, base = j0, CH, c["A_LBO"], c["A_PART"], 0
                        else:
                            ky, ci = row // NCI, cit * NCI + row % NCI
                            yy = y + ky - 1
                            rok = ci < cin and 0 <= yy < H
                            src = x[b, ci, yy] if rok else None
                            jbase, nchunk, lbo, part, base = j0 - 1, CH + 2, c["B_LBO"], c["B_PART"], 2 * c["A_PART"]
                        for jq in range(4):
                            for q in range(3):
                                cc = jq + 4 * q
                                if cc >= nchunk:
                                    continue
                                v = np.zeros(4, np.float32)
                                for m in range(4):
                                    p = jbase + cc + m * S
                                    if rok and 0 <= p < W:
                                        v[m] = src[p]
                                if is_a and cit == 0 and rok:                  # fused bias gradient: A items of ci tile 0
                                    dbias[co] += float(v.astype(np.float64).sum())
                                hi, lo = split(v)
                                d = (base + row * 16 + cc * lbo) // 4
                                smem[d:d + 4] = hi
                                smem[d + part // 4:d + part // 4 + 4] = lo
                assert not np.isnan(smem).any(), "a slot the tensor core reads was never staged"
                # ---- issuer: 4 K steps x 3 horizontal taps x 3 split terms
                sA, sB = 0, 2 * c["A_PART"]
                for t in range(CH // 2):
                    a_hi = operand(smem, sA + 2 * t * c["A_LBO"], MT, c["A_LBO"]).astype(np.float64)
                    a_lo = operand(smem, sA + c["A_PART"] + 2 * t * c["A_LBO"], MT, c["A_LBO"]).astype(np.float64)
                    for kx in range(3):
                        b_hi = operand(smem, sB + (2 * t + kx) * c["B_LBO"], N, c["B_LBO"]).astype(np.float64)
                        b_lo = operand(smem, sB + c["B_PART"] + (2 * t + kx) * c["B_LBO"], N, c["B_LBO"]).astype(np.float64)
                        acc[:, kx * N:(kx + 1) * N] += a_hi @ b_hi.T + a_hi @ b_lo.T + a_lo @ b_hi.T
            # ---- epilogue: two passes of 64 output channels through a [64][9*NCI+1] tile in dW's (ci, ky, kx) order
            pitch = 9 * NCI + 1
            assert "constexpr int WG_EPI_PITCH = 9 * WG_NCI + 1;" in open(CU).read()
            for ps in range(2):
                tile = np.full(64 * pitch, np.nan)
                for warp in range(8):
                    q, half = warp & 3, warp >> 2
                    if (q >> 1) != ps:
                        continue
                    for blk in range(half, 3 * N // 16, 2):
                        c0 = blk * 16
                        kx, n0 = divmod(c0, N)
                        ky, cil = divmod(n0, NCI)
                        for lane in range(32):
                            for j in range(16):
                                tile[((q & 1) * 32 + lane) * pitch + (cil + j) * 9 + ky * 3 + kx] = acc[q * 32 + lane, c0 + j]
                nci = min(NCI, cin - cit * NCI)
                for r in range(64):
                    co = cot * MT + ps * 64 + r
                    if co >= cout:
                        break
                    for e in range(nci * 9):
                        ci_l, t9 = divmod(e, 9)
                        ky, kx = divmod(t9, 3)
                        assert not np.isnan(tile[r * pitch + e])
                        dw[co, cit * NCI + ci_l, ky, kx] += tile[r * pitch + e]
    return dw, dbias


@pytest.mark.parametrize("b,cin,cout,h,w,splits_hint", [(2, 5, 3, 3, 32, 148), (1, 50, 130, 2, 64, 7), (1, 3, 8, 1, 32, 1),
                                                        (1, 2, 2, 70, 32, 1), (1, 100, 140, 2, 96, 13)])
def test_wgrad_plan_matches_torch(b, cin, cout, h, w, splits_hint):
    g = torch.Generator().manual_seed(cin * 100 + cout)
    x = torch.randn(b, cin, h, w, generator=g)
    go = torch.randn(b, cout, h, w, generator=g)
    wt = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, None, padding=1).backward(go.double())
    got, got_bias = emulate(x.numpy(), go.numpy(), splits_hint)
    want = wt.grad.numpy()
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err <= 2e-6, err          # only the dropped lo*lo terms (2^-22 relative) separate the two
    np.testing.assert_allclose(got_bias, go.double().sum((0, 2, 3)).numpy(), rtol=0, atol=1e-9)   # every element exactly once


def test_wgrad_constants_fit_the_hardware():
    c = constants()
    assert 3 * c["WG_N"] <= 512 and c["WG_N"] % 16 == 0 and 16 <= c["WG_N"] <= 256      # TMEM columns, MMA N for M = 128
    assert 2 * c["STAGE"] + 64 <= 227 * 1024                                              # shared memory per CTA
    assert c["A_LBO"] % 128 == 0 and c["B_LBO"] % 128 == 0                                # as in the validated forward kernel
    assert (c["WG_MT"] // 8 + c["WG_N"] // 8) <= 5 * (c["WG_PRODUCERS"] // 32)            # work items per producer warp <= 5
