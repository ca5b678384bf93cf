"""CPU tests of the C oracle (oracle/warp_ops.*) — no GPU needed.

Pins the restatement against (a) the reference's own known answers
(test_local_attn_reshape.py:29-43), (b) its gradcheck recipes
(test_block_extractor.py:77-81, test_local_attn_reshape.py:66-70),
(c) algebraic identities (pixel_shuffle / unfold), (d) torch's grid_sample
(the third-party kernel behind WarpNet), and (e) golden outputs of the
reference's CUDA kernels recorded on a B200 (tests/golden/ref_cuda_*.npz).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch.autograd import gradcheck

from tests.golden import cases

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class _Resample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W, a, b, ks, dil):
        ctx.W, ctx.k = W, (ks, dil)
        ctx.save_for_backward(a, b)
        return W.resample2d_forward(a, b, ks, dil)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g1, g2 = ctx.W.resample2d_backward(a, b, g.contiguous(), *ctx.k)
        return None, g1, g2, None, None


class _Block(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W, s, f, k):
        ctx.W, ctx.k = W, k
        ctx.save_for_backward(s, f)
        return W.block_extractor_forward(s, f, k)

    @staticmethod
    def backward(ctx, g):
        s, f = ctx.saved_tensors
        gs, gf = ctx.W.block_extractor_backward(s, f, g.contiguous(), ctx.k)
        return None, gs, gf, None


class _Reshape(torch.autograd.Function):
    @staticmethod
    def forward(ctx, W, x, k):
        ctx.W, ctx.k = W, k
        ctx.save_for_backward(x)
        return W.local_attn_reshape_forward(x, k)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return None, ctx.W.local_attn_reshape_backward(x, g.contiguous(), ctx.k), None


def test_local_attn_reshape_known_answer(oracle_warp):
    c = cases.local_attn_reshape_case("kat_0_8")
    out = oracle_warp.local_attn_reshape_forward(torch.tensor(c["x"], dtype=torch.float32), 3)
    assert out.shape == (2, 1, 30, 30)
    assert out[0, 0, :3, :3].tolist() == [[0, 1, 2], [3, 4, 5], [6, 7, 8]]
    assert torch.equal(out[1, 0, 27:, 27:], out[0, 0, :3, :3])


@pytest.mark.parametrize("k,shape", [(3, (4, 9, 14, 10)), (5, (2, 25, 6, 7)), (7, (1, 49, 5, 3)), (1, (2, 1, 4, 4))])
def test_local_attn_reshape_is_pixel_shuffle(oracle_warp, k, shape):
    x = torch.rand(*shape, dtype=torch.float64)
    out = oracle_warp.local_attn_reshape_forward(x, k)
    assert torch.equal(out, F.pixel_shuffle(x, k))
    g = torch.rand_like(out)
    assert torch.equal(oracle_warp.local_attn_reshape_backward(x, g, k), F.pixel_unshuffle(g, k))


def test_local_attn_reshape_gradcheck(oracle_warp):
    x = torch.rand(4, 9, 14, 10, dtype=torch.float64, requires_grad=True)
    assert gradcheck(lambda t: _Reshape.apply(oracle_warp, t, 3), (x,), fast_mode=True)
    x2 = torch.rand(1, 9, 4, 3, dtype=torch.float64, requires_grad=True)
    assert gradcheck(lambda t: _Reshape.apply(oracle_warp, t, 3), (x2,))


@pytest.mark.parametrize("k", [3, 5, 7])
def test_block_extractor_constant_flow_is_unfold(oracle_warp, k):
    # losses.py:212-217: flow == k//2 turns K4 into an im2col, bit-exactly
    b, c, hs, ws = 2, 3, 16, 13
    src = torch.rand(b, c, hs, ws)
    hf, wf = hs - k + 1, ws - k + 1
    flow = torch.full((b, 2, hf, wf), float(k // 2))
    out = oracle_warp.block_extractor_forward(src, flow, k)
    unf = F.unfold(src, k).view(b, c, k, k, hf, wf).permute(0, 1, 4, 2, 5, 3).reshape(b, c, hf * k, wf * k)
    assert torch.equal(out, unf)


def test_block_extractor_zero_flow_centre_is_identity(oracle_warp):
    # test_block_extractor.py:49-57: zero flow, k=3 -> centre of each block is the source pixel
    src = torch.rand(2, 3, 9, 7)
    out = oracle_warp.block_extractor_forward(src, torch.zeros(2, 2, 9, 7), 3)
    assert torch.equal(out[:, :, 1::3, 1::3], src)


def test_block_extractor_gradcheck_reference_recipe(oracle_warp):
    torch.manual_seed(0)
    s = torch.rand(4, 6, 14, 10, dtype=torch.float64, requires_grad=True)
    f = (torch.rand(4, 2, 14, 10, dtype=torch.float64) * 1.8).requires_grad_()
    # full recipe shape in fast mode (random projections), a slice of it exhaustively
    assert gradcheck(lambda a, b: _Block.apply(oracle_warp, a, b, 3), (s, f), fast_mode=True)
    s2 = s.detach()[:1, :2, :6, :5].clone().requires_grad_()
    f2 = f.detach()[:1, :, :6, :5].clone().requires_grad_()
    assert gradcheck(lambda a, b: _Block.apply(oracle_warp, a, b, 3), (s2, f2))


@pytest.mark.parametrize("ks", [2, 4, 6])
def test_resample2d_gradcheck(oracle_warp, ks):
    # away from integer coordinates and with xf,yf >= 0
    torch.manual_seed(ks)
    a = torch.rand(2, 3, 9, 8, dtype=torch.float64, requires_grad=True)
    fl = torch.rand(2, 2, 9, 8, dtype=torch.float64) * 0.6 + 0.2
    sg = torch.rand(2, 1, 9, 8, dtype=torch.float64) * 2 + 1
    b = torch.cat([fl, sg], 1).requires_grad_()
    assert gradcheck(lambda u, v: _Resample.apply(oracle_warp, u, v, ks, 1), (a, b), eps=1e-6, atol=1e-5)


def test_resample2d_large_sigma_is_tap_average(oracle_warp):
    # sigma -> inf: all weights equal -> plain mean of the ks*ks taps
    src = torch.rand(1, 2, 12, 12, dtype=torch.float64)
    in2 = torch.zeros(1, 3, 12, 12, dtype=torch.float64)
    in2[:, 2] = 1e6
    out = oracle_warp.resample2d_forward(src, in2, 4, 1)
    pad = F.pad(src, (1, 2, 1, 2), mode="replicate")      # clamped taps
    ref = F.avg_pool2d(pad, 4, stride=1)
    assert torch.allclose(out, ref, rtol=0, atol=1e-9)


def test_resample2d_zero_sigma_and_degenerate_kernel(oracle_warp):
    src = torch.rand(1, 2, 6, 6)
    in2 = torch.zeros(1, 3, 6, 6)
    # sigma == 0 takes the EPS branch of SAFE_DIV: weight exp(-d^2/1e-8) -> 1 at d == 0
    out = oracle_warp.resample2d_forward(src, in2, 2, 1)
    assert torch.equal(out, src)
    # kernel_size < 2: no taps, 0/EPS = 0
    assert torch.count_nonzero(oracle_warp.resample2d_forward(src, in2, 1, 1)) == 0


def test_resample2d_truncation_quirk_only_for_negative_coords(oracle_warp):
    # N2: backward-input1 weights use alpha = xf - int(xf).  For xf,yf >= 0 that is the
    # forward's floor-alpha and the analytic gradient wrt input1 is exact.  For negative
    # non-integer coordinates the weights differ; with ks=2 both taps clamp to column 0
    # so it cannot be seen, with ks=4 and xf in (-1,0) the tap at column 1 shows it.
    torch.manual_seed(1)
    sig = torch.full((1, 1, 6, 6), 2.0, dtype=torch.float64)
    pos = torch.cat([torch.rand(1, 2, 6, 6, dtype=torch.float64) * 0.8 + 0.1, sig], 1)
    neg = pos.clone()
    neg[:, 0, :, 0] -= 1.0                               # x = 0: xf in (-0.9, -0.1)
    fn = lambda b, ks: (lambda a: _Resample.apply(oracle_warp, a, b, ks, 1))
    a = torch.rand(1, 1, 6, 6, dtype=torch.float64, requires_grad=True)
    assert gradcheck(fn(pos, 2), (a,), eps=1e-6, atol=1e-6)
    assert gradcheck(fn(pos, 4), (a,), eps=1e-6, atol=1e-6)
    assert gradcheck(fn(neg, 2), (a,), eps=1e-6, atol=1e-6)
    assert not gradcheck(fn(neg, 4), (a,), eps=1e-6, atol=1e-6, raise_exception=False)


@pytest.mark.parametrize("dt,tol", [(torch.float32, 2e-5), (torch.float64, 1e-12)])
def test_grid_warp_matches_torch_grid_sample(oracle_warp, dt, tol):
    torch.manual_seed(0)
    for (b, c, hi, wi, ho, wo) in [(2, 5, 13, 11, 7, 9), (1, 3, 128, 128, 32, 32), (2, 4, 8, 8, 8, 8)]:
        img = torch.rand(b, c, hi, wi, dtype=dt, requires_grad=True)
        fl = (torch.rand(b, 2, ho, wo, dtype=dt) * 2.6 - 1.3).requires_grad_()   # includes out-of-image taps
        ref = F.grid_sample(img, fl.permute(0, 2, 3, 1), mode="bilinear", align_corners=False)
        go = torch.randn_like(ref)
        ref.backward(go)
        out = oracle_warp.grid_warp_forward(img.detach(), fl.detach())
        gi, gf = oracle_warp.grid_warp_backward(img.detach(), fl.detach(), go)
        assert (out - ref).abs().max() <= tol
        assert (gi - img.grad).abs().max() <= tol * 10
        assert (gf - fl.grad).abs().max() <= tol * 10 * max(hi, wi)


def test_oracle_accepts_strided_views(oracle_warp):
    base = torch.rand(2, 9, 10, 12, dtype=torch.float64)
    view = base.permute(0, 1, 3, 2)                     # non-contiguous
    assert torch.equal(oracle_warp.local_attn_reshape_forward(view, 3), F.pixel_shuffle(view.contiguous(), 3))


# ---------------------------------------------------------------- golden vectors
def _gold(tag):
    p = os.path.join(GOLD, "ref_cuda_%s.npz" % tag)
    if not os.path.exists(p):
        pytest.skip("golden vectors %s not recorded yet (tests/golden/make_golden_gpu.py)" % p)
    return np.load(p)


def _cmp(got, want, big, rtol):
    got = got.numpy()
    if big:
        got = got.reshape(-1)[::cases.SUBSAMPLE]
    assert got.shape == want.shape
    scale = max(1e-30, float(np.abs(want).max()))
    assert float(np.abs(got - want).max()) <= rtol * scale


@pytest.mark.parametrize("tag,dt,rtol", [("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-12)])
def test_oracle_matches_reference_cuda_golden(oracle_warp, tag, dt, rtol):
    g = _gold(tag)
    for name in cases.RESAMPLE2D_CASES:
        c = cases.resample2d_case(name)
        in1, in2, go = (torch.tensor(c[k], dtype=dt) for k in ("in1", "in2", "gout"))
        big = name.startswith("cfg1")
        _cmp(oracle_warp.resample2d_forward(in1, in2, c["ks"], c["dil"]), g["resample2d/%s/out" % name], big, rtol)
        g1, g2 = oracle_warp.resample2d_backward(in1, in2, go, c["ks"], c["dil"])
        _cmp(g1, g["resample2d/%s/gin1" % name], big, rtol * 5)
        _cmp(g2, g["resample2d/%s/gin2" % name], big, rtol * 5)
    for name in cases.BLOCK_EXTRACTOR_CASES:
        c = cases.block_extractor_case(name)
        s, f, go = (torch.tensor(c[k], dtype=dt) for k in ("src", "flow", "gout"))
        _cmp(oracle_warp.block_extractor_forward(s, f, c["k"]), g["block_extractor/%s/out" % name], False, rtol)
        gs, gf = oracle_warp.block_extractor_backward(s, f, go, c["k"])
        _cmp(gs, g["block_extractor/%s/gsrc" % name], False, rtol * 5)
        _cmp(gf, g["block_extractor/%s/gflow" % name], False, rtol * 5)
    for name in cases.LOCAL_ATTN_RESHAPE_CASES:
        c = cases.local_attn_reshape_case(name)
        x, go = torch.tensor(c["x"], dtype=dt), torch.tensor(c["gout"], dtype=dt)
        assert np.array_equal(oracle_warp.local_attn_reshape_forward(x, c["k"]).numpy(), g["local_attn_reshape/%s/out" % name])
        assert np.array_equal(oracle_warp.local_attn_reshape_backward(x, go, c["k"]).numpy(), g["local_attn_reshape/%s/gin" % name])
