"""GPU parity tests: the sm_100a kernels (through the C ABI) against the C
oracle on the same seeded inputs, against the reference's own CUDA extensions
(oracle/_ref) where they were prebuilt, and through size-independent
properties at benchmark sizes.

Tolerances (stated per SURVEY 8d / north_star):
  fp32  forward 1e-5, backward 1e-4 (scatter order), relative to max|ref|
  fp64  1e-12 forward, 1e-11 backward
  gather/index ops (local_attn_reshape, block_extractor with integer flow): bit-exact
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch.autograd import gradcheck

from tests.golden import cases

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
TOL = {torch.float32: (1e-5, 1e-4), torch.float64: (1e-12, 1e-11)}


def rel_err(got, want):
    want = want.to(got.device)
    scale = max(1e-30, float(want.abs().max()))
    return float((got - want).abs().max()) / scale


def to_dev(c, keys, dt):
    return [torch.tensor(c[k], dtype=dt, device=DEV) for k in keys]


@pytest.fixture(scope="module")
def E():
    import ffwm_b200  # noqa: F401  loads libffwm_b200.so or raises
    from ffwm_b200 import external_function
    return external_function


@pytest.fixture(scope="module")
def ops():
    from ffwm_b200 import ops
    return ops


# ----------------------------------------------------------------- resample2d
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", cases.RESAMPLE2D_CASES)
def test_resample2d_vs_oracle(ops, oracle_warp, name, dt):
    c = cases.resample2d_case(name)
    in1, in2, go = to_dev(c, ("in1", "in2", "gout"), dt)
    out = torch.empty_like(go)
    ops.resample2d_forward(in1, in2, out, c["ks"], c["dil"])
    g1, g2 = torch.zeros_like(in1), torch.full_like(in2, float("nan"))
    ops.resample2d_backward(in1, in2, go, g1, g2, c["ks"], c["dil"])
    ref_out = oracle_warp.resample2d_forward(in1.cpu(), in2.cpu(), c["ks"], c["dil"])
    r1, r2 = oracle_warp.resample2d_backward(in1.cpu(), in2.cpu(), go.cpu(), c["ks"], c["dil"])
    ft, bt = TOL[dt]
    assert rel_err(out, ref_out) <= ft
    assert rel_err(g1, r1) <= bt
    assert rel_err(g2, r2) <= bt


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_resample2d_vs_reference_cuda(ops, ref_cuda, dt):
    rs = ref_cuda["resample2d_cuda"]
    ft, bt = TOL[dt]
    for name in cases.RESAMPLE2D_CASES:
        c = cases.resample2d_case(name)
        in1, in2, go = to_dev(c, ("in1", "in2", "gout"), dt)
        out, ref = torch.empty_like(go), torch.zeros_like(go)
        ops.resample2d_forward(in1, in2, out, c["ks"], c["dil"])
        rs.forward(in1, in2, ref, c["ks"], c["dil"])
        g1, g2 = torch.zeros_like(in1), torch.empty_like(in2)
        r1, r2 = torch.zeros_like(in1), torch.zeros_like(in2)
        ops.resample2d_backward(in1, in2, go, g1, g2, c["ks"], c["dil"])
        rs.backward(in1, in2, go, r1, r2, c["ks"], c["dil"])
        assert rel_err(out, ref) <= ft, name
        assert rel_err(g1, r1) <= bt, name
        assert rel_err(g2, r2) <= bt, name


def test_resample2d_large_kernel_generic_path(ops, oracle_warp):
    torch.manual_seed(3)
    in1 = torch.rand(1, 2, 20, 20, dtype=torch.float64, device=DEV)
    in2 = torch.cat([torch.randn(1, 2, 9, 9, dtype=torch.float64) * 2, torch.full((1, 1, 9, 9), 3.0, dtype=torch.float64)], 1).to(DEV)
    go = torch.randn(1, 2, 9, 9, dtype=torch.float64, device=DEV)
    for ks in (10, 0, 1):
        out = torch.empty_like(go)
        ops.resample2d_forward(in1, in2, out, ks, 1)
        assert rel_err(out, oracle_warp.resample2d_forward(in1.cpu(), in2.cpu(), ks, 1)) <= 1e-12 or ks < 2
        g1, g2 = torch.zeros_like(in1), torch.empty_like(in2)
        ops.resample2d_backward(in1, in2, go, g1, g2, ks, 1)
        if ks >= 2:
            r1, r2 = oracle_warp.resample2d_backward(in1.cpu(), in2.cpu(), go.cpu(), ks, 1)
            assert rel_err(g1, r1) <= 1e-11 and rel_err(g2, r2) <= 1e-11
        else:
            assert torch.count_nonzero(out) == 0


def test_resample2d_module_autograd(E):
    torch.manual_seed(0)
    a = torch.rand(2, 3, 9, 8, dtype=torch.float64, device=DEV, requires_grad=True)
    fl = (torch.rand(2, 2, 9, 8, dtype=torch.float64, device=DEV) * 0.6 + 0.2).requires_grad_()
    for ks in (2, 4):
        mod = E.Resample2d(ks, 1, sigma=2)
        assert gradcheck(lambda u, v: mod(u, v), (a, fl), eps=1e-6, atol=1e-5, nondet_tol=1e-10)


def test_resample2d_model_shapes_and_partial_grads(ops, oracle_warp):
    # PerceptualCorrectness would call it with (b,c,h,w) VGG maps, ks=4 sigma=2 (losses.py:329)
    torch.manual_seed(5)
    in1 = torch.rand(2, 64, 32, 32, device=DEV)
    in2 = torch.cat([torch.randn(2, 2, 32, 32) * 2, torch.full((2, 1, 32, 32), 2.0)], 1).to(DEV)
    go = torch.randn(2, 64, 32, 32, device=DEV)
    r1, r2 = oracle_warp.resample2d_backward(in1.cpu(), in2.cpu(), go.cpu(), 4, 1)
    g1 = torch.zeros_like(in1)
    ops.resample2d_backward(in1, in2, go, g1, None, 4, 1)       # only grad_input1
    assert rel_err(g1, r1) <= 1e-4
    g2 = torch.empty_like(in2)
    ops.resample2d_backward(in1, in2, go, None, g2, 4, 1)       # only grad_input2
    assert rel_err(g2, r2) <= 1e-4


# ------------------------------------------------------------ block_extractor
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", cases.BLOCK_EXTRACTOR_CASES)
def test_block_extractor_vs_oracle(ops, oracle_warp, name, dt):
    c = cases.block_extractor_case(name)
    s, f, go = to_dev(c, ("src", "flow", "gout"), dt)
    out = torch.empty_like(go)
    ops.block_extractor_forward(s, f, out, c["k"])
    gs, gf = torch.zeros_like(s), torch.full_like(f, float("nan"))
    ops.block_extractor_backward(s, f, go, gs, gf, c["k"])
    ref = oracle_warp.block_extractor_forward(s.cpu(), f.cpu(), c["k"])
    rs_, rf = oracle_warp.block_extractor_backward(s.cpu(), f.cpu(), go.cpu(), c["k"])
    ft, bt = TOL[dt]
    assert rel_err(out, ref) <= ft
    assert rel_err(gs, rs_) <= bt
    assert rel_err(gf, rf) <= bt


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_block_extractor_vs_reference_cuda(ops, ref_cuda, dt):
    be = ref_cuda["block_extractor_cuda"]
    ft, bt = TOL[dt]
    for name in cases.BLOCK_EXTRACTOR_CASES:
        c = cases.block_extractor_case(name)
        s, f, go = to_dev(c, ("src", "flow", "gout"), dt)
        out, ref = torch.empty_like(go), torch.zeros_like(go)
        ops.block_extractor_forward(s, f, out, c["k"])
        be.forward(s, f, ref, c["k"])
        gs, gf = torch.zeros_like(s), torch.empty_like(f)
        rs_, rf = torch.zeros_like(s), torch.zeros_like(f)
        ops.block_extractor_backward(s, f, go, gs, gf, c["k"])
        be.backward(s, f, go, rs_, rf, c["k"])
        assert rel_err(out, ref) <= ft, name
        assert rel_err(gs, rs_) <= bt, name
        assert rel_err(gf, rf) <= bt, name


@pytest.mark.parametrize("S,k", [(32, 3), (64, 5), (128, 7)])
def test_block_extractor_live_use_is_bit_exact_unfold(E, S, k):
    # train_flow.py: 1-channel coordinate grids, flow == k//2 (losses.py:212-217), B=6
    torch.manual_seed(k)
    grid = torch.rand(6, 1, S, S, device=DEV) * 128
    hf = S - k + 1
    f = torch.zeros(6, 2, hf, hf, device=DEV) + float(k // 2)
    out = E.BlockExtractor(k)(grid, f)
    unf = F.unfold(grid, k).view(6, 1, k, k, hf, hf).permute(0, 1, 4, 2, 5, 3).reshape(6, 1, hf * k, hf * k)
    assert torch.equal(out, unf)


def test_block_extractor_gradcheck_reference_recipe(E):
    torch.manual_seed(0)
    s = torch.rand(4, 6, 14, 10, dtype=torch.float64, device=DEV, requires_grad=True)
    f = (torch.rand(4, 2, 14, 10, dtype=torch.float64, device=DEV) * 1.8).requires_grad_()
    ext = E.BlockExtractor(3)
    assert gradcheck(ext, (s, f), fast_mode=True, nondet_tol=1e-10)   # grad_source is a RED.ADD scatter
    s2 = s.detach()[:1, :2, :6, :5].clone().requires_grad_()
    f2 = f.detach()[:1, :, :6, :5].clone().requires_grad_()
    assert gradcheck(ext, (s2, f2), nondet_tol=1e-10)


def test_block_extractor_grad_flow_is_deterministic(ops):
    torch.manual_seed(1)
    s = torch.rand(2, 64, 32, 32, device=DEV)
    f = torch.rand(2, 2, 32, 32, device=DEV) * 1.8
    go = torch.randn(2, 64, 96, 96, device=DEV)
    outs = []
    for _ in range(3):
        gf = torch.empty_like(f)
        ops.block_extractor_backward(s, f, go, None, gf, 3)
        outs.append(gf)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


# --------------------------------------------------------- local_attn_reshape
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("name", cases.LOCAL_ATTN_RESHAPE_CASES)
def test_local_attn_reshape_bit_exact(E, oracle_warp, name, dt):
    c = cases.local_attn_reshape_case(name)
    x, go = to_dev(c, ("x", "gout"), dt)
    x.requires_grad_()
    out = E.LocalAttnReshape()(x, c["k"])
    out.backward(go)
    assert torch.equal(out.cpu(), oracle_warp.local_attn_reshape_forward(x.detach().cpu(), c["k"]))
    assert torch.equal(x.grad.cpu(), oracle_warp.local_attn_reshape_backward(x.detach().cpu(), go.cpu(), c["k"]))
    assert torch.equal(out, F.pixel_shuffle(x.detach(), c["k"]))
    if name == "kat_0_8":       # the reference's printed known answer
        assert out[0, 0, :3, :3].tolist() == [[0, 1, 2], [3, 4, 5], [6, 7, 8]]


def test_local_attn_reshape_vs_reference_cuda(ops, ref_cuda):
    lar = ref_cuda["local_attn_reshape_cuda"]
    for name in cases.LOCAL_ATTN_RESHAPE_CASES:
        c = cases.local_attn_reshape_case(name)
        x, go = to_dev(c, ("x", "gout"), torch.float32)
        out, ref = torch.empty_like(go), torch.zeros_like(go)
        ops.local_attn_reshape_forward(x, out, c["k"])
        lar.forward(x, ref, c["k"])
        gi, ri = torch.empty_like(x), torch.zeros_like(x)
        ops.local_attn_reshape_backward(go, gi, c["k"])
        lar.backward(x, go, ri, c["k"])
        assert torch.equal(out, ref) and torch.equal(gi, ri), name


@pytest.mark.parametrize("shape,k", [((6, 9, 30, 30), 3), ((6, 25, 60, 60), 5), ((6, 49, 122, 122), 7)])
def test_local_attn_reshape_live_shapes_round_trip(E, shape, k):
    # forward then backward of the same data is the identity (each address exactly once)
    x = torch.randn(*shape, device=DEV, requires_grad=True)
    out = E.LocalAttnReshape()(x, k)
    out.backward(out.detach())
    assert torch.equal(x.grad, x.detach())


def test_local_attn_reshape_gradcheck(E):
    x = torch.rand(4, 9, 14, 10, dtype=torch.float64, device=DEV, requires_grad=True)
    assert gradcheck(lambda t: E.LocalAttnReshape()(t, 3), (x,), fast_mode=True)


# ------------------------------------------------------------------ grid_warp
@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(2, 5, 13, 11, 7, 9), (8, 3, 128, 128, 32, 32), (8, 128, 32, 32, 32, 32),
                                   (2, 64, 64, 64, 64, 64), (1, 1, 4, 4, 5, 3)])
def test_grid_warp_vs_oracle_and_torch(E, oracle_warp, shape, dt):
    b, c, hi, wi, ho, wo = shape
    torch.manual_seed(hi + c)
    img = torch.rand(b, c, hi, wi, dtype=dt, device=DEV, requires_grad=True)
    fl = (torch.rand(b, 2, ho, wo, dtype=dt, device=DEV) * 2.6 - 1.3).requires_grad_()
    out = E.grid_warp(img, fl)
    go = torch.randn_like(out)
    out.backward(go)
    ft, bt = TOL[dt]
    ref = oracle_warp.grid_warp_forward(img.detach().cpu(), fl.detach().cpu())
    ri, rf = oracle_warp.grid_warp_backward(img.detach().cpu(), fl.detach().cpu(), go.cpu())
    assert rel_err(out, ref) <= ft
    assert rel_err(img.grad, ri) <= bt
    assert rel_err(fl.grad, rf) <= bt
    # and against the library call the reference makes on this GPU
    img2, fl2 = img.detach().clone().requires_grad_(), fl.detach().clone().requires_grad_()
    tref = F.grid_sample(img2, fl2.permute(0, 2, 3, 1), mode="bilinear", align_corners=False)
    tref.backward(go)
    assert rel_err(out, tref) <= ft
    assert rel_err(img.grad, img2.grad) <= bt
    assert rel_err(fl.grad, fl2.grad) <= bt


def test_grid_warp_gradcheck_and_identity(E):
    torch.manual_seed(0)
    img = torch.rand(2, 3, 7, 6, dtype=torch.float64, device=DEV, requires_grad=True)
    fl = (torch.rand(2, 2, 5, 4, dtype=torch.float64, device=DEV) * 1.6 - 0.8).requires_grad_()
    assert gradcheck(E.grid_warp, (img, fl), eps=1e-6, atol=1e-5, nondet_tol=1e-10)
    # identity grid reproduces the image
    h, w = 16, 12
    ys = (torch.arange(h, dtype=torch.float64, device=DEV) * 2 + 1) / h - 1
    xs = (torch.arange(w, dtype=torch.float64, device=DEV) * 2 + 1) / w - 1
    grid = torch.stack([xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)]).unsqueeze(0)
    im = torch.rand(1, 4, h, w, dtype=torch.float64, device=DEV)
    assert torch.allclose(E.grid_warp(im, grid), im, atol=1e-12)


# -------------------------------------------------- strides, empties, drop-in
def test_kernels_honour_strides(ops, oracle_warp):
    torch.manual_seed(2)
    base = torch.rand(2, 10, 12, 6, device=DEV)
    src = base.permute(0, 3, 1, 2)                       # (2,6,10,12) channels-last view, non-contiguous
    flow = (torch.rand(2, 10, 12, 2, device=DEV) * 1.8).permute(0, 3, 1, 2)
    out_store = torch.empty(2, 30, 36, 6, device=DEV)
    out = out_store.permute(0, 3, 1, 2)
    ops.block_extractor_forward(src, flow, out, 3)
    ref = oracle_warp.block_extractor_forward(src.cpu(), flow.cpu(), 3)
    assert rel_err(out, ref) <= 1e-5
    go = torch.randn(2, 30, 36, 6, device=DEV).permute(0, 3, 1, 2)   # the reference passes non-contiguous grads through
    gs, gf = torch.zeros_like(src), torch.empty_like(flow)
    ops.block_extractor_backward(src, flow, go, gs, gf, 3)
    rs_, rf = oracle_warp.block_extractor_backward(src.cpu(), flow.cpu(), go.cpu(), 3)
    assert rel_err(gs, rs_) <= 1e-4 and rel_err(gf, rf) <= 1e-4
    img_out = torch.empty(2, 10, 12, 6, device=DEV).permute(0, 3, 1, 2)
    g = (torch.rand(2, 2, 10, 12, device=DEV) * 2 - 1)
    ops.grid_warp_forward(src, g, img_out)
    assert rel_err(img_out, oracle_warp.grid_warp_forward(src.cpu(), g.cpu())) <= 1e-5


def test_empty_and_ragged_inputs(ops, E):
    e = torch.empty(0, 3, 8, 8, device=DEV)
    ops.grid_warp_forward(e, torch.empty(0, 2, 4, 4, device=DEV), torch.empty(0, 3, 4, 4, device=DEV))
    ops.resample2d_forward(e, torch.empty(0, 3, 8, 8, device=DEV), e.clone(), 2, 1)
    out = E.BlockExtractor(3)(torch.rand(1, 1, 1, 1, device=DEV), torch.zeros(1, 2, 1, 1, device=DEV))
    assert out.shape == (1, 1, 3, 3)
    with pytest.raises(RuntimeError):
        ops.block_extractor_forward(torch.rand(1, 1, 4, 4, device=DEV), torch.zeros(1, 2, 4, 4, device=DEV),
                                    torch.empty(1, 1, 11, 12, device=DEV), 3)


def test_dropin_modules_match_reference_calling_convention(ref_cuda):
    import ffwm_b200
    ffwm_b200.install_dropin()
    import block_extractor_cuda as mine
    theirs = ref_cuda["block_extractor_cuda"]
    torch.manual_seed(4)
    s = torch.rand(2, 3, 9, 9, device=DEV)
    f = torch.rand(2, 2, 9, 9, device=DEV) * 1.8
    a, b = torch.zeros(2, 3, 27, 27, device=DEV), torch.zeros(2, 3, 27, 27, device=DEV)
    assert mine.forward(s, f, a, 3) == 1 and theirs.forward(s, f, b, 3) == 1
    assert rel_err(a, b) <= 1e-6


# --------------------------------- benchmark sizes through size-independent properties
def test_full_size_properties(ops):
    torch.manual_seed(0)
    b, c, r = 4, 64, 256
    src = torch.rand(b, c, r, r, device=DEV)
    # zero displacement, any sigma: resample2d(ks=2) keeps pixels whose 4 taps are all the
    # same clamped pixel only at the border; in general it is a normalised average, so it is
    # bounded by min/max of the source and linear in the source.
    in2 = torch.cat([torch.randn(b, 2, r, r, device=DEV) * 2, torch.full((b, 1, r, r), 2.0, device=DEV)], 1)
    o1, o2, o3 = (torch.empty_like(src) for _ in range(3))
    ops.resample2d_forward(src, in2, o1, 4, 1)
    assert float(o1.min()) >= 0.0 and float(o1.max()) <= 1.0
    src2 = torch.rand_like(src)
    ops.resample2d_forward(src2, in2, o2, 4, 1)
    ops.resample2d_forward(src + 2 * src2, in2, o3, 4, 1)
    assert rel_err(o3, o1 + 2 * o2) <= 1e-5                       # linearity in input1
    # adjoint identity: <resample(a), g> == <a, grad_input1(g)>.  It only holds where the
    # reference's backward is consistent with its forward, i.e. for non-negative sampling
    # coordinates (K2 truncates where K1 floors), so displacements are made >= 0.
    in2p = torch.cat([in2[:, :2].abs(), in2[:, 2:]], 1)
    ops.resample2d_forward(src, in2p, o1, 4, 1)
    g = torch.randn_like(src)
    g1 = torch.zeros_like(src)
    ops.resample2d_backward(src, in2p, g, g1, None, 4, 1)
    lhs, rhs = float((o1.double() * g.double()).sum()), float((src.double() * g1.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)
    # block_extractor k=3 with zero flow: block centres reproduce the source bit-exactly
    flow = torch.zeros(b, 2, r, r, device=DEV)
    ob = torch.empty(b, c, 3 * r, 3 * r, device=DEV)
    ops.block_extractor_forward(src, flow, ob, 3)
    assert torch.equal(ob[:, :, 1::3, 1::3], src)
    # grid_warp adjoint identity
    gw = torch.rand(b, 2, r, r, device=DEV) * 2 - 1
    ow = torch.empty_like(src)
    ops.grid_warp_forward(src, gw, ow)
    gi = torch.zeros_like(src)
    ops.grid_warp_backward(src, gw, g, gi, None)
    lhs, rhs = float((ow.double() * g.double()).sum()), float((src.double() * gi.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


# --------------------------------- tiled scatter (large maps) vs the direct kernels and the oracle
@pytest.mark.gpu
@pytest.mark.parametrize("ks,fscale", [(2, 2.0), (4, 2.0), (4, 12.0)])
def test_resample2d_tiled_scatter_matches_direct_and_oracle(ops, oracle_warp, ks, fscale):
    """(2,48,100,90): ragged tiles, 32+16 channel groups; fscale=12 px pushes most taps out of the
    tile's halo (far list) and against the image border (clamped duplicates)."""
    g = torch.Generator().manual_seed(ks * 10 + int(fscale))
    b, c, h, w = 2, 48, 100, 90
    in1 = torch.rand(b, c, h, w, generator=g) * 2 - 1
    in2 = torch.cat([torch.randn(b, 2, h, w, generator=g) * fscale, torch.rand(b, 1, h, w, generator=g) * 2 + 1], 1)
    go = torch.randn(b, c, h, w, generator=g)
    d = [t.to(DEV) for t in (in1, in2, go)]
    g1_t, g2_t = torch.zeros_like(d[0]), torch.empty_like(d[1])
    ops.resample2d_backward(d[0], d[1], d[2], g1_t, g2_t, ks, 1)
    with _env(FFWM_DISABLE_TILED="1"):
        g1_d, g2_d = torch.zeros_like(d[0]), torch.empty_like(d[1])
        ops.resample2d_backward(d[0], d[1], d[2], g1_d, g2_d, ks, 1)
    assert rel_err(g1_t, g1_d) <= 2e-5
    assert rel_err(g2_t, g2_d) <= 5e-5                   # tiled gather vs direct kernel: summation order only
    out_t, out_d = torch.empty_like(d[0]), torch.empty_like(d[0])
    ops.resample2d_forward(d[0], d[1], out_t, ks, 1)
    with _env(FFWM_DISABLE_TILED="1"):
        ops.resample2d_forward(d[0], d[1], out_d, ks, 1)
    assert rel_err(out_t, out_d) <= 1e-6
    assert rel_err(out_t.cpu(), oracle_warp.resample2d_forward(in1, in2, ks, 1)) <= 1e-5
    w1, w2 = oracle_warp.resample2d_backward(in1, in2, go, ks, 1)
    assert rel_err(g1_t.cpu(), w1) <= 1e-4
    assert rel_err(g2_t.cpu()[:, :2], w2[:, :2]) <= 1e-4
    # accumulate semantics: a second call adds on top
    ops.resample2d_backward(d[0], d[1], d[2], g1_t, None, ks, 1)
    assert rel_err(g1_t, 2 * g1_d) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("noise", [0.03, 0.6])
def test_grid_warp_tiled_scatter_matches_direct_and_torch(ops, noise):
    g = torch.Generator().manual_seed(7)
    b, c, h, w = 2, 40, 96, 112
    img = torch.rand(b, c, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    grid = torch.stack((xs, ys), 0).unsqueeze(0).repeat(b, 1, 1, 1) + noise * torch.randn(b, 2, h, w, generator=g)
    go = torch.randn(b, c, h, w, generator=g)
    d = [t.to(DEV) for t in (img, grid, go)]
    gi_t, gf_t = torch.zeros_like(d[0]), torch.empty_like(d[1])
    ops.grid_warp_backward(d[0], d[1], d[2], gi_t, gf_t)
    with _env(FFWM_DISABLE_TILED="1"):
        gi_d, gf_d = torch.zeros_like(d[0]), torch.empty_like(d[1])
        ops.grid_warp_backward(d[0], d[1], d[2], gi_d, gf_d)
    assert rel_err(gi_t, gi_d) <= 2e-5 and rel_err(gf_t, gf_d) <= 5e-5
    out_t, out_d = torch.empty_like(d[0]), torch.empty_like(d[0])
    ops.grid_warp_forward(d[0], d[1], out_t)
    ops.grid_warp_forward(d[0], d[1], out_d)
    assert torch.equal(out_t, out_d)                      # deterministic
    x = img.double().requires_grad_(True)
    gr = grid.double().requires_grad_(True)
    torch.nn.functional.grid_sample(x, gr.permute(0, 2, 3, 1), align_corners=False).backward(go.double())
    assert rel_err(gi_t.cpu().double(), x.grad) <= 1e-5
    assert rel_err(gf_t.cpu().double(), gr.grad) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("k,mode", [(3, "rand1.8"), (3, "wild"), (2, "rand1.8"), (5, "rand1.8")])
def test_block_extractor_tiled_paths_match_direct_and_oracle(ops, oracle_warp, k, mode):
    """Large enough for the tiled forward (k=2,3) and the tiled grad_source scatter."""
    g = torch.Generator().manual_seed(31 + k)
    b, c, h, w = 2, 24, 40, 37
    src = torch.rand(b, c, h, w, generator=g)
    flow = torch.rand(b, 2, h, w, generator=g) * 1.8 if mode == "rand1.8" else torch.randn(b, 2, h, w, generator=g) * 9
    go = torch.randn(b, c, k * h, k * w, generator=g)
    d = [t.to(DEV) for t in (src, flow, go)]
    out = torch.empty_like(d[2])
    ops.block_extractor_forward(d[0], d[1], out, k)
    assert rel_err(out.cpu(), oracle_warp.block_extractor_forward(src, flow, k)) <= 1e-6
    gs_t, gf_t = torch.zeros_like(d[0]), torch.empty_like(d[1])
    ops.block_extractor_backward(d[0], d[1], d[2], gs_t, gf_t, k)
    with _env(FFWM_DISABLE_TILED="1"):
        gs_d, gf_d = torch.zeros_like(d[0]), torch.empty_like(d[1])
        ops.block_extractor_backward(d[0], d[1], d[2], gs_d, gf_d, k)
    assert rel_err(gs_t, gs_d) <= 2e-5 and rel_err(gf_t, gf_d) <= 5e-5
    ws, wf = oracle_warp.block_extractor_backward(src, flow, go, k)
    assert rel_err(gs_t.cpu(), ws) <= 1e-4 and rel_err(gf_t.cpu(), wf) <= 1e-4


# --------------------------------- tiled kernel generations on small ragged shapes (forced) vs direct kernels / oracle
class _env:
    """Runtime options of the library for the duration of a block (ffwm_set_option; the library reads the FFWM_*
    environment only once, at load).  _env(FFWM_FORCE_TILED="1") -> option FORCE_TILED = 1, restored on exit."""

    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        from ffwm_b200 import _lib
        self.old = {k: _lib.set_option(k, 0 if v is None else int(v)) for k, v in self.kv.items()}

    def __exit__(self, *a):
        from ffwm_b200 import _lib
        for k, v in self.old.items():
            _lib.set_option(k, v)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 48, 100, 90), (1, 32, 33, 16), (3, 17, 40, 21)])
@pytest.mark.parametrize("ks,fscale", [(2, 2.0), (4, 2.0), (4, 12.0)])
def test_resample2d_forced_tiled_matches_direct_and_oracle(ops, oracle_warp, ks, fscale, shape):
    """FFWM_FORCE_TILED sends small ragged shapes (partial strips / tiles, partial channel groups, heights that
    are not a multiple of the 8-row step) through the rolling forward (roll_gather.cuh), the row-owner scatter
    (scatter_rows.cuh) and the accumulate-then-weigh flow gradient (gather_quad.cuh); fscale=12 px pushes most
    windows out of the staged region (far paths) and against the border (padding = the reference's clamps).
    These kernels use expf instead of the double exp: 1e-6-level differences from the direct kernels."""
    g = torch.Generator().manual_seed(ks * 10 + int(fscale) + shape[1])
    b, c, h, w = shape
    in1 = torch.rand(b, c, h, w, generator=g) * 2 - 1
    in2 = torch.cat([torch.randn(b, 2, h, w, generator=g) * fscale, torch.rand(b, 1, h, w, generator=g) * 2 + 1], 1)
    go = torch.randn(b, c, h, w, generator=g)
    d = [t.to(DEV) for t in (in1, in2, go)]
    out_r, out_d = torch.empty_like(d[0]), torch.empty_like(d[0])
    g1_r, g1_d = torch.zeros_like(d[0]), torch.zeros_like(d[0])
    g2_r, g2_d = torch.empty_like(d[1]), torch.empty_like(d[1])
    with _env(FFWM_FORCE_TILED="1"):
        ops.resample2d_forward(d[0], d[1], out_r, ks, 1)
        ops.resample2d_backward(d[0], d[1], d[2], g1_r, g2_r, ks, 1)
    with _env(FFWM_DISABLE_TILED="1"):
        ops.resample2d_forward(d[0], d[1], out_d, ks, 1)
        ops.resample2d_backward(d[0], d[1], d[2], g1_d, g2_d, ks, 1)
    assert rel_err(out_r, out_d) <= 3e-6
    assert rel_err(g1_r, g1_d) <= 2e-5
    assert rel_err(g2_r, g2_d) <= 5e-5
    assert rel_err(out_r.cpu(), oracle_warp.resample2d_forward(in1, in2, ks, 1)) <= 1e-5
    w1, w2 = oracle_warp.resample2d_backward(in1, in2, go, ks, 1)
    assert rel_err(g1_r.cpu(), w1) <= 1e-4
    assert rel_err(g2_r.cpu(), w2) <= 1e-4
    # a strided (channels-last) grad_input1 takes the scalar flush of the scatter
    g1_s = torch.zeros(b, h, w, c, device=DEV).permute(0, 3, 1, 2)
    with _env(FFWM_FORCE_TILED="1"):
        ops.resample2d_backward(d[0], d[1], d[2], g1_s, None, ks, 1)
    assert rel_err(g1_s, g1_d) <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 40, 96, 112), (1, 33, 37, 50), (2, 16, 20, 24)])
@pytest.mark.parametrize("noise", [0.03, 0.6])
def test_grid_warp_forced_tiled_matches_direct_and_torch(ops, noise, shape):
    g = torch.Generator().manual_seed(7 + shape[1])
    b, c, h, w = shape
    img = torch.rand(b, c, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), indexing="ij")
    grid = torch.stack((xs, ys), 0).unsqueeze(0).repeat(b, 1, 1, 1) + noise * torch.randn(b, 2, h, w, generator=g)
    go = torch.randn(b, c, h, w, generator=g)
    d = [t.to(DEV) for t in (img, grid, go)]
    gi_r, gi_d = torch.zeros_like(d[0]), torch.zeros_like(d[0])
    gf_r, gf_d = torch.empty_like(d[1]), torch.empty_like(d[1])
    with _env(FFWM_FORCE_TILED="1"):
        ops.grid_warp_backward(d[0], d[1], d[2], gi_r, gf_r)
    with _env(FFWM_DISABLE_TILED="1"):
        ops.grid_warp_backward(d[0], d[1], d[2], gi_d, gf_d)
    assert rel_err(gi_r, gi_d) <= 2e-5
    assert rel_err(gf_r, gf_d) <= 5e-5
    x = img.double().requires_grad_(True)
    gr = grid.double().requires_grad_(True)
    torch.nn.functional.grid_sample(x, gr.permute(0, 2, 3, 1), align_corners=False).backward(go.double())
    assert rel_err(gi_r.cpu().double(), x.grad) <= 1e-5
    assert rel_err(gf_r.cpu().double(), gr.grad) <= 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 24, 40, 37), (1, 40, 33, 64)])
@pytest.mark.parametrize("k,mode", [(3, "rand1.8"), (3, "randn2"), (3, "wild"), (2, "rand1.8"), (2, "wild")])
def test_block_extractor_window_forward_matches_oracle(ops, oracle_warp, k, mode, shape):
    g = torch.Generator().manual_seed(31 + k + shape[1])
    b, c, h, w = shape
    src = torch.rand(b, c, h, w, generator=g)
    flow = {"rand1.8": lambda: torch.rand(b, 2, h, w, generator=g) * 1.8,
            "randn2": lambda: torch.randn(b, 2, h, w, generator=g) * 2,
            "wild": lambda: torch.randn(b, 2, h, w, generator=g) * 9}[mode]()
    d = [t.to(DEV) for t in (src, flow)]
    out_r = torch.empty(b, c, k * h, k * w, device=DEV)
    ops.block_extractor_forward(d[0], d[1], out_r, k)
    assert rel_err(out_r.cpu(), oracle_warp.block_extractor_forward(src, flow, k)) <= 1e-6
    # integer flow: bit-exact unfold (the live use in the reference)
    iflow = torch.full((b, 2, h, w), float(k // 2), device=DEV)
    ops.block_extractor_forward(d[0], iflow, out_r, k)
    want = F.unfold(F.pad(d[0], (0, k - 1, 0, k - 1), mode="replicate"), k).view(b, c, k, k, h, w)   # taps x+j clamp at the edge
    want = want.permute(0, 1, 4, 2, 5, 3).reshape(b, c, k * h, k * w)
    assert torch.equal(out_r, want)
    # a strided (channels-last) output
    out_s = torch.empty(b, k * h, k * w, c, device=DEV).permute(0, 3, 1, 2)
    ops.block_extractor_forward(d[0], d[1], out_s, k)
    assert rel_err(out_s.cpu(), oracle_warp.block_extractor_forward(src, flow, k)) <= 1e-6
