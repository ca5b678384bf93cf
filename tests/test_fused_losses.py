"""the fused correlation column-max (csrc/corr_max.cu) and the fused affine regularisation
(csrc/affine_reg.cu) against golden vectors produced by the REFERENCE's own loss modules in float64
(tests/golden/make_golden_fused_losses.py -> ref_fused_losses.npz) and against the unfused chain of validated kernels.

Tolerances: affine regularisation fp64 1e-10, fp32 2e-5 of the loss / 1e-4 of max|grad| (the quadratic form cancels:
grid values are ~64, residuals ~1); correlation max 2e-5 absolute on cosine values in [0, 1] (3xBF16 split), loss 1e-4."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden_fused_losses as G  # noqa: E402  (input builders only; nothing of the reference is touched at import)

GOLD = np.load(os.path.join(HERE, "golden", "ref_fused_losses.npz"))
DEV = "cuda:0"


def test_golden_is_the_affine_fit_residual():
    """CPU: the reference's five-pass chain equals mean over windows of w^T Q w — the identity the fused kernel implements."""
    from ffwm_b200.losses import AffineRegularizationLoss
    for kz in (3, 5, 7):
        reg = AffineRegularizationLoss(kz)
        flow = G.affine_inputs(kz)
        grid = reg.flow2grid(flow)
        q = reg.kernel.reshape(kz * kz, kz * kz).double()
        total = 0.0
        for ch in range(2):
            win = torch.nn.functional.unfold(grid[:, ch:ch + 1], kz)                    # (b, kz^2, windows)
            total += torch.einsum("bmw,mn,bnw->bw", win, q, win).mean()
        assert abs(float(total) - float(GOLD["affine/kz%d/loss" % kz])) <= 1e-9 * abs(float(total))


@pytest.mark.gpu
@pytest.mark.parametrize("kz", [3, 5, 7])
@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_fused_affine_regularisation_matches_reference(kz, dt):
    from ffwm_b200 import _lib, losses
    reg = losses.AffineRegularizationLoss(kz)
    flow = G.affine_inputs(kz, dt).to(DEV).requires_grad_(True)
    n0 = _lib.kernel_launches()
    loss = reg(flow)
    loss.backward()
    assert _lib.kernel_launches() - n0 == 4                      # one kernel per plane and direction
    want, wgrad = float(GOLD["affine/kz%d/loss" % kz]), GOLD["affine/kz%d/grad" % kz]
    ltol, gtol = (1e-10, 1e-10) if dt == torch.float64 else (2e-5, 1e-4)
    assert abs(float(loss) - want) <= ltol * abs(want)
    assert np.abs(flow.grad.double().cpu().numpy() - wgrad).max() <= gtol * np.abs(wgrad).max()
    # and the unfused chain on the validated K4 / K6 kernels gives the same number
    flow2 = G.affine_inputs(kz, dt).to(DEV).requires_grad_(True)
    losses.FUSED_AFFINE = False
    try:
        loss2 = reg(flow2)
        loss2.backward()
    finally:
        losses.FUSED_AFFINE = True
    assert abs(float(loss2) - float(loss)) <= max(ltol, 1e-12) * abs(want) * 2
    assert float((flow2.grad - flow.grad).abs().max()) <= 2 * gtol * np.abs(wgrad).max() + 1e-12


@pytest.mark.gpu
def test_affine_reg_ragged_and_rejections():
    from ffwm_b200 import ops
    q = torch.eye(9, device=DEV)
    g = torch.randn(3, 1, 5, 11, device=DEV)
    out = torch.empty(3, 1, 3, 9, device=DEV)
    ops.affine_reg_forward(g, q, out, 3)                         # Q = I: sum of squares of the window / 9
    want = torch.nn.functional.unfold(g, 3).pow(2).sum(1).view(3, 1, 3, 9) / 9
    assert float((out - want).abs().max()) <= 1e-5
    with pytest.raises(RuntimeError):
        ops.affine_reg_forward(g, q, torch.empty(3, 1, 3, 8, device=DEV), 3)
    with pytest.raises(RuntimeError):
        ops.affine_reg_forward(g, torch.eye(16, device=DEV), torch.empty(3, 1, 2, 8, device=DEV), 4)


@pytest.mark.gpu
@pytest.mark.parametrize("c,hw", G.CORR_CASES)
def test_corr_max_matches_reference(c, hw):
    from ffwm_b200 import ops
    src, tgt, flow, mask = G.corr_inputs(c, hw, torch.float32)
    got = ops.corr_max(src.to(DEV), tgt.to(DEV), 1e-8)
    want = GOLD["corr/%d_%d/cmax" % (c, hw)]
    assert got.shape == want.shape
    assert np.abs(got.double().cpu().numpy() - want).max() <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("c,hw", G.CORR_CASES)
def test_perceptual_correctness_with_fused_column_max(c, hw):
    from ffwm_b200 import _lib, losses
    src, tgt, flow, mask = [t.to(DEV) for t in G.corr_inputs(c, hw, torch.float32)]
    pc = object.__new__(losses.PerceptualCorrectness)
    torch.nn.Module.__init__(pc)
    pc.eps = 1e-8
    pc.target_vgg, pc.source_vgg = {"x": tgt}, {"x": src}
    flow.requires_grad_(True)
    n0 = _lib.kernel_launches()
    lm = pc.calculate_loss(flow, "x", mask, True)
    assert _lib.kernel_launches() - n0 >= 3                      # prepass + GEMM/max + the grid warp
    lu = pc.calculate_loss(flow, "x", None, True)
    assert abs(float(lm) - float(GOLD["corr/%d_%d/loss_masked" % (c, hw)])) <= 1e-4 * abs(float(GOLD["corr/%d_%d/loss_masked" % (c, hw)]))
    assert abs(float(lu) - float(GOLD["corr/%d_%d/loss" % (c, hw)])) <= 1e-4 * abs(float(GOLD["corr/%d_%d/loss" % (c, hw)]))
    lm.backward()                                                # the flow gradient goes through the warp only
    assert torch.isfinite(flow.grad).all()


@pytest.mark.gpu
def test_corr_max_full_size_properties():
    """BASELINE size (relu1_1: C = 64, 128x128, 16384 x 16384 products per sample): checked through properties —
    a target pixel that also occurs among the sources reaches 1; permuting the sources changes nothing."""
    from ffwm_b200 import ops
    g = torch.Generator().manual_seed(5)
    src = torch.randn(2, 64, 128, 128, generator=g).clamp_min(0).to(DEV)
    tgt = torch.randn(2, 64, 128, 128, generator=g).clamp_min(0).to(DEV)
    tgt[:, :, 5, :] = src[:, :, 77, :] * 3.0                      # same directions, other norms
    a = ops.corr_max(src, tgt, 1e-8)
    assert float((a.view(2, 128, 128)[:, 5] - 1).abs().max()) <= 2e-5
    assert float(a.max()) <= 1 + 2e-5 and float(a.min()) >= 0
    perm = torch.randperm(128 * 128, generator=g).to(DEV)
    b = ops.corr_max(src.view(2, 64, -1)[:, :, perm].view_as(src).contiguous(), tgt, 1e-8)
    assert float((a - b).abs().max()) <= 1e-6                     # max over a permuted set (accumulation order per pair is unchanged)
    # a row block against torch in float64
    sn = src[0].view(64, -1).double()
    tn = tgt[0].view(64, -1)[:, :512].double()
    sn = sn / (sn.norm(dim=0, keepdim=True) + 1e-8)
    tn = tn / (tn.norm(dim=0, keepdim=True) + 1e-8)
    want = (sn.t() @ tn).max(dim=0)[0]
    assert float((a[0, :512].double() - want).abs().max()) <= 2e-5
