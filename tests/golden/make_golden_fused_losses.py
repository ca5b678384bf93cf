"""Generates tests/golden/ref_fused_losses.npz with the REFERENCE's own loss modules (baseline/_ref, byte-identical
staged copy) on the CPU in float64:

  * `AffineRegularizationLoss(kz)(flow)` (models/losses.py:181-223) for kz = 3, 5, 7 — value and gradient w.r.t. the
    flow.  The reference has no CPU implementation of block_extractor / local_attn_reshape, so its instances get the
    C-oracle-backed callables (harness shim, as in baseline/ref_harness.py); nothing in the reference is edited.
  * `PerceptualCorrectness.calculate_loss` (:341-371) on seeded feature maps (the VGG pass is bypassed by setting
    `target_vgg` / `source_vgg` directly) — value, and the intermediate `correction_max`.

    python tests/golden/make_golden_fused_losses.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from baseline import ref_harness as H  # noqa: E402


def affine_inputs(kz, dtype=torch.float64):
    g = torch.Generator().manual_seed(40 + kz)
    s = {3: 32, 5: 40, 7: 37}[kz]
    lin = torch.linspace(-1, 1, s, dtype=torch.float64)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    base = torch.stack((gx, gy), 0).unsqueeze(0).repeat(2, 1, 1, 1)
    return (base + 0.05 * torch.randn(2, 2, s, s, generator=g, dtype=torch.float64)).to(dtype)


def corr_inputs(c, hw, dtype=torch.float64):
    g = torch.Generator().manual_seed(c + hw)
    src = torch.randn(2, c, hw, hw, generator=g, dtype=torch.float64).clamp_min(0)
    tgt = torch.randn(2, c, hw, hw, generator=g, dtype=torch.float64).clamp_min(0)
    lin = torch.linspace(-1, 1, hw, dtype=torch.float64)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    flow = torch.stack((gx, gy), 0).unsqueeze(0).repeat(2, 1, 1, 1) + 0.1 * torch.randn(2, 2, hw, hw, generator=g, dtype=torch.float64)
    mask = (torch.rand(2, 1, hw, hw, generator=g, dtype=torch.float64) > 0.3).double()
    return src.to(dtype), tgt.to(dtype), flow.to(dtype), mask.to(dtype)


CORR_CASES = ((64, 32), (128, 24), (256, 16), (64, 19))


def main():
    H.import_reference()
    import importlib
    RL = importlib.import_module("models.losses")
    from oracle import train_cpu

    class _Extract(torch.nn.Module):
        def __init__(self, kz):
            super().__init__()
            self.kz = kz

        def forward(self, s, f):
            return train_cpu._BlockExtractorCPU.apply(s, f, self.kz)

    class _Reshape(torch.nn.Module):
        def forward(self, x, k):
            return train_cpu._LocalAttnReshapeCPU.apply(x, k)

    out = {}
    for kz in (3, 5, 7):
        reg = RL.AffineRegularizationLoss(kz)
        reg.extractor, reg.reshape = _Extract(kz), _Reshape()
        flow = affine_inputs(kz).requires_grad_(True)
        loss = reg(flow)
        loss.backward()
        out["affine/kz%d/loss" % kz] = loss.detach().numpy()
        out["affine/kz%d/grad" % kz] = flow.grad.numpy()
    for c, hw in CORR_CASES:
        src, tgt, flow, mask = corr_inputs(c, hw)
        pc = object.__new__(RL.PerceptualCorrectness)
        torch.nn.Module.__init__(pc)
        pc.eps = 1e-8
        pc.target_vgg, pc.source_vgg = {"x": tgt}, {"x": src}
        out["corr/%d_%d/loss_masked" % (c, hw)] = pc.calculate_loss(flow, "x", mask, True).numpy()
        out["corr/%d_%d/loss" % (c, hw)] = pc.calculate_loss(flow, "x", None, True).numpy()
        # the intermediate the fused kernel replaces (models/losses.py:347-353)
        b = src.size(0)
        sa = src.view(b, c, -1).transpose(1, 2)
        ta = tgt.view(b, c, -1)
        sn = sa / (sa.norm(dim=2, keepdim=True) + 1e-8)
        tn = ta / (ta.norm(dim=1, keepdim=True) + 1e-8)
        out["corr/%d_%d/cmax" % (c, hw)] = torch.max(torch.bmm(sn, tn), dim=1)[0].numpy()
    np.savez_compressed(os.path.join(HERE, "ref_fused_losses.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
