"""Fused guided-filter kernels (ffwm_b200/csrc/guided_filter.cu) against the torch-op mirror of the reference formula
in float64 (forward 1e-5, gradient 1e-4 relative to max|ref|: the path's stated tolerances).

Validated on a B200 in round 2 (gpurun call 1, profiles/r02a_*)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape,r", [((8, 3, 128, 128), 32), ((8, 3, 64, 64), 16), ((8, 3, 32, 32), 8), ((2, 3, 20, 37), 4),
                                     ((1, 1, 12, 40), 5)])
def test_fused_guided_filter_matches_reference_formula(shape, r):
    from ffwm_b200.external_function import GuidedFilter, GuidedFilterFunction
    g = torch.Generator().manual_seed(r)
    x, y, gq = torch.rand(*shape, generator=g), torch.rand(*shape, generator=g), torch.randn(*shape, generator=g)
    x64 = x.double().to(DEV).requires_grad_()
    q64 = GuidedFilter(r)(x64, y.double().to(DEV))
    q64.backward(gq.double().to(DEV))
    xd = x.to(DEV).requires_grad_()
    q = GuidedFilterFunction.apply(xd, y.to(DEV), r, 1e-8)
    q.backward(gq.to(DEV))
    torch.cuda.synchronize()
    assert rel(q, q64) <= 1e-5
    assert rel(xd.grad, x64.grad) <= 1e-4


def test_module_switches_to_the_fused_path(monkeypatch):
    from ffwm_b200 import external_function as E
    g = torch.Generator().manual_seed(1)
    x, y = torch.rand(2, 3, 64, 64, generator=g).to(DEV), torch.rand(2, 3, 64, 64, generator=g).to(DEV)
    want = E.GuidedFilter(16)(x, y)
    monkeypatch.setattr(E, "FUSED_GF", True)
    assert rel(E.GuidedFilter(16)(x, y), want) <= 1e-5
    with pytest.raises(AssertionError):                      # the reference's size assert still applies
        E.GuidedFilter(40)(x, y)
