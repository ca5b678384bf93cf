"""The formulas restated in ffwm_b200/csrc/mfm.cu (mfm_max / mfm_ga / mfm_gb) against PyTorch's own
torch.max(a, b) forward and autograd backward on the CPU, including ties (gradient split in half) and NaN.
The kernels were written without GPU access; this pins the semantics they transcribe."""
import numpy as np
import torch


def mfm_max(a, b):
    return np.where(a != a, a, np.where(b != b, b, np.maximum(a, b)))


def mfm_ga(a, b, g):
    return np.where(a < b, 0.0, np.where(a == b, g * 0.5, g))


def mfm_gb(a, b, g):
    return np.where(a > b, 0.0, np.where(a == b, g * 0.5, g))


def test_kernel_formulas_are_atens():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(4000, generator=g)
    b = torch.randn(4000, generator=g)
    b[::5] = a[::5]                                    # ties
    a[7], b[11], a[13], b[13] = float("nan"), float("nan"), float("nan"), float("nan")
    a[17], b[17] = float("inf"), float("inf")
    a[19], b[19] = 0.0, -0.0
    go = torch.randn(4000, generator=g)
    ar, br = a.clone().requires_grad_(), b.clone().requires_grad_()
    out = torch.max(ar, br)
    out.backward(go)
    an, bn, gn = a.numpy(), b.numpy(), go.numpy()
    with np.errstate(invalid="ignore"):
        np.testing.assert_array_equal(mfm_max(an, bn), out.detach().numpy())
        np.testing.assert_array_equal(mfm_ga(an, bn, gn).astype(np.float32), ar.grad.numpy())
        np.testing.assert_array_equal(mfm_gb(an, bn, gn).astype(np.float32), br.grad.numpy())


def test_source_holds_the_same_formulas():
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ffwm_b200", "csrc", "mfm.cu")).read()
    assert "return a != a ? a : (b != b ? b : fmaxf(a, b));" in src
    assert "return a < b ? 0.f : (a == b ? g * 0.5f : g);" in src
    assert "return a > b ? 0.f : (a == b ? g * 0.5f : g);" in src
