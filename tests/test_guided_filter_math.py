"""The arithmetic plan of ffwm_b200/csrc/guided_filter.cu, restated with torch ops and checked on the CPU against the
mirror of the reference formula (ffwm_b200.external_function.GuidedFilter, itself pinned to the reference by
tests/test_capi.py::test_guided_filter_box_sums and the train-step goldens) and against autograd of that formula for
the hand-derived backward pass.  The kernels themselves were written without GPU access; this pins what they compute."""
import pytest
import torch

from ffwm_b200.external_function import GuidedFilter


def window_sum(t, r, dim):
    """truncated window sum taken directly (gf_row_kernel / gf_col_kernel)"""
    n = t.size(dim)
    out = torch.zeros_like(t)
    for i in range(n):
        lo, hi = max(0, i - r), min(n - 1, i + r)
        out.select(dim, i).copy_(t.narrow(dim, lo, hi - lo + 1).sum(dim))
    return out


def box(t, r):
    return window_sum(window_sum(t, r, 3), r, 2)          # rows (along W) first, then columns, as the kernels do


def count(h, w, r, dtype):
    ci = torch.tensor([min(i + r, h - 1) - max(i - r, 0) + 1 for i in range(h)], dtype=dtype)
    cj = torch.tensor([min(j + r, w - 1) - max(j - r, 0) + 1 for j in range(w)], dtype=dtype)
    return (ci[:, None] * cj[None, :]).view(1, 1, h, w)   # gf_cnt(i) * gf_cnt(j)


def kernel_plan(x, y, gq, r, eps):
    """the four forward and four backward kernels of guided_filter.cu, in their order and with their formulas"""
    n = count(x.size(2), x.size(3), r, x.dtype)
    sx, sy, sxy, sxx = (box(t, r) for t in (x, y, x * y, x * x))            # GfLoadXY + rows + columns
    mx, my = sx / n, sy / n                                                  # GfStoreCoef
    cov, var = sxy / n - mx * my, sxx / n - mx * mx
    a = cov / (var + eps)
    b = my - a * mx
    ve = var + eps
    m_a = box(a, r) / n                                                      # GfLoad2 + GfStoreOut
    q = m_a * x + box(b, r) / n
    p_a, p_b = box(gq * x / n, r), box(gq / n, r)                            # GfLoadG
    g_a = p_a - p_b * mx                                                     # GfStoreGCoef
    gcov, gvar = g_a / ve, -g_a * a / ve
    gmx = -p_b * a - gcov * my - 2.0 * mx * gvar
    gx = gq * m_a + y * box(gcov / n, r) + 2.0 * x * box(gvar / n, r) + box(gmx / n, r)   # GfLoad3 + GfStoreGx
    return q, gx


@pytest.mark.parametrize("shape,r", [((2, 3, 20, 22), 4), ((1, 3, 36, 34), 16), ((1, 1, 12, 40), 5)])
def test_kernel_plan_matches_reference_formula_and_autograd(shape, r):
    g = torch.Generator().manual_seed(r)
    x = torch.rand(*shape, generator=g, dtype=torch.float64, requires_grad=True)
    y = torch.rand(*shape, generator=g, dtype=torch.float64)
    gq = torch.randn(*shape, generator=g, dtype=torch.float64)
    q_ref = GuidedFilter(r)(x, y)
    q_ref.backward(gq)
    q, gx = kernel_plan(x.detach(), y, gq, r, 1e-8)
    assert (q - q_ref.detach()).abs().max() <= 1e-12
    assert (gx - x.grad).abs().max() <= 1e-11 * max(1.0, float(x.grad.abs().max()))


def test_direct_window_sums_agree_with_cumsum_differences_in_fp32():
    """What the GPU parity tolerance rests on: in float32, at the train step's largest size (128x128, r = 32), the
    directly summed windows and the reference's cumsum differences give the same filter output to ~1e-5."""
    g = torch.Generator().manual_seed(0)
    x, y = torch.rand(1, 3, 128, 128, generator=g), torch.rand(1, 3, 128, 128, generator=g)
    want = GuidedFilter(32)(x.double(), y.double())
    q32, _ = kernel_plan(x, y, torch.zeros_like(x), 32, 1e-8)
    ref32 = GuidedFilter(32)(x, y)
    err_direct = float((q32.double() - want).abs().max())
    err_cumsum = float((ref32.double() - want).abs().max())
    assert err_direct <= 2e-5 and err_direct <= 2 * err_cumsum + 1e-6
