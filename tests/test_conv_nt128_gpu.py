"""tcgen05 3x3 convolution with 128 output channels per CTA (conv3x3_tc_kernel<128, 128>) against float64.

Validated on a B200 in round 2 (gpurun call 1, profiles/r02a_*)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("b,cin,cout,h", [(1, 8, 128, 2), (2, 16, 65, 8), (2, 195, 195, 10), (1, 128, 128, 128), (1, 20, 130, 5),
                                          (1, 3, 256, 3), (1, 24, 300, 1)])
def test_conv3x3_forward_nt128_matches_fp64(b, cin, cout, h):
    from ffwm_b200 import ops
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    x = torch.randn(b, cin, h, 128, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    xd, wd, bd = x.to(DEV), w.to(DEV), bias.to(DEV)
    out = torch.full((b, cout, h, 128), float("nan"), device=DEV)
    ops.conv3x3_forward(xd, ops.conv3x3_pack_weights(wd, nt=128), bd, out, nt=128)      # default math: 3xBF16
    torch.cuda.synchronize()
    assert rel(out.cpu(), want) <= (3e-5 if cin * 9 < 2048 else 5e-5)      # default operand math (3xBF16)
    out64 = torch.empty_like(out)
    ops.conv3x3_forward(xd, ops.conv3x3_pack_weights(wd), bd, out64)
    assert rel(out, out64) <= 1e-6      # same MMAs in the same order per output element: expected to agree bit for bit


def test_conv3x3_dgrad_nt128():
    from ffwm_b200 import ops
    g = torch.Generator().manual_seed(5)
    b, cin, cout, h = 2, 150, 70, 9
    x = torch.randn(b, cin, h, 128, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64) / (cin * 9) ** 0.5
    go = torch.randn(b, cout, h, 128, generator=g, dtype=torch.float64)
    F.conv2d(x, w, None, padding=1).backward(go)
    gx = torch.empty(b, cin, h, 128, device=DEV)
    ops.conv3x3_forward(go.float().to(DEV), ops.conv3x3_pack_weights(w.float().to(DEV), dgrad=True, nt=128), None, gx, nt=128)
    assert rel(gx.cpu(), x.grad) <= 3e-5


def test_nt128_rejected_for_other_widths():
    from ffwm_b200 import ops
    x = torch.zeros(1, 8, 4, 64, device=DEV)
    with pytest.raises(RuntimeError):
        ops.conv3x3_forward(x, ops.conv3x3_pack_weights(torch.zeros(128, 8, 3, 3, device=DEV), nt=128), None,
                            torch.zeros(1, 128, 4, 64, device=DEV), nt=128)
