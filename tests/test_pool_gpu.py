"""csrc/pool.cu against ATen's max_pool2d on a B200: bit-exact forward, and a backward pass that routes every window's gradient
to the same element as ATen (first maximum; ties are the rule after a ReLU; NaN wins), ceil and floor mode, odd sizes."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(2, 3, 8, 8), (8, 48, 128, 128), (3, 5, 7, 9), (1, 4, 2, 2), (2, 2, 33, 18), (6, 512, 16, 16)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("ceil_mode", [False, True])
def test_matches_aten_bit_for_bit(shape, ceil_mode):
    from ffwm_b200.pool import max_pool2x2
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g)
    x = torch.where(torch.rand(shape, generator=g) < 0.4, torch.zeros(()), x).relu_()     # many ties (zeros), as after a ReLU
    if shape[2] >= 8:
        x[0, 0, 3, 2] = float("nan")
    xa, xb = x.cuda().requires_grad_(), x.cuda().requires_grad_()
    ya = F.max_pool2d(xa, 2, 2, ceil_mode=ceil_mode)
    yb = max_pool2x2(xb, ceil_mode=ceil_mode)
    assert ya.shape == yb.shape and torch.equal(torch.nan_to_num(ya, nan=-7.0), torch.nan_to_num(yb, nan=-7.0))
    go = torch.randn(ya.shape, generator=g).cuda()
    ya.backward(go)
    yb.backward(go)
    assert torch.equal(xa.grad, xb.grad)


def test_modules_use_the_kernel_and_decline_other_pools():
    from ffwm_b200 import _lib
    from ffwm_b200.pool import MaxPool2d
    x = torch.randn(2, 4, 16, 16, device="cuda")
    n0 = _lib.kernel_launches()
    y = MaxPool2d(2, 2, ceil_mode=True)(x)
    assert _lib.kernel_launches() - n0 == 1 and torch.equal(y, F.max_pool2d(x, 2, 2))
    n0 = _lib.kernel_launches()
    z = MaxPool2d(3, 2, padding=1)(x)
    assert _lib.kernel_launches() == n0 and torch.equal(z, F.max_pool2d(x, 3, 2, 1))
    assert torch.equal(MaxPool2d(2, 2)(x.cpu()), F.max_pool2d(x.cpu(), 2, 2))
