"""csrc/conv_few.cu (direct fp32 convolutions for <= 4 input channels, or <= 4 output channels at stride 1) against float64
torch: forward mode, and the (in_major, flip) reading of the weight that makes the same kernels a data gradient.
fp32 FFMA accumulation in one chain of up to 3136 products: 4e-6 of max|out|."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# b, cin, cout, h, w, k, stride, pad
FEW_CASES = [
    (4, 1, 96, 128, 128, 5, 1, 2),      # LightCNN stem
    (2, 3, 64, 128, 128, 7, 1, 3),      # generator stem
    (3, 3, 64, 64, 64, 3, 2, 1),        # discriminator's first convolution
    (2, 3, 64, 128, 128, 3, 1, 1),      # FlowNet conv0 / VGG conv1_1
    (1, 4, 40, 33, 17, 3, 1, 0), (2, 2, 5, 9, 70, 4, 2, 1), (1, 1, 1, 6, 6, 2, 1, 0),
    (2, 96, 1, 128, 128, 5, 1, 2),      # (the stem's data gradient as a forward problem)
    (2, 195, 3, 128, 128, 3, 1, 1),     # reconstruction head
    (2, 64, 2, 32, 32, 3, 1, 1),        # flow head
    (1, 50, 4, 37, 45, 3, 1, 1), (1, 9, 3, 40, 130, 7, 1, 3), (2, 5, 1, 12, 260, 1, 1, 0),
]


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("b,cin,cout,h,w,k,s,p", FEW_CASES)
def test_forward_matches_fp64(b, cin, cout, h, w, k, s, p):
    from ffwm_b200 import ops
    g = torch.Generator().manual_seed(cin * 100 + cout + k)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(x.double(), wt.double(), bias.double(), stride=s, padding=p)
    out = torch.full(want.shape, float("nan"), device=DEV)
    ops.conv_few(x.to(DEV), wt.to(DEV), False, False, bias.to(DEV), out, s, p)
    assert rel(out.cpu(), want) <= 4e-6, rel(out.cpu(), want)


@pytest.mark.parametrize("b,cin,cout,h,w,k,p", [(2, 1, 96, 128, 128, 5, 2), (2, 3, 64, 64, 64, 7, 3), (2, 195, 3, 64, 64, 3, 1), (1, 40, 2, 33, 47, 3, 0)])
def test_data_gradient_through_in_major_and_flip(b, cin, cout, h, w, k, p):
    """grad_input of conv2d(x, W, stride 1, pad p) = conv_few(grad_out, W, in_major=1, flip=1, pad k-1-p)."""
    from ffwm_b200 import ops
    g = torch.Generator().manual_seed(7 + cin + cout)
    x = torch.randn(b, cin, h, w, generator=g, dtype=torch.float64, requires_grad=True)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    out = F.conv2d(x, wt.double(), None, padding=p)
    go = torch.randn(out.shape, generator=g)
    out.backward(go.double())
    gx = torch.full((b, cin, h, w), float("nan"), device=DEV)
    ops.conv_few(go.to(DEV), wt.to(DEV), True, True, None, gx, 1, k - 1 - p)
    assert rel(gx.cpu(), x.grad) <= 4e-6, rel(gx.cpu(), x.grad)


def test_strided_views_and_rejections():
    from ffwm_b200 import ops
    g = torch.Generator().manual_seed(3)
    big = torch.randn(2, 6, 40, 44, generator=g).to(DEV)
    x = big[:, 1:4, 2:34, 4:40]                                          # a view: 3 channels, 32 x 36
    wt = torch.randn(8, 3, 3, 3, generator=g).to(DEV)
    outbig = torch.zeros(2, 10, 34, 40, device=DEV)
    out = outbig[:, 1:9, 1:33, 2:38]
    ops.conv_few(x, wt, False, False, None, out, 1, 1)
    want = F.conv2d(x.double(), wt.double(), None, padding=1)
    assert rel(out, want) <= 2e-6 and float(outbig[:, 0].abs().max()) == 0 and float(outbig[:, :, 0].abs().max()) == 0
    with pytest.raises(RuntimeError):                                    # 8 -> 8 channels: not a degenerate shape
        ops.conv_few(torch.zeros(1, 8, 8, 8, device=DEV), torch.zeros(8, 8, 3, 3, device=DEV), False, False, None, torch.zeros(1, 8, 8, 8, device=DEV), 1, 1)
    with pytest.raises(RuntimeError):                                    # few outputs need stride 1
        ops.conv_few(torch.zeros(1, 8, 8, 8, device=DEV), torch.zeros(2, 8, 3, 3, device=DEV), False, False, None, torch.zeros(1, 2, 4, 4, device=DEV), 2, 1)
    with pytest.raises(NotImplementedError):
        ops.conv_few(torch.zeros(1, 3, 8, 8), torch.zeros(8, 3, 3, 3), False, False, None, torch.zeros(1, 8, 8, 8), 1, 1)


def test_modules_route_degenerate_layers_to_the_direct_kernels():
    from ffwm_b200 import _lib, conv as C
    torch.manual_seed(5)
    m = C.Conv2d(1, 96, 5, 1, 2).to(DEV)
    x = torch.randn(2, 1, 64, 64, device=DEV, requires_grad=True)
    n0 = _lib.kernel_launches()
    y = m(x)
    assert _lib.kernel_launches() - n0 == 1                              # no weight packing, one direct kernel
    ref = F.conv2d(x.double(), m.weight.double(), m.bias.double(), padding=2)
    assert rel(y, ref) <= 2e-6
    go = torch.randn_like(y)
    y.backward(go)
    gx_ref = torch.autograd.grad(ref, x, go.double())[0] if False else None
    xd = x.detach().double().requires_grad_()
    F.conv2d(xd, m.weight.double(), m.bias.double(), padding=2).backward(go.double())
    assert rel(x.grad, xd.grad) <= 2e-6
