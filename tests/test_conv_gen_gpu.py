"""General tcgen05 convolution (ffwm_b200/csrc/conv_gen_tc.cu) against PyTorch float64 on the same inputs: every
kernel size / stride / padding / map size the networks of the path use (in both operand maths) outside the 3x3 stride-1 tiles of
conv3x3_tc.cu, forward convolution and transposed convolution, plus ragged shapes.
Tolerance relative to max|ref|: 3e-5 (K = Cin*kh*kw < 2048), 6e-5 (K < 8192), 1e-4 above (the truncating fp32
accumulation of the tensor core grows with the chain length; cuDNN's strict fp32 kernels measure 1e-5..5e-5 on these
shapes) — all inside the path's 1e-4 contract."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def tol(k):
    return 3e-5 if k < 2048 else 6e-5 if k < 8192 else 1e-4


# b, cin, cout, h, w, k, stride, pad
CONV_CASES = [
    (2, 3, 64, 128, 128, 7, 1, 3),      # generator stem
    (2, 64, 128, 128, 128, 4, 2, 1),    # generator encoders (4x4 stride 2)
    (2, 128, 256, 64, 64, 4, 2, 1),
    (2, 64, 64, 128, 128, 1, 1, 0),     # 1x1 residual input
    (8, 384, 384, 16, 16, 3, 1, 1),     # 3x3 on small maps
    (6, 512, 512, 8, 8, 3, 1, 1),
    (6, 1024, 1024, 2, 2, 3, 1, 1),     # FlowNet's deepest layer (K = 9216, split K)
    (6, 512, 1024, 4, 4, 3, 2, 1),      # stride-2 3x3
    (6, 3, 64, 128, 128, 3, 2, 1),
    (2, 1026, 512, 4, 4, 3, 1, 1),      # FlowNet inter convs (odd channel counts)
    (2, 770, 2, 8, 8, 3, 1, 1),         # flow head (2 output channels)
    (2, 1, 96, 128, 128, 5, 1, 2),      # LightCNN stem
    (2, 192, 384, 16, 16, 1, 1, 0),
    (2, 128, 128, 16, 16, 3, 1, 0),     # discriminator's unpadded 3x3
    (2, 256, 1, 16, 16, 1, 1, 0),       # discriminator head
    (1, 5, 7, 9, 11, 3, 1, 1), (3, 17, 33, 13, 7, 5, 2, 2), (1, 16, 16, 1, 1, 1, 1, 0), (2, 40, 300, 6, 5, 3, 2, 0),
    (1, 20, 24, 31, 33, 7, 2, 3), (2, 8, 8, 10, 10, 2, 2, 0), (1, 33, 5, 12, 12, 4, 1, 2),
]
# b, cin, cout, h, w, k, stride, pad, output_padding
TCONV_CASES = [
    (6, 1024, 512, 2, 2, 4, 2, 1, 0),   # FlowNet deconv5
    (6, 1026, 256, 4, 4, 4, 2, 1, 0),
    (6, 194, 32, 32, 32, 4, 2, 1, 0),
    (6, 2, 2, 64, 64, 4, 2, 1, 0),      # flow upsampler
    (2, 12, 20, 7, 9, 3, 2, 1, 0), (2, 12, 20, 7, 9, 3, 2, 1, 1), (1, 9, 6, 5, 5, 4, 2, 1, 0), (2, 16, 16, 6, 6, 3, 1, 1, 0),
    (1, 6, 40, 5, 4, 5, 2, 2, 1), (1, 4, 4, 3, 3, 2, 2, 0, 0), (1, 7, 3, 4, 6, 7, 2, 3, 1),
]


class _Ops:
    """ffwm_b200.ops with one operand math bound to the general convolution calls (the weight gradient is 3xBF16 only)."""

    def __init__(self, math):
        from ffwm_b200 import ops as o
        self.o, self.math = o, math
        self.conv_wgrad = o.conv_wgrad

    def conv_pack_weights(self, *a):
        return self.o.conv_pack_weights(*a, math=self.math)

    def conv_forward(self, *a):
        return self.o.conv_forward(*a, math=self.math)


@pytest.fixture(scope="module", params=[0, 1], ids=["tf32x3", "bf16x3"])
def ops(request):
    return _Ops(request.param)


@pytest.mark.parametrize("b,cin,cout,h,w,k,s,p", CONV_CASES)
def test_conv_forward_matches_fp64(ops, b, cin, cout, h, w, k, s, p):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g)
    want = F.conv2d(x.double(), wt.double(), bias.double(), stride=s, padding=p)
    xd, wd, bd = x.to(DEV), wt.to(DEV), bias.to(DEV)
    out = torch.full(want.shape, float("nan"), device=DEV)
    ops.conv_forward(xd, ops.conv_pack_weights(wd, False, s, p, False), bd, out, k, k, s, p, False)
    torch.cuda.synchronize()
    assert rel(out.cpu(), want) <= tol(cin * k * k), (rel(out.cpu(), want), tol(cin * k * k))
    # the data gradient of this convolution = the transposed mode on the same weight (in_major)
    go = torch.randn(want.shape, generator=g)
    xr = x.double().requires_grad_(True)
    F.conv2d(xr, wt.double(), None, stride=s, padding=p).backward(go.double())
    gx = torch.full(x.shape, float("nan"), device=DEV)
    ops.conv_forward(go.to(DEV), ops.conv_pack_weights(wd, True, s, p, True), None, gx, k, k, s, p, True)
    assert rel(gx.cpu(), xr.grad) <= tol(cout * k * k), (rel(gx.cpu(), xr.grad), tol(cout * k * k))


@pytest.mark.parametrize("b,cin,cout,h,w,k,s,p,op", TCONV_CASES)
def test_conv_transpose_matches_fp64(ops, b, cin, cout, h, w, k, s, p, op):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k + 7)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.randn(cin, cout, k, k, generator=g) / (cin * k * k / s / s) ** 0.5
    bias = torch.randn(cout, generator=g)
    xr = x.double().requires_grad_(True)
    want = F.conv_transpose2d(xr, wt.double(), bias.double(), stride=s, padding=p, output_padding=op)
    xd, wd, bd = x.to(DEV), wt.to(DEV), bias.to(DEV)
    out = torch.full(want.shape, float("nan"), device=DEV)
    ops.conv_forward(xd, ops.conv_pack_weights(wd, True, s, p, True), bd, out, k, k, s, p, True)
    torch.cuda.synchronize()
    assert rel(out.cpu(), want.detach()) <= tol(cin * k * k), (rel(out.cpu(), want.detach()), tol(cin * k * k))
    go = torch.randn(want.shape, generator=g)
    want.backward(go.double())
    gx = torch.full(x.shape, float("nan"), device=DEV)
    ops.conv_forward(go.to(DEV), ops.conv_pack_weights(wd, False, s, p, False), None, gx, k, k, s, p, False)
    assert rel(gx.cpu(), xr.grad) <= tol(cout * k * k), (rel(gx.cpu(), xr.grad), tol(cout * k * k))


def test_strided_views_and_rejections(ops):
    g = torch.Generator().manual_seed(3)
    big = torch.randn(2, 40, 20, 24, generator=g).to(DEV)
    x = big[:, 4:28, 2:18, 3:21]                                     # non-contiguous view
    wt = (torch.randn(10, 24, 3, 3, generator=g) / 15).to(DEV)
    out_big = torch.zeros(2, 16, 16, 18, device=DEV)
    out = out_big[:, 3:13]                                           # channel-sliced output
    ops.conv_forward(x, ops.conv_pack_weights(wt, False, 1, 1, False), None, out, 3, 3, 1, 1, False)
    want = F.conv2d(x.double(), wt.double(), None, padding=1)
    assert rel(out, want) <= 3e-5 and float(out_big[:, :3].abs().max()) == 0 and float(out_big[:, 13:].abs().max()) == 0
    with pytest.raises(RuntimeError):                                # output size does not match the geometry
        ops.conv_forward(x, ops.conv_pack_weights(wt, False, 1, 1, False), None, torch.zeros(2, 10, 15, 18, device=DEV), 3, 3, 1, 1, False)
    with pytest.raises(RuntimeError):                                # stride 3 is not supported
        ops.conv_pack_weights(wt, False, 3, 1, False)


@pytest.mark.parametrize("case", ["conv_s2", "conv_1x1", "convT", "conv_7x7"])
def test_modules_autograd_match_fp64(case):
    """ffwm_b200.conv.Conv2d / ConvTranspose2d route these shapes through the general kernel: forward and grad_input
    on tcgen05, weight / bias gradients on cuDNN; all three against float64 autograd."""
    from ffwm_b200 import _lib
    from ffwm_b200.conv import Conv2d, ConvTranspose2d
    torch.backends.cudnn.allow_tf32 = False          # the library weight gradient in strict fp32 (TF32 measures 3e-4 here)
    torch.manual_seed(1)
    if case == "conv_s2":
        m, x = Conv2d(24, 40, 4, 2, 1), torch.randn(2, 24, 20, 28)
    elif case == "conv_1x1":
        m, x = Conv2d(24, 40, 1, 1, 0, bias=False), torch.randn(2, 24, 9, 16)
    elif case == "conv_7x7":
        m, x = Conv2d(3, 16, 7, 1, 3), torch.randn(2, 3, 30, 30)
    else:
        m, x = ConvTranspose2d(24, 12, 4, 2, 1), torch.randn(2, 24, 9, 7)
    m = m.to(DEV)
    x = x.to(DEV).requires_grad_(True)
    n0 = _lib.kernel_launches()
    out = m(x)
    go = torch.randn_like(out)
    out.backward(go)
    assert _lib.kernel_launches() - n0 >= 4                          # pack + conv, forward and data gradient
    ref = torch.nn.ConvTranspose2d(24, 12, 4, 2, 1) if case == "convT" else torch.nn.Conv2d(
        m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding, bias=m.bias is not None)
    ref = ref.double()
    ref.load_state_dict({k: v.detach().double().cpu() for k, v in m.state_dict().items()})
    xr = x.detach().double().cpu().requires_grad_(True)
    outr = ref(xr)
    outr.backward(go.double().cpu())
    assert rel(out.cpu(), outr.detach()) <= 3e-5 and rel(x.grad.cpu(), xr.grad) <= 3e-5
    assert rel(m.weight.grad.cpu(), ref.weight.grad) <= 1e-4
    if m.bias is not None:
        assert rel(m.bias.grad.cpu(), ref.bias.grad) <= 1e-4


# ---------------------------------------------------------------- weight gradient (csrc/conv_gen_wgrad_tc.cu)
WGRAD_TOL = 4e-5      # of max|ref|; chains are capped at 768 accumulator updates per CTA (1.5e-5 measured on the step's shapes)


# stride-1 shapes whose rows fill whole 64-pixel stages take the weight gradient's ROW mode (one tile per kernel row shared by
# its taps): widths 128 / 64 (one or two segments per row), 32 / 16 (two / four rows per stage), 5x5, 2x2, no padding
# (and at least 96 input channels: narrower tiles stay in the gathered mode)
ROW_MODE_CASES = [(8, 195, 195, 128, 128, 3, 1, 1), (2, 64, 64, 64, 64, 3, 1, 1), (2, 96, 200, 32, 32, 3, 1, 1), (2, 100, 50, 64, 64, 5, 1, 2),
                  (2, 120, 40, 18, 18, 3, 1, 0), (1, 130, 24, 17, 33, 2, 1, 0), (3, 104, 136, 16, 16, 3, 1, 1), (1, 260, 20, 16, 64, 3, 1, 1)]


@pytest.mark.parametrize("b,cin,cout,h,w,k,s,p", CONV_CASES + ROW_MODE_CASES)
def test_conv_wgrad_matches_fp64(ops, b, cin, cout, h, w, k, s, p):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k + 1)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.zeros(cout, cin, k, k, dtype=torch.float64, requires_grad=True)
    out = F.conv2d(x.double(), wt, None, stride=s, padding=p)
    go = torch.randn(out.shape, generator=g)
    out.backward(go.double())
    gw = torch.full((cout, cin, k, k), float("nan"), device=DEV)
    ops.conv_wgrad(go.to(DEV), x.to(DEV), gw, s, p)
    torch.cuda.synchronize()
    assert rel(gw.cpu(), wt.grad) <= WGRAD_TOL, rel(gw.cpu(), wt.grad)


@pytest.mark.parametrize("b,cin,cout,h,w,k,s,p,op", TCONV_CASES)
def test_conv_transpose_wgrad_matches_fp64(ops, b, cin, cout, h, w, k, s, p, op):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k + 2)
    x = torch.randn(b, cin, h, w, generator=g)
    wt = torch.zeros(cin, cout, k, k, dtype=torch.float64, requires_grad=True)
    out = F.conv_transpose2d(x.double(), wt, None, stride=s, padding=p, output_padding=op)
    go = torch.randn(out.shape, generator=g)
    out.backward(go.double())
    gw = torch.full((cin, cout, k, k), float("nan"), device=DEV)
    ops.conv_wgrad(x.to(DEV), go.to(DEV), gw, s, p)
    torch.cuda.synchronize()
    assert rel(gw.cpu(), wt.grad) <= WGRAD_TOL, rel(gw.cpu(), wt.grad)


def test_conv_wgrad_is_deterministic_and_takes_strided_targets(ops):
    g = torch.Generator().manual_seed(11)
    x, go = torch.randn(4, 40, 24, 24, generator=g).to(DEV), torch.randn(4, 72, 12, 12, generator=g).to(DEV)
    a, b2 = torch.empty(72, 40, 4, 4, device=DEV), torch.empty(72, 40, 4, 4, device=DEV)
    ops.conv_wgrad(go, x, a, 2, 1)
    ops.conv_wgrad(go, x, b2, 2, 1)
    assert torch.equal(a, b2)
    big = torch.zeros(80, 44, 4, 4, device=DEV)
    ops.conv_wgrad(go, x, big[4:76, 2:42], 2, 1)                      # a view into a larger tensor
    assert torch.equal(big[4:76, 2:42], a) and float(big[:4].abs().max()) == 0 and float(big[:, 42:].abs().max()) == 0
