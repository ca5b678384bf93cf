"""csrc/batch_norm.cu against torch's batch norm (float64) — training-mode statistics, running-stat update, the absorbed
LeakyReLU and residual add, both directions.  Tolerances: 2e-6 of the output scale forward, 2e-5 of max|grad| backward
(fp32 arithmetic against a float64 reference; north_star asks <= 1e-4)."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

pytestmark = pytest.mark.gpu


def _ref(x, res, w, b, rm, rv, momentum, eps, slope):
    y = F.batch_norm(x, rm, rv, w, b, True, momentum, eps)
    if res is not None:
        y = y + res
    return y if slope is None else F.leaky_relu(y, slope)


CASES = [
    # n, c, h, w, slope, residual, offset
    (4, 16, 64, 64, 0.2, False, 0.0),
    (2, 5, 7, 9, 0.2, False, 0.0),           # ragged: scalar path
    (8, 195, 32, 32, 0.2, True, 0.0),        # residual block tail
    (3, 4, 2, 2, None, False, 0.0),          # FlowNet's deepest map, no activation
    (2, 8, 128, 128, None, False, 50.0),     # |mean| >> std: conditioning of the variance (no activation: the sign of a
                                             # pre-activation within 3e-6 of zero is not defined at this conditioning)
    (1, 3, 96, 96, None, True, -3.0),        # several chunks per plane, batch 1
    (8, 64, 128, 128, 0.2, False, 0.5),
    (2, 6, 16, 16, 0.2, True, 0.0),          # one block per channel (<= 8192 values per channel): 1 chunk per thread
    (8, 12, 16, 32, None, False, 1.0),       # ... 4 chunks per thread
    (7, 3, 20, 28, 0.2, False, 0.0),         # ... a ragged number of chunks
]


@pytest.mark.parametrize("n,c,h,w,slope,use_res,offset", CASES)
def test_forward_backward_match_torch_float64(n, c, h, w, slope, use_res, offset):
    from ffwm_b200.norm import BatchNormFunction
    g = torch.Generator().manual_seed(1000 * n + c + h)
    dev = "cuda:0"
    x = (torch.randn(n, c, h, w, generator=g) * (0.5 + torch.rand(1, c, 1, 1, generator=g)) + offset + torch.randn(1, c, 1, 1, generator=g)).to(dev)
    res = torch.randn(n, c, h, w, generator=g).to(dev) if use_res else None
    wt, bs = (torch.rand(c, generator=g) + 0.5).to(dev), torch.randn(c, generator=g).to(dev)
    rm, rv = torch.randn(c, generator=g).to(dev), (torch.rand(c, generator=g) + 0.5).to(dev)
    go = torch.randn(n, c, h, w, generator=g).to(dev)
    momentum, eps = 0.1, 1e-5

    xs, ws, bs_ = x.clone().requires_grad_(), wt.clone().requires_grad_(), bs.clone().requires_grad_()
    rs = res.clone().requires_grad_() if use_res else None
    rm1, rv1 = rm.clone(), rv.clone()
    y = BatchNormFunction.apply(xs, rs, ws, bs_, rm1, rv1, momentum, eps, 1.0 if slope is None else slope)
    y.backward(go)

    xd, wd, bd = x.double().requires_grad_(), wt.double().requires_grad_(), bs.double().requires_grad_()
    rd = res.double().requires_grad_() if use_res else None
    rm2, rv2 = rm.double(), rv.double()
    yd = _ref(xd, rd, wd, bd, rm2, rv2, momentum, eps, slope)
    yd.backward(go.double())

    def close(a, b, tol, what):
        scale = float(b.detach().abs().max()) + 1e-30
        err = float((a.detach().double() - b.detach()).abs().max()) / scale
        assert err <= tol, "%s: %.3g > %.3g" % (what, err, tol)

    close(y, yd, 2e-6, "y")
    close(rm1, rm2, 1e-6, "running_mean")
    close(rv1, rv2, 2e-6, "running_var")
    # a pre-activation within rounding of zero may take the other branch than float64: exclude those from the comparison
    safe = torch.ones_like(x, dtype=torch.bool)
    if slope is not None:
        pre = F.batch_norm(x.double(), None, None, wt.double(), bs.double(), True, 0.0, eps) + (res.double() if use_res else 0)
        safe = pre.abs() > 1e-5 * pre.abs().max()
        assert float(safe.float().mean()) > 0.999
    close(xs.grad * safe, xd.grad * safe, 2e-5, "grad_x")
    close(ws.grad, wd.grad, 2e-5, "grad_gamma")
    close(bs_.grad, bd.grad, 2e-5, "grad_beta")
    if use_res:
        close(rs.grad * safe, rd.grad * safe, 2e-5, "grad_residual")


def test_small_maps_also_pass_on_the_two_kernel_path():
    from ffwm_b200 import _lib
    old = _lib.set_option("BN_NO_SMALL", 1)
    try:
        for case in [c for c in CASES if c[0] * c[2] * c[3] <= 8192]:
            test_forward_backward_match_torch_float64(*case)
    finally:
        _lib.set_option("BN_NO_SMALL", old)


def test_module_is_a_drop_in_for_bn_plus_leaky_relu():
    from ffwm_b200.norm import BatchNorm2d, fuse_activations
    torch.manual_seed(3)
    ref = nn.Sequential(nn.Conv2d(6, 10, 3, padding=1), nn.BatchNorm2d(10), nn.LeakyReLU(0.2, True)).cuda()
    mine = nn.Sequential(*fuse_activations([nn.Conv2d(6, 10, 3, padding=1), BatchNorm2d(10), nn.LeakyReLU(0.2, True)])).cuda()
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(4, 6, 24, 20, device="cuda")
    for _ in range(3):                                   # running statistics accumulate identically
        ya, yb = ref(x), mine(x)
        torch.testing.assert_close(yb, ya, rtol=1e-5, atol=1e-5)
        ya.square().sum().backward()
        yb.square().sum().backward()
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), mine.state_dict().items()):
        torch.testing.assert_close(vb, va, rtol=1e-5, atol=1e-6, msg=ka)
    for (name, pa), pb in zip(ref.named_parameters(), mine.parameters()):
        if name == "0.bias":
            # a bias in front of a batch norm has an exactly zero gradient (the mean is subtracted again): both sides hold
            # nothing but the rounding noise of their own summation order there (|g| ~ 1e-3 of the weight gradients)
            assert float(pb.grad.abs().max()) < 1e-2 * float(mine[0].weight.grad.abs().max())
            continue
        torch.testing.assert_close(pb.grad, pa.grad, rtol=2e-4, atol=2e-4 * float(pa.grad.abs().max()), msg=name)
    ref.eval()
    mine.eval()                                          # inference: torch's path + the absorbed activation
    torch.testing.assert_close(mine(x), ref(x), rtol=1e-5, atol=1e-5)


def test_runs_are_bit_identical():
    from ffwm_b200.norm import BatchNormFunction
    torch.manual_seed(5)
    x = torch.randn(8, 32, 64, 64, device="cuda")
    w, b = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda")
    go = torch.randn_like(x)
    outs = []
    for _ in range(2):
        xs = x.clone().requires_grad_()
        ws = w.clone().requires_grad_()
        y = BatchNormFunction.apply(xs, None, ws, b, torch.zeros(32, device="cuda"), torch.ones(32, device="cuda"), 0.1, 1e-5, 0.2)
        y.backward(go)
        outs.append((y.detach().clone(), xs.grad.clone(), ws.grad.clone()))
    for a, c in zip(*outs):
        assert torch.equal(a, c)


def test_rejects_single_value_per_channel_and_cpu_tensors():
    from ffwm_b200 import ops
    x = torch.zeros(1, 4, 1, 1, device="cuda")
    v = torch.zeros(4, device="cuda")
    with pytest.raises((ValueError, RuntimeError)):
        ops.batch_norm_forward(x, None, v, v, v.clone(), v.clone(), 0.1, 1e-5, 1.0, torch.empty_like(x), v.clone(), v.clone())
    with pytest.raises(NotImplementedError):
        ops.batch_norm_forward(x.cpu(), None, v, v, None, None, 0.1, 1e-5, 1.0, torch.empty_like(x), v.clone(), v.clone())


@pytest.mark.parametrize("shape", [(8, 195, 32, 32), (2, 5, 7, 9), (1, 3, 1, 1), (4, 16, 128, 128), (3, 7, 70, 66)])
def test_channel_sum_matches_float64(shape):
    from ffwm_b200 import ops
    torch.manual_seed(sum(shape))
    x = torch.randn(shape, device="cuda") + 0.25
    got = ops.channel_sum(x)
    want = x.double().sum((0, 2, 3))
    scale = float(x.double().abs().sum((0, 2, 3)).max())
    assert float((got.double() - want).abs().max()) <= 2e-7 * scale
    assert torch.equal(got, ops.channel_sum(x))              # deterministic


def test_deferred_counters_on_the_kernel_path():
    from ffwm_b200.norm import BatchNorm2d, DeferredCounters
    bn = BatchNorm2d(6).cuda()
    ctr = DeferredCounters([bn])
    x = torch.randn(4, 6, 8, 8, device="cuda")
    for _ in range(3):
        bn(x)
    assert int(bn.num_batches_tracked) == 0                  # counted on the host, not yet added
    ctr.flush()
    assert int(bn.num_batches_tracked) == 3
    ctr.flush()
    assert int(bn.num_batches_tracked) == 3
