"""Batched spectral norm (ffwm_b200/spectral.py) against torch's per-layer hooks on the CPU: same outputs, gradients,
u/v buffer updates, eval-mode behaviour and state_dict keys; only the summation order (bmm vs mv) differs."""
import copy

import torch

from ffwm_b200 import base_networks as BN
from ffwm_b200.spectral import batch_spectral_norm


def _flat(ts):
    return torch.cat([t.reshape(-1) for t in ts])


def test_discriminator_batched_sn_matches_per_layer_hooks(monkeypatch):
    monkeypatch.setattr(BN, "BATCHED_SN", False)         # the reference copy keeps torch's per-layer hooks
    torch.manual_seed(0)
    ref = BN.MSDiscriminator(128, sigmoid=False).double()
    net = copy.deepcopy(ref)
    mgr = batch_spectral_norm(net)
    assert mgr.n_layers == 9 and len(mgr.groups) == 3
    x = torch.rand(2, 3, 128, 128, dtype=torch.float64)
    for _ in range(3):                                   # the power iteration advances on every training forward
        ya, yb = ref(x), net(x)
        assert (ya - yb).abs().max() <= 1e-12
        ref.zero_grad()
        net.zero_grad()
        ya.square().sum().backward()
        yb.square().sum().backward()
        ga, gb = _flat(p.grad for p in ref.parameters()), _flat(p.grad for p in net.parameters())
        assert (ga - gb).abs().max() <= 1e-11 * float(ga.abs().max())
    for (na, ba), (nb, bb) in zip(ref.named_buffers(), net.named_buffers()):
        assert na == nb and (ba.double() - bb.double()).abs().max() <= 1e-12
    assert list(ref.state_dict()) == list(net.state_dict())
    ref.eval()
    net.eval()
    u0 = net.nets[0][0].weight_u.clone()
    assert (ref(x) - net(x)).abs().max() <= 1e-12
    assert torch.equal(u0, net.nets[0][0].weight_u)      # no power iteration in eval mode


def test_generator_batched_sn_matches_per_layer_hooks_fp32(monkeypatch):
    monkeypatch.setattr(BN, "BATCHED_SN", False)
    torch.manual_seed(1)
    ref = BN.FFWM(sn=True)
    net = copy.deepcopy(ref)
    mgr = batch_spectral_norm(net)
    assert mgr.n_layers == 52 and len(mgr.groups) == 19
    # spectral norm itself, layer by layer: run both sets of hooks without the (CUDA-only) warps of the forward pass
    ref.train()
    net.train()
    for m in ref.modules():
        for hook in m._forward_pre_hooks.values():
            hook(m, None)
    mgr._update(net, None)
    for ma, mb in zip([m for m in ref.modules() if hasattr(m, "weight_orig")],
                      [m for m in net.modules() if hasattr(m, "weight_orig")]):
        scale = float(ma.weight.detach().abs().max())
        assert (ma.weight - mb.weight).abs().max() <= 2e-6 * scale
        assert (ma.weight_u - mb.weight_u).abs().max() <= 2e-6 and (ma.weight_v - mb.weight_v).abs().max() <= 2e-6
    # gradient through sigma
    la = sum((m.weight ** 2).sum() for m in ref.modules() if hasattr(m, "weight_orig"))
    lb = sum((m.weight ** 2).sum() for m in net.modules() if hasattr(m, "weight_orig"))
    la.backward()
    lb.backward()
    for pa, pb in zip(ref.parameters(), net.parameters()):
        if pa.grad is not None:
            assert (pa.grad - pb.grad).abs().max() <= 1e-5 * max(1.0, float(pa.grad.abs().max()))


def test_fused_plan_table_layout():
    """Host side of the fused path (ffwm_b200/spectral.py:_Plan): the device table csrc/spectral_norm.cu walks — 8 int64 per
    layer, then three block-prefix arrays with ceil(w/32), ceil(h/8), ceil(h*w/4096) blocks per layer."""
    import math
    from torch.nn.utils import spectral_norm
    from ffwm_b200.spectral import _Plan
    torch.manual_seed(0)
    mods = [spectral_norm(torch.nn.Conv2d(3, 8, 3)), spectral_norm(torch.nn.Conv2d(8, 40, 5)), spectral_norm(torch.nn.Conv2d(40, 4, 1))]
    plan = _Plan(mods, 1e-12)
    t = plan.table.tolist()
    n = len(mods)
    oe = ot = os_ = 0
    for i, m in enumerate(mods):
        h, w = m.weight_orig.shape[0], m.weight_orig[0].numel()
        assert t[8 * i:8 * i + 8] == [m.weight_orig.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr(), h, w, oe, ot, os_]
        assert plan.slices[i] == (oe, h * w, tuple(m.weight_orig.shape))
        oe, ot, os_ = oe + h * w, ot + w, os_ + h
    assert (plan.total, plan.sum_w, plan.sum_h) == (oe, ot, os_)
    pre = t[8 * n:]
    for k, per in enumerate((lambda h, w: math.ceil(w / 32), lambda h, w: math.ceil(h / 8), lambda h, w: math.ceil(h * w / 4096))):
        want = [0]
        for m in mods:
            want.append(want[-1] + per(m.weight_orig.shape[0], m.weight_orig[0].numel()))
        assert pre[k * (n + 1):(k + 1) * (n + 1)] == want and plan.blocks[k] == want[-1]
    assert plan.key == _Plan.pointers(mods)
