"""csrc/spectral_norm.cu (all spectrally normalised layers of a network in three launches forward, two backward) against
torch's own per-layer spectral_norm hooks on a B200: normalised weights, u / v updates over several steps, gradients of
weight_orig, eval mode.  fp32 against fp32 with different summation orders: 2e-6 on weights and buffers, 2e-5 on gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _sn_modules(net):
    return [m for m in net.modules() if hasattr(m, "weight_orig")]


def _run_hooks(net):
    for m in net.modules():
        for hook in m._forward_pre_hooks.values():
            hook(m, None)


@pytest.mark.parametrize("which", ["generator", "discriminator"])
def test_fused_spectral_norm_matches_torch_hooks(monkeypatch, which):
    from ffwm_b200 import base_networks as BN, spectral
    monkeypatch.setattr(BN, "BATCHED_SN", False)                 # build with torch's per-layer hooks, batch one copy by hand
    torch.manual_seed(2)
    with torch.device("cuda"):
        ref = BN.FFWM(sn=True) if which == "generator" else BN.MSDiscriminator(128, sigmoid=False)
    net = copy.deepcopy(ref)
    monkeypatch.setattr(spectral, "FUSED_SN", True)
    mgr = spectral.batch_spectral_norm(net)
    assert mgr._fused_ok()
    ref.train(), net.train()
    from ffwm_b200 import _lib
    for step in range(3):                                        # the power iteration advances on every training forward
        _run_hooks(ref)
        n0 = _lib.kernel_launches()
        mgr._update(net, None)
        assert _lib.kernel_launches() - n0 == 3
        ma, mb = _sn_modules(ref), _sn_modules(net)
        assert len(ma) == len(mb) == mgr.n_layers
        torch.manual_seed(10 + step)
        loss_a = loss_b = 0.0
        for a, b in zip(ma, mb):
            scale = float(a.weight.detach().abs().max())
            assert float((a.weight - b.weight).abs().max()) <= 2e-6 * scale
            assert float((a.weight_u - b.weight_u).abs().max()) <= 2e-6 and float((a.weight_v - b.weight_v).abs().max()) <= 2e-6
            r = torch.randn_like(a.weight)
            loss_a = loss_a + (a.weight * r).sum() + a.weight.square().sum()
            loss_b = loss_b + (b.weight * r).sum() + b.weight.square().sum()
        for p in list(ref.parameters()) + list(net.parameters()):
            p.grad = None
        loss_a.backward()
        n0 = _lib.kernel_launches()
        loss_b.backward()
        assert _lib.kernel_launches() - n0 == 2
        for a, b in zip(ma, mb):
            ga, gb = a.weight_orig.grad, b.weight_orig.grad
            assert float((ga - gb).abs().max()) <= 2e-5 * float(ga.abs().max())
    ref.eval(), net.eval()
    u0 = _sn_modules(net)[0].weight_u.clone()
    _run_hooks(ref)
    mgr._update(net, None)
    for a, b in zip(_sn_modules(ref), _sn_modules(net)):
        assert float((a.weight - b.weight).abs().max()) <= 2e-6 * float(a.weight.abs().max())
    assert torch.equal(u0, _sn_modules(net)[0].weight_u)          # no power iteration in eval mode


def test_fused_path_declines_what_it_does_not_cover(monkeypatch):
    from ffwm_b200 import spectral
    from torch.nn.utils import spectral_norm
    net = torch.nn.Sequential(spectral_norm(torch.nn.Conv2d(4, 8, 3), n_power_iterations=2)).cuda()
    mgr = spectral.batch_spectral_norm(net)
    assert not mgr._fused_ok()                                    # two power iterations: the batched torch formulation
    ref = copy.deepcopy(net[0].weight_orig)
    y = net(torch.randn(1, 4, 8, 8, device="cuda"))
    assert y.shape == (1, 8, 6, 6) and torch.equal(ref, net[0].weight_orig)
    cpu = torch.nn.Sequential(spectral_norm(torch.nn.Conv2d(4, 8, 3)))
    assert not spectral.batch_spectral_norm(cpu)._fused_ok()      # CPU tensors
