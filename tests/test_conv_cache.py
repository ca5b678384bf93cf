"""Host logic of the packed-weight cache (ffwm_b200/conv.py:_cached) — no GPU needed: `pack` is a counter here."""
import torch

from ffwm_b200 import conv as C


def _counter():
    calls = []

    def pack():
        calls.append(1)
        return len(calls)
    return calls, pack


def test_frozen_weight_is_packed_once_and_again_after_an_in_place_update():
    w = torch.nn.Parameter(torch.randn(4, 4, 3, 3), requires_grad=False)
    calls, pack = _counter()
    assert C._cached(w, "k", pack) == 1 and C._cached(w, "k", pack) == 1 and len(calls) == 1
    with torch.no_grad():
        w.mul_(2.0)                                           # bumps the version counter
    assert C._cached(w, "k", pack) == 2
    assert C._cached(w, "other key", pack) == 3 and C._cached(w, "k", pack) == 2


def test_weight_seen_trainable_is_never_cached_again():
    """torch.optim.Adam(fused=True) does not bump `_version`; the reference freezes netD only around G's backward."""
    w = torch.nn.Parameter(torch.randn(4, 4, 3, 3))
    calls, pack = _counter()
    w.requires_grad_(False)
    C._cached(w, "k", pack), C._cached(w, "k", pack)
    assert len(calls) == 1                                    # G step: looks frozen
    w.requires_grad_(True)
    C._cached(w, "k", pack)                                   # D step: trainable -> packed fresh, remembered
    w.grad = torch.ones_like(w)
    v = w._version
    torch.optim.Adam([w], lr=0.1, fused=True).step()
    assert w._version == v                                    # the hazard this guards against
    w.requires_grad_(False)
    n = len(calls)
    C._cached(w, "k", pack), C._cached(w, "k", pack)
    assert len(calls) == n + 2                                # next G step: packed again, both times


def test_degenerate_channel_routing_rules(monkeypatch):
    """ffwm_b200/conv.py:_few_forward / _few_dgrad — which shapes go to the direct kernels (csrc/conv_few.cu): <= 4 input channels
    always; ONE output channel at stride 1 on maps >= 32x32 with <= 512 input channels; the data gradient reads the weight
    transposed and flipped with padding k - 1 - p."""
    calls = []
    monkeypatch.setattr(C.ops, "conv_few", lambda *a: calls.append(a[2:4] + a[6:8]))      # (in_major, flip, stride, pad)
    monkeypatch.setattr(C, "FEW", True)
    z = torch.zeros
    assert C._few_forward(z(2, 1, 64, 64), z(96, 1, 5, 5), None, z(2, 96, 64, 64), 1, 2) and calls[-1] == (False, False, 1, 2)
    assert C._few_forward(z(2, 3, 64, 64), z(64, 3, 3, 3), None, z(2, 64, 32, 32), 2, 1) and calls[-1] == (False, False, 2, 1)
    assert C._few_forward(z(2, 96, 64, 64), z(1, 96, 5, 5), None, z(2, 1, 64, 64), 1, 2)
    n = len(calls)
    assert not C._few_forward(z(2, 96, 64, 64), z(3, 96, 3, 3), None, z(2, 3, 64, 64), 1, 1)      # 3 outputs: tensor cores
    assert not C._few_forward(z(2, 96, 16, 16), z(1, 96, 3, 3), None, z(2, 1, 16, 16), 1, 1)      # small map
    assert not C._few_forward(z(2, 1024, 64, 64), z(1, 1024, 3, 3), None, z(2, 1, 64, 64), 1, 1)  # too many input channels
    assert not C._few_forward(z(2, 64, 64, 64), z(64, 64, 3, 3), None, z(2, 64, 64, 64), 1, 1)
    assert len(calls) == n
    # data gradients: conv (Cout, Cin) = (96, 1): grad_input has one channel; (3, 195): grad_out has three
    assert C._few_dgrad(z(2, 96, 64, 64), z(96, 1, 5, 5), z(2, 1, 64, 64), 1, 2) and calls[-1] == (True, True, 1, 2)
    assert C._few_dgrad(z(2, 3, 64, 64), z(3, 195, 3, 3), z(2, 195, 64, 64), 1, 1) and calls[-1] == (True, True, 1, 1)
    assert not C._few_dgrad(z(2, 64, 32, 32), z(64, 3, 3, 3), z(2, 3, 64, 64), 2, 1)              # stride 2: transposed convolution
    assert not C._few_dgrad(z(2, 64, 64, 64), z(64, 3, 3, 3), z(2, 3, 64, 64), 1, 1)              # three outputs
    monkeypatch.setattr(C, "FEW", False)
    assert not C._few_forward(z(2, 1, 64, 64), z(96, 1, 5, 5), None, z(2, 96, 64, 64), 1, 2)
