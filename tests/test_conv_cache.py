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
