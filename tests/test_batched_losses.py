"""PerceptualLoss.many / IdentityLoss.many (one network pass per resolution over all generated images, one over all
targets) against the pair-by-pair calls of the reference's forward: same losses, same gradients."""
import torch

from ffwm_b200 import losses
from ffwm_b200.light_cnn import LightCNN_29Layers


def test_perceptual_many_matches_pairwise_calls():
    torch.manual_seed(0)
    crit = losses.PerceptualLoss().double()
    g = torch.Generator().manual_seed(1)
    shapes = [(2, 3, 32, 32), (1, 3, 48, 48), (2, 3, 32, 32), (3, 3, 32, 32)]
    xs = [torch.rand(*s, generator=g, dtype=torch.float64, requires_grad=True) for s in shapes]
    ys = [torch.rand(*s, generator=g, dtype=torch.float64) for s in shapes]
    want = [crit(x, y) for x, y in zip(xs, ys)]
    sum(w * (i + 1) for i, w in enumerate(want)).backward()
    gw = [x.grad.clone() for x in xs]
    for x in xs:
        x.grad = None
    got = crit.many(list(zip(xs, ys)))
    sum(w * (i + 1) for i, w in enumerate(got)).backward()
    for a, b in zip(got, want):
        assert abs(float(a) - float(b)) <= 1e-12 * max(1.0, abs(float(b)))
    for x, gref in zip(xs, gw):
        assert (x.grad - gref).abs().max() <= 1e-12 * max(1.0, float(gref.abs().max()))


def test_identity_many_matches_pairwise_calls():
    torch.manual_seed(0)
    net = LightCNN_29Layers(num_classes=10).double().eval()
    for p in net.parameters():
        p.requires_grad_(False)
    crit = losses.IdentityLoss(net)
    g = torch.Generator().manual_seed(2)
    a, b = (torch.rand(2, 3, 128, 128, generator=g, dtype=torch.float64, requires_grad=True) for _ in range(2))
    t = torch.rand(2, 3, 128, 128, generator=g, dtype=torch.float64)
    want = [crit(a, t), crit(b, t)]
    (want[0] + 2 * want[1]).backward()
    ga, gb = a.grad.clone(), b.grad.clone()
    a.grad = b.grad = None
    got = crit.many([(a, t), (b, t)])
    (got[0] + 2 * got[1]).backward()
    assert all(abs(float(x) - float(y)) <= 1e-12 for x, y in zip(got, want))
    assert (a.grad - ga).abs().max() <= 1e-12 and (b.grad - gb).abs().max() <= 1e-12
