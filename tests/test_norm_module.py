"""Host side of ffwm_b200/norm.py on the CPU: the fused module layout keeps the reference's state_dict keys and computes
conv -> BatchNorm2d -> LeakyReLU exactly as torch does (torch's own kernels run here; the CUDA kernels: test_batch_norm_gpu.py)."""
import functools

import torch
from torch import nn

from ffwm_b200 import base_networks as B
from ffwm_b200.norm import AbsorbedLeakyReLU, BatchNorm2d, as_product_norm, fuse_activations


def test_fuse_activations_keeps_keys_and_function():
    torch.manual_seed(0)
    ref = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.BatchNorm2d(8), nn.LeakyReLU(0.2, True))
    mods = fuse_activations([nn.Conv2d(3, 8, 3, padding=1), BatchNorm2d(8), nn.LeakyReLU(0.2, True)])
    assert isinstance(mods[2], AbsorbedLeakyReLU) and mods[1].act_slope == 0.2
    mine = nn.Sequential(*mods)
    assert list(mine.state_dict()) == list(ref.state_dict())
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(2, 3, 9, 9)
    assert torch.equal(mine(x), ref(x))                      # training mode
    assert torch.equal(mine.state_dict()["1.running_var"], ref.state_dict()["1.running_var"])
    ref.eval(), mine.eval()
    assert torch.equal(mine(x), ref(x))


def test_residual_block_tail_equals_the_spelled_out_form():
    torch.manual_seed(1)
    blk = B.ResidualBlock(6, 6, sn=True)
    assert isinstance(blk.blocks[-1], BatchNorm2d) and blk.blocks[-1].act_slope is None and blk.blocks[1].act_slope == B.LRELU_SLOPE
    x = torch.randn(2, 6, 8, 8)
    blk.eval()                                               # (spectral norm does not iterate in eval mode)
    want = blk.activ(blk.blocks(x) + blk.input(x))
    assert torch.allclose(blk(x), want, rtol=0, atol=0)


def test_norm_factories_of_the_reference_map_to_the_product_class():
    assert as_product_norm(nn.BatchNorm2d) is BatchNorm2d
    p = as_product_norm(functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True))
    assert isinstance(p(4), BatchNorm2d)
    assert as_product_norm(nn.InstanceNorm2d) is nn.InstanceNorm2d
    net = B.FlowNet(4)
    assert all(isinstance(m, BatchNorm2d) for m in net.modules() if isinstance(m, nn.BatchNorm2d))
    assert not any(type(m) is nn.LeakyReLU for m in net.modules())


def test_deferred_counters_add_up_like_torch():
    from ffwm_b200.norm import DeferredCounters
    torch.manual_seed(2)
    a = nn.Sequential(nn.Conv2d(3, 4, 1), BatchNorm2d(4))
    b = nn.Sequential(nn.Conv2d(3, 4, 1), BatchNorm2d(4))
    b.load_state_dict(a.state_dict())
    ctr = DeferredCounters([b])
    x = torch.randn(2, 3, 5, 5)
    for _ in range(3):
        a(x), b(x)
    # (on the CPU the layer takes torch's path, which counts by itself: deferral only applies to the CUDA kernels' path)
    ctr.flush()
    assert int(a[1].num_batches_tracked) == int(b[1].num_batches_tracked) == 3


def test_deep_copies_leave_the_deferred_counters():
    import copy
    from ffwm_b200.norm import DeferredCounters
    net = nn.Sequential(nn.Conv2d(3, 4, 1), BatchNorm2d(4))
    DeferredCounters([net])
    twin = copy.deepcopy(net)
    assert net[1]._deferred is not None and twin[1]._deferred is None
    assert list(twin.state_dict()) == list(net.state_dict()) and twin[1].weight is not net[1].weight
    twin(torch.randn(2, 3, 4, 4))
    assert int(twin[1].num_batches_tracked) == 1
