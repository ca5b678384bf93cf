"""FlowNet pre-training step (the only live user of block_extractor /
local_attn_reshape) against the reference's `FlowNetModel`, recorded on the CPU by
tests/golden/make_golden_flownet_step.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import model_cases as MC  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "ref_flownet_step.json")))


def build(device):
    from ffwm_b200.train_step import FlowNetTrainer
    tr = FlowNetTrainer(device)
    MC.fill_state(tr.flowNet, torch.float32)
    tr.Correctness.vgg.load_torchvision(MC.vgg_torchvision_state(tr.Correctness.vgg))
    tr.Correctness.to(device)
    return tr


def run(tr):
    from oracle.train_cpu import synthetic_batch
    got = []
    for step in range(2):
        tr.set_input(synthetic_batch(1, seed=700 + step))
        tr.optimize_parameters()
        got.append(tr.get_current_losses())
    return got


def compare(tr, got, rtol):
    for g, w in zip(got, GOLD["steps"]):
        for k, v in w.items():
            assert abs(g[k] - v) <= rtol * max(abs(v), 1e-3), (k, g[k], v)
    p = MC.sub(dict(tr.flowNet.named_parameters())[GOLD["probe"]])[:64]
    want = np.array(GOLD["param_after"])
    assert np.abs(p - want).max() <= 10 * rtol * max(np.abs(want).max(), 1e-6)


def test_flownet_step_matches_reference_on_cpu_ops():
    from oracle import train_cpu
    with train_cpu.cpu_ops():
        tr = build("cpu")
        got = run(tr)
    compare(tr, got, 2e-4)


@pytest.mark.gpu
def test_flownet_step_matches_reference_on_gpu():
    """Product path: block_extractor, local_attn_reshape and the grid warp on the sm_100a kernels."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from ffwm_b200 import _lib
    n0 = _lib.kernel_launches()
    tr = build("cuda:0")
    got = run(tr)
    assert _lib.kernel_launches() - n0 >= 2 * (6 + 6 + 6 + 6 + 4)     # K4..K7 x 6 per step + warps
    compare(tr, got, 5e-3)
