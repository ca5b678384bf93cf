"""The FFWM train step against the UNMODIFIED reference `FFWMModel`, whose losses
over two optimisation steps were recorded on the CPU by tests/golden/make_golden_train_step.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import model_cases as MC  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "ref_train_step.json")))


def build_trainer(device):
    from ffwm_b200 import base_networks as B, light_cnn as L
    from ffwm_b200.train_step import FFWMTrainer
    states = {"lightcnn_state": MC.fill_state(L.LightCNN_29Layers(), torch.float32).state_dict(),
              "flownetf_state": MC.fill_state(B.FlowNet(64), torch.float32).state_dict(),
              "flownetb_state": MC.fill_state(B.FlowNet(64), torch.float32).state_dict()}
    tr = FFWMTrainer(device, **states)
    tr.criterionPerceptual.vgg.load_torchvision(MC.vgg_torchvision_state(tr.criterionPerceptual.vgg))
    tr.criterionPerceptual.to(device)
    MC.fill_state(tr.netG, torch.float32)
    MC.fill_state(tr.netD, torch.float32)
    return tr


def run_two_steps(tr):
    from oracle.train_cpu import synthetic_batch
    got = []
    for step in range(2):
        tr.set_input(synthetic_batch(2, seed=500 + step))
        tr.optimize_parameters()
        got.append(tr.get_current_losses())
    return got


def compare(tr, got, rtol):
    for g, w in zip(got, GOLD["steps"]):
        for k, v in w.items():
            assert abs(g[k] - v) <= rtol * max(abs(v), 1e-3), (k, g[k], v)
    for net, key in GOLD["probes"].items():
        p = MC.sub(dict(getattr(tr, net).named_parameters())[key])[:64]
        want = np.array(GOLD["params_after"][net])
        assert np.abs(p - want).max() <= rtol * max(np.abs(want).max(), 1e-6) * 10, net


def test_train_step_matches_reference_on_cpu_ops():
    """Host logic (loss weights, order of passes, optimisers) with every warp served by the CPU
    oracle: must reproduce the reference's CPU run almost exactly (same PyTorch CPU kernels)."""
    from oracle import train_cpu
    torch.manual_seed(0)
    with train_cpu.cpu_ops():
        tr = build_trainer("cpu")
        got = run_two_steps(tr)
    compare(tr, got, 2e-4)


@pytest.mark.gpu
def test_train_step_matches_reference_on_gpu():
    """The product path: cuDNN fp32 convolutions + the hand-written warp kernels on cuda:0."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from ffwm_b200 import _lib
    n0 = _lib.kernel_launches()
    tr = build_trainer("cuda:0")
    got = run_two_steps(tr)
    assert _lib.kernel_launches() - n0 >= 2 * 16          # 16 grid warps per step went through the C ABI
    compare(tr, got, 5e-3)


@pytest.mark.gpu
def test_cuda_graph_replay_matches_eager_launches():
    """The captured step (benchmarks use it) must compute what the eager step computes.  Two eager
    runs already differ from each other (atomic scatter order, cuDNN algorithm choice, and Adam's
    sign-like first updates amplify the noise), so the graph is held to that measured run-to-run
    spread, not to bit equality."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from ffwm_b200.train_step import FFWMTrainer
    from oracle.train_cpu import synthetic_batch
    batches = [synthetic_batch(2, seed=900 + i) for i in range(3)]

    def run(graph, segmented=False):
        torch.manual_seed(3)
        tr = FFWMTrainer("cuda:0", graph=graph)
        if graph:
            tr.enable_cuda_graph(batches[0], warmup=3, segmented=segmented)   # 3 real steps on batches[0]
            assert tr.graph_kernel_nodes >= 16
        else:
            for _ in range(3):
                tr.step(batches[0])
        out = []
        for b in batches[1:]:
            tr.step(b)
            out.append(tr.get_current_losses())
        return out

    e1, e2, g, gs = run(False), run(False), run(True), run(True, segmented=True)   # gs: the data-parallel layout
    for a, b, c, d in zip(e1, e2, g, gs):
        for k in a:
            spread = abs(a[k] - b[k])
            # the LightCNN identity loss (max-feature-map routing, magnitude 1e-2) and the VGG perceptual term that dominates
            # loss_G amplify the noise most: two EAGER runs were measured 5 % apart in loss_G after these five steps
            # (profiles/r02v_graph_vs_eager.txt: 10.947 vs 11.507), while loss_l1 / loss_adv / loss_D agree to 3-4 digits
            floor = {"loss_iden": 1e-1, "loss_G": 1e-1, "loss_prc": 1e-1, "loss_fc": 1e-1}.get(k, 2e-2)
            tol = max(5 * spread, floor * max(abs(a[k]), 1e-3))
            assert abs(a[k] - c[k]) <= tol, ("one graph", k, a[k], b[k], c[k])
            assert abs(a[k] - d[k]) <= tol, ("three graphs", k, a[k], b[k], d[k])
