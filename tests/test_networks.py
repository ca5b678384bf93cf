"""Network rows of the hot path: ffwm_b200's FlowNet / FFWM generator /
MSDiscriminator / LightCNN-29 against golden vectors produced by the REFERENCE's own modules
(tests/golden/make_golden_models.py: float64, CPU, parameters filled by state_dict key)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import model_cases as MC  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_models_f64.npz"))
KEYS = json.load(open(os.path.join(HERE, "golden", "ref_state_keys.json")))


def gold(prefix):
    return {k[len(prefix) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(prefix + "/")}


def check(got, want, rtol):
    assert set(got) == set(want)
    for k in want:
        scale = max(float(np.abs(want[k]).max()), 1e-30)
        err = float(np.abs(got[k] - want[k]).max()) / scale
        assert err <= rtol, "%s: rel err %.3e > %.1e" % (k, err, rtol)


@pytest.fixture(scope="module")
def nets():
    from ffwm_b200 import base_networks, light_cnn
    return base_networks, light_cnn


# ------------------------------------------------------------------ checkpoint contract (CPU)
def test_state_dict_keys_match_the_reference(nets):
    B, L = nets
    mine = {"FlowNet64": B.FlowNet(64), "FFWM_sn": B.FFWM(sn=True),
            "MSDiscriminator128": B.MSDiscriminator(128, sigmoid=False), "LightCNN_29Layers": L.LightCNN_29Layers()}
    for name, net in mine.items():
        got = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert list(got) == list(KEYS[name]), name          # same keys in the same order
        assert got == KEYS[name], name


def test_parameter_counts(nets):
    B, L = nets
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert abs(n(B.FlowNet(64)) / 1e6 - 52.36) < 0.01       # SURVEY 6 [probe]
    assert abs(n(B.FFWM(sn=True)) / 1e6 - 16.25) < 0.01
    assert abs(n(B.MSDiscriminator(128, sigmoid=False)) / 1e6 - 1.12) < 0.01
    assert B.MSDiscriminator(128).max_n_scales == 3


# ------------------------------------------------------------------ conv-only nets run on CPU too
@pytest.mark.parametrize("which", ["flownet16", "netD", "lightcnn"])
def test_conv_networks_match_reference_cpu_f64(nets, which):
    B, L = nets
    if which == "flownet16":
        got = MC.run_flownet(MC.fill_state(B.FlowNet(16)))
    elif which == "netD":
        got = MC.run_netd(MC.fill_state(B.MSDiscriminator(128, sigmoid=False)))
    else:
        got = MC.run_lightcnn(MC.fill_state(L.LightCNN_29Layers(num_classes=100)))
    check(got, gold(which), 1e-9)


def test_generator_refuses_cpu_tensors(nets):
    B, _ = nets
    g = B.FFWM(sn=True)
    x, flows = MC.netg_inputs(torch.zeros(1))
    with pytest.raises(NotImplementedError):      # the warp is CUDA-only, as in the reference's ops
        g(x.float(), flow=[f.float() for f in flows])


# ------------------------------------------------------------------ GPU: every net, f64 and f32
# fp32 tolerances relative to max|ref| (reference = the reference's modules in float64 on the CPU), set from the
# measurement on a B200 in profiles/r02h_network_accuracy.txt — (outputs, gradients), with what was measured:
#                product (tcgen05, 3xTF32 forwards / 3xBF16 gradients)    cuDNN strict fp32      cuDNN TF32 (torch default)
#   flownet16    1.2e-5 , 2.0e-5                                          1.4e-5 , 8.5e-6        4.5e-3 , 7.7e-2
#   netD         1.5e-6 , 1.2e-5                                          1.3e-6 , 5.8e-7        5.2e-4 , 4.0e-2
#   lightcnn     3.7e-5 , 4.1e-2                                          8.7e-7 , 5.7e-3        7.8e-4 , 1.3e-1
#   netG         5.3e-5 , 2.5e-2                                          3.1e-5 , 1.4e-2        5.5e-3 , 1.8e-1
# FlowNet and the discriminator are well conditioned and held to 1e-4 — the path's contract — end to end, gradients
# included (30 conv + BatchNorm(batch of 2) layers of back-propagation).  LightCNN's max-feature-map is piecewise
# linear: rounding flips a few max selections and reroutes their gradient, and the generator's gradients pass through
# ~60 spectral-normed conv + BatchNorm layers; there even cuDNN's strict fp32 is 0.6-1.4e-2 away from float64, so those
# gradients are held to a few times that.  LightCNN has no normalisation layers, so the tensor core's truncating fp32
# accumulation (a systematic ~1e-6 shrink per layer) adds up over its 29 layers: outputs 4e-5.  The float64 runs are
# the tight check (1e-8; float64 never takes the tensor-core path).
F32_TOL = {"flownet16": (1e-4, 1e-4), "netD": (1e-4, 1e-4), "lightcnn": (1e-4, 8e-2), "netG": (2e-4, 5e-2)}


def check_split(got, want, tol_out, tol_grad):
    assert set(got) == set(want)
    for k in want:
        scale = max(float(np.abs(want[k]).max()), 1e-30)
        err = float(np.abs(got[k] - want[k]).max()) / scale
        rtol = tol_grad if k.startswith("grad/") else tol_out
        assert err <= rtol, "%s: rel err %.3e > %.1e" % (k, err, rtol)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("which", ["flownet16", "netD", "lightcnn", "netG"])
def test_networks_match_reference_on_gpu(nets, which, dt):
    B, L = nets
    tol_out, tol_grad = (1e-8, 1e-8) if dt == torch.float64 else F32_TOL[which]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    if which == "flownet16":
        got = MC.run_flownet(MC.fill_state(B.FlowNet(16), dt).to(dev))
    elif which == "netD":
        got = MC.run_netd(MC.fill_state(B.MSDiscriminator(128, sigmoid=False), dt).to(dev))
    elif which == "lightcnn":
        got = MC.run_lightcnn(MC.fill_state(L.LightCNN_29Layers(num_classes=100), dt).to(dev))
    else:
        got = MC.run_netg(MC.fill_state(B.FFWM(sn=True), dt).to(dev))
    check_split(got, gold(which), tol_out, tol_grad)
