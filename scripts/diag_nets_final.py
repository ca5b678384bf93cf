"""Whole-network fp32 errors of the SHIPPED configuration on the GPU against the reference's float64 goldens (outputs and
gradients per probed key), next to the library in strict fp32 and in TF32:  python scripts/diag_nets_final.py
'product' = every hand-written kernel on (tcgen05 / direct convolutions, batch norm + LeakyReLU + residual, spectral norm,
pooling, MFM); LightCNN additionally with the 3xBF16 forwards it gets as a frozen loss network."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import model_cases as MC  # noqa: E402
from ffwm_b200 import base_networks as B, conv, light_cnn as L, norm, pool, spectral  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_models_f64.npz"))
dev = torch.device("cuda", 0)


def gold(prefix):
    return {k[len(prefix) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(prefix + "/")}


def run(which, lossnet):
    if which == "flownet16":
        return MC.run_flownet(MC.fill_state(B.FlowNet(16), torch.float32).to(dev))
    if which == "netD":
        return MC.run_netd(MC.fill_state(B.MSDiscriminator(128, sigmoid=False), torch.float32).to(dev))
    if which == "lightcnn":
        net = MC.fill_state(L.LightCNN_29Layers(num_classes=100), torch.float32).to(dev)
        if lossnet:
            conv.set_forward_math(net, conv.LOSSNET_MATH_FWD)
        return MC.run_lightcnn(net)
    return MC.run_netg(MC.fill_state(B.FFWM(sn=True), torch.float32).to(dev))


CONFIGS = [("library fp32", False, False, False), ("library TF32 (torch default)", False, True, False), ("product", True, False, False),
           ("product, as a frozen loss network", True, False, True)]
for which in ["flownet16", "netD", "lightcnn", "netG"]:
    want = gold(which)
    for name, on, tf32, lossnet in CONFIGS:
        if lossnet and which != "lightcnn":
            continue
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
        conv.ENABLED = norm.ENABLED = pool.ENABLED = spectral.FUSED_SN = conv.FEW = L.FUSED_MFM = on
        got = run(which, lossnet)
        errs = {k: float(np.abs(got[k] - want[k]).max() / max(np.abs(want[k]).max(), 1e-30)) for k in want}
        print("%-10s %-34s %s" % (which, name, "  ".join("%s=%.1e" % (k.split("/")[-1][:18], v) for k, v in errs.items())), flush=True)
