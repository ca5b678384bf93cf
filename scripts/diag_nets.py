"""Per-key fp32 errors of the mirrored networks on the GPU against the reference's float64 goldens, under different
kernel selections (diagnostic for the tolerances of tests/test_networks.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import model_cases as MC  # noqa: E402
from ffwm_b200 import _lib, base_networks as B, conv, light_cnn as L  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_models_f64.npz"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)


def gold(prefix):
    return {k[len(prefix) + 1:]: GOLD[k] for k in GOLD.files if k.startswith(prefix + "/")}


def run(which):
    if which == "flownet16":
        return MC.run_flownet(MC.fill_state(B.FlowNet(16), torch.float32).to(dev))
    if which == "netD":
        return MC.run_netd(MC.fill_state(B.MSDiscriminator(128, sigmoid=False), torch.float32).to(dev))
    if which == "lightcnn":
        return MC.run_lightcnn(MC.fill_state(L.LightCNN_29Layers(num_classes=100), torch.float32).to(dev))
    return MC.run_netg(MC.fill_state(B.FFWM(sn=True), torch.float32).to(dev))


CONFIGS = [("library fp32", dict(ENABLED=False)),
           ("library TF32 (torch default)", dict(ENABLED=False, TF32=True)),
           ("3x3: tcgen05 fwd, library dgrad", dict(ENABLED=True, GENERAL=False, WGRAD_GEN=False, _DIAG_DGRAD_LIB=True)),
           ("3x3: library fwd, tcgen05 dgrad", dict(ENABLED=True, GENERAL=False, WGRAD_GEN=False, _DIAG_FWD_LIB=True)),
           ("3x3 tcgen05, 3xBF16 forwards", dict(ENABLED=True, GENERAL=False, WGRAD_GEN=False, MATH=1)),
           ("everything, 3xBF16 forwards", dict(ENABLED=True, GENERAL=True, WGRAD_GEN=True, MATH=1)),
           ("3x3 tcgen05 only", dict(ENABLED=True, GENERAL=False, WGRAD_GEN=False)),
           ("3x3 + general fwd/dgrad", dict(ENABLED=True, GENERAL=True, WGRAD_GEN=False)),
           ("product default (all convolutions on tcgen05)", dict(ENABLED=True, GENERAL=True, WGRAD_GEN=True))]
for which in sys.argv[1:] or ["flownet16", "netD", "lightcnn", "netG"]:
    want = gold(which)
    for name, cfg in CONFIGS:
        conv._DIAG_DGRAD_LIB = conv._DIAG_FWD_LIB = False
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = bool(cfg.get("TF32"))
        conv.MATH_FWD = cfg.get("MATH", 0)
        for k, v in cfg.items():
            if k not in ("TF32", "MATH"):
                setattr(conv, k, v)
        got = run(which)
        errs = {k: float(np.abs(got[k] - want[k]).max() / max(np.abs(want[k]).max(), 1e-30)) for k in want}
        print("%-10s %-30s %s" % (which, name, "  ".join("%s=%.1e" % (k.split("/")[-1][:18], v) for k, v in errs.items())), flush=True)
