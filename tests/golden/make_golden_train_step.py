"""Generates tests/golden/ref_train_step.json by running the REFERENCE's UNMODIFIED
`models.ffwm_model.FFWMModel` (and its losses, networks, LightCNN) from /root/reference on the CPU
in this build container: two `optimize_parameters()` steps on the SURVEY 8(d) cfg3 synthetic batch
(B=2), every network's parameters filled deterministically by state_dict key
(tests/golden/model_cases.py).  Harness shims only: `numpy.int`, a
pre-seeded torchvision VGG19 cache file in a scratch TORCH_HOME, state_dict files at the paths
`opt.lightcnn / opt.flownetf / opt.flownetb`.  Nothing in the reference is edited.

    python tests/golden/make_golden_train_step.py
"""
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("FFWM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
np.int = int
import model_cases as MC  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    scratch = tempfile.mkdtemp(prefix="ffwm_gold_")
    try:
        os.environ["TORCH_HOME"] = scratch
        import torchvision
        vgg = MC.fill_state(torchvision.models.vgg19(weights=None), torch.float32)
        os.makedirs(os.path.join(scratch, "hub", "checkpoints"))
        torch.save(vgg.state_dict(), os.path.join(scratch, "hub", "checkpoints", "vgg19-dcbb9e9d.pth"))
        from models import base_networks as RB
        from lightcnn.light_cnn import LightCNN_29Layers
        paths = {}
        for name, net in (("lightcnn", LightCNN_29Layers()), ("flownetf", RB.FlowNet(64)), ("flownetb", RB.FlowNet(64))):
            paths[name] = os.path.join(scratch, name + ".pth")
            torch.save(MC.fill_state(net, torch.float32).state_dict(), paths[name])
        opt = types.SimpleNamespace(gpu_ids=[], isTrain=True, checkpoints_dir=scratch, name="gold", preprocess="none",
                                    crop=False, **paths)
        from models.ffwm_model import FFWMModel
        model = FFWMModel(opt)
        MC.fill_state(model.netG, torch.float32)
        MC.fill_state(model.netD, torch.float32)
        from oracle.train_cpu import synthetic_batch
        out = {"steps": []}
        for step in range(2):
            model.set_train_input(synthetic_batch(2, seed=500 + step))
            model.optimize_parameters()
            out["steps"].append({k: float(v) for k, v in model.get_current_losses().items()})
        probes = {"netG": "rec2.0.weight_orig", "netD": "nets.0.0.weight_orig", "flowNetF": "conv0.0.weight", "flowNetB": "predict_flow0.0.weight"}
        out["params_after"] = {n: MC.sub(dict(getattr(model, n).named_parameters())[k]).tolist()[:64] for n, k in probes.items()}
        out["probes"] = probes
        json.dump(out, open(os.path.join(HERE, "ref_train_step.json"), "w"), indent=1)
        print(json.dumps(out["steps"], indent=1))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
