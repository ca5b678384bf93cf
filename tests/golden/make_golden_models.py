"""Generates tests/golden/ref_models_f64.npz by importing the REFERENCE's own Python modules
(models/base_networks.py, lightcnn/light_cnn.py) read-only from /root/reference in the build
container and running them on the CPU in float64 with parameters filled deterministically by
state_dict key (tests/golden/model_cases.py).  The GPU suite rebuilds the same parameters into
ffwm_b200's networks and compares.  Harness shims only: numpy.int.

    python tests/golden/make_golden_models.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get("FFWM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
np.int = int            # models/base_networks.py:366 uses the removed alias
import model_cases as MC  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    from models import base_networks as RB           # the reference
    from lightcnn.light_cnn import LightCNN_29Layers as RefLightCNN
    store = {}

    def put(prefix, d):
        for k, v in d.items():
            store[prefix + "/" + k] = v

    put("flownet16", MC.run_flownet(MC.fill_state(RB.FlowNet(16))))
    put("netD", MC.run_netd(MC.fill_state(RB.MSDiscriminator(128, sigmoid=False))))
    put("lightcnn", MC.run_lightcnn(MC.fill_state(RefLightCNN(num_classes=100))))
    put("netG", MC.run_netg(MC.fill_state(RB.FFWM(sn=True))))
    # the checkpoint contract: state_dict keys and shapes of the networks at their production sizes
    import json
    keys = {}
    for name, net in (("FlowNet64", RB.FlowNet(64)), ("FFWM_sn", RB.FFWM(sn=True)),
                      ("MSDiscriminator128", RB.MSDiscriminator(128, sigmoid=False)), ("LightCNN_29Layers", RefLightCNN())):
        keys[name] = {k: list(v.shape) for k, v in net.state_dict().items()}
    json.dump(keys, open(os.path.join(HERE, "ref_state_keys.json"), "w"))
    np.savez_compressed(os.path.join(HERE, "ref_models_f64.npz"), **store)
    print("wrote", len(store), "arrays", sum(v.nbytes for v in store.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
