"""Generates tests/golden/ref_orchestrators.json: the reference's UNMODIFIED orchestrators at the
BASELINE batch sizes, run on the CPU in this build container from the byte-identical staged copy
(baseline/_ref, baseline/stage_ref.py) through baseline/ref_harness.py (harness shims only):

  * `FFWMModel.optimize_parameters()` x 2 at batch 8 (BASELINE config 3), seeds 800, 801
  * `FlowNetModel.optimize_parameters()` x 2 at batch 6 (BASELINE config 2), seeds 810, 811

tests/test_orchestrators_gpu.py steps the SAME reference classes on the sm_100a product
(`ffwm_b200.compat.install()`) on a B200 and compares.

    python tests/golden/make_golden_orchestrators.py
"""
import json
import os
import shutil
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from baseline import ref_harness as H  # noqa: E402

SUB = 11
FFWM_PROBES = {"netG": "rec2.0.weight_orig", "netD": "nets.0.0.weight_orig", "flowNetF": "conv0.0.weight",
               "flowNetB": "predict_flow0.0.weight"}
FLOW_PROBE = "predict_flow0.0.weight"


def sub(t):
    return t.detach().double().cpu().reshape(-1)[::SUB][:64].tolist()


def main():
    torch.set_num_threads(os.cpu_count())
    scratch = tempfile.mkdtemp(prefix="ffwm_gold_")
    out = {}
    try:
        model = H.reference_ffwm_model("cpu", scratch=scratch)
        steps = []
        for step in range(2):
            model.set_train_input(H.synthetic_batch(8, seed=800 + step))
            model.optimize_parameters()
            steps.append({k: float(v) for k, v in model.get_current_losses().items()})
        out["ffwm_b8"] = {"steps": steps, "probes": FFWM_PROBES,
                          "params_after": {n: sub(dict(getattr(model, n).named_parameters())[k]) for n, k in FFWM_PROBES.items()}}
        print(json.dumps(steps, indent=1))
        del model
        model = H.reference_flownet_model("cpu", scratch=scratch)
        steps = []
        for step in range(2):
            model.set_train_input(H.synthetic_batch(6, seed=810 + step))
            model.optimize_parameters()
            steps.append({k: float(getattr(model, k)) for k in ("loss", "loss_reg", "loss_lm", "loss_cor")})
        out["flownet_b6"] = {"steps": steps, "probe": FLOW_PROBE,
                             "param_after": sub(dict(model.flowNet.named_parameters())[FLOW_PROBE])}
        print(json.dumps(steps, indent=1))
        json.dump(out, open(os.path.join(HERE, "ref_orchestrators.json"), "w"), indent=1)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
