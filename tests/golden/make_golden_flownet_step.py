"""Generates tests/golden/ref_flownet_step.json: the REFERENCE's `models.flownet_model.FlowNetModel`
(FlowNet pre-training: perceptual-correctness + affine-regularisation + landmark losses) run on the
CPU in this build container for two `optimize_parameters()` steps.

The reference has no CPU implementation of block_extractor / local_attn_reshape
(NotImplementedError for CPU tensors), so — harness shims only, nothing in the reference is
edited — the `extractor` / `reshape` attributes of its AffineRegularizationLoss INSTANCES are
replaced by callables backed by the C oracle (which is pinned to the outputs of the reference's own
CUDA kernels, tests/golden/ref_cuda_*.npz), `criterionLD` gets the torch-1.5 integer division
(SURVEY 8c), and VGG19 comes from a pre-seeded cache file.

    python tests/golden/make_golden_flownet_step.py
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
        from oracle import train_cpu
        from models.flownet_model import FlowNetModel
        opt = types.SimpleNamespace(gpu_ids=[], isTrain=True, checkpoints_dir=scratch, name="gold", preprocess="none")
        model = FlowNetModel(opt)
        model.reverse = False                      # train_flow.py sets it from --reverse
        MC.fill_state(model.flowNet, torch.float32)
        class _Extract(torch.nn.Module):
            def __init__(self, kz):
                super().__init__()
                self.kz = kz

            def forward(self, s, f):
                return train_cpu._BlockExtractorCPU.apply(s, f, self.kz)

        class _Reshape(torch.nn.Module):
            def forward(self, x, k):
                return train_cpu._LocalAttnReshapeCPU.apply(x, k)

        for reg in model.Regularization.method_dic.values():
            reg.extractor = _Extract(reg.kz)
            reg.reshape = _Reshape()
        inner = model.criterionLD.criterionLD

        def ld_forward(flows, lm_S, lm_F, gate, self=model.criterionLD):
            total = 0
            for i, flow in enumerate(flows):
                scale = self.img_size // flow.size(3)
                total += self.weights[i] * inner(flow, torch.div(lm_S, scale, rounding_mode='floor'),
                                                 torch.div(lm_F, scale, rounding_mode='floor'), gate)
            return total
        model.criterionLD = ld_forward
        out = {"steps": []}
        for step in range(2):
            model.set_train_input(train_cpu.synthetic_batch(1, seed=700 + step))
            model.optimize_parameters()
            out["steps"].append({k: float(getattr(model, k)) for k in ("loss", "loss_reg", "loss_lm", "loss_cor")})
        probe = "predict_flow0.0.weight"
        out["probe"] = probe
        out["param_after"] = MC.sub(dict(model.flowNet.named_parameters())[probe]).tolist()[:64]
        json.dump(out, open(os.path.join(HERE, "ref_flownet_step.json"), "w"), indent=1)
        print(json.dumps(out["steps"], indent=1))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
