"""Generates tests/golden/ref_guided_filter_f64.npz by importing the REFERENCE's own GuidedFilter / FastGuidedFilter
(models/external_function.py:164-277) read-only from /root/reference in the build container and running them on the
CPU in float64 on seeded inputs (output and gradient with respect to x).  tests/test_guided_filter_math.py compares
ffwm_b200's mirror and the kernel plan of csrc/guided_filter.cu with it.

    python tests/golden/make_golden_guided_filter.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FFWM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

sys.path.insert(0, HERE)
from gf_cases import CASES, fast_inputs, inputs  # noqa: E402


def main():
    from models import external_function as REF_EF          # the reference (its CUDA imports sit in try/except)
    store = {}
    for name, shape, r in CASES:
        x, y, gq = inputs(shape, r)
        x.requires_grad_()
        q = REF_EF.GuidedFilter(r)(x, y)
        q.backward(gq)
        store[name + "/q"], store[name + "/gx"] = q.detach().numpy(), x.grad.numpy()
    # FastGuidedFilter: coefficients at low resolution, applied at high resolution
    lr_x, lr_y, hr_x = fast_inputs()
    store["fast/q"] = REF_EF.FastGuidedFilter(4)(lr_x, lr_y, hr_x).numpy()
    out = os.path.join(HERE, "ref_guided_filter_f64.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, {k: v.shape for k, v in store.items()})


if __name__ == "__main__":
    main()
