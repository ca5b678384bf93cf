"""The reference's orchestrators import on top of ffwm_b200's modules (drop-in tier 2).  Needs the
reference checkout, which only exists in the build container: skipped elsewhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FFWM_REFERENCE", os.path.join(os.sep, "root", "reference"))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_models_bind_to_ffwm_b200():
    code = r"""
import sys, numpy as np
np.int = int
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ffwm_b200
import models                      # the reference package (its __init__ only defines helpers)
ffwm_b200.compat.install()
from models import ffwm_model, flownet_model          # UNMODIFIED reference orchestrators
assert ffwm_model.base_networks is ffwm_b200.base_networks
assert ffwm_model.losses is ffwm_b200.losses
assert ffwm_model.external_function is ffwm_b200.external_function
assert ffwm_model.LightCNN_29Layers is ffwm_b200.light_cnn.LightCNN_29Layers
assert flownet_model.losses.BlockExtractor is ffwm_b200.external_function.BlockExtractor
import resample2d_cuda, block_extractor_cuda, local_attn_reshape_cuda
assert resample2d_cuda.forward.__module__.startswith('ffwm_b200.dropin')
print('ok')
""" % (ROOT, REF)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
