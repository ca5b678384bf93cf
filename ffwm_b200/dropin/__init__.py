"""Drop-in replacements for the reference's three pybind modules.

`install()` registers them in sys.modules under the reference's names so that
the UNMODIFIED reference file models/external_function.py (`import
resample2d_cuda` etc., :6-12) binds to the sm_100a kernels.
"""
import sys

from . import block_extractor_cuda, local_attn_reshape_cuda, resample2d_cuda

NAMES = ("resample2d_cuda", "block_extractor_cuda", "local_attn_reshape_cuda")


def install():
    sys.modules["resample2d_cuda"] = resample2d_cuda
    sys.modules["block_extractor_cuda"] = block_extractor_cuda
    sys.modules["local_attn_reshape_cuda"] = local_attn_reshape_cuda
