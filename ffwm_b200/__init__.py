"""ffwm_b200 — B200-native (sm_100a) implementation of FFWM's flow-warping hot path.

Importing the package loads libffwm_b200.so (hand-written CUDA behind the C
ABI in include/ffwm_b200.h) and fails loudly if it has not been built: there
is no CPU or eager-PyTorch fallback for the kernels.
"""
from . import _lib

_lib.lib()   # raises ImportError when the library is missing or stale

from . import ops, external_function, dropin  # noqa: E402
from . import conv, base_networks, light_cnn, losses, parallel, train_step, compat  # noqa: E402
from .dropin import install as install_dropin  # noqa: E402

__all__ = ["ops", "external_function", "dropin", "install_dropin", "base_networks", "light_cnn", "losses",
           "parallel", "train_step", "compat", "conv"]
