import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_warp():
    from oracle import warp
    warp.build()
    return warp


@pytest.fixture(scope="session")
def ref_cuda():
    """The reference's own CUDA extensions, prebuilt into oracle/_ref by
    oracle/build_ref.py (GPU box only uses the prebuilt files)."""
    import torch  # noqa: F401  (pybind modules need libtorch loaded)
    d = os.path.join(ROOT, "oracle", "_ref")
    if not all(os.path.exists(os.path.join(d, n + ".so")) for n in
               ("resample2d_cuda", "block_extractor_cuda", "local_attn_reshape_cuda")):
        pytest.skip("oracle/_ref not built")
    import importlib.util
    mods = {}
    for n in ("resample2d_cuda", "block_extractor_cuda", "local_attn_reshape_cuda"):
        spec = importlib.util.spec_from_file_location(n, os.path.join(d, n + ".so"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[n] = m
    return mods
