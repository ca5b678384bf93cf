"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU,
exports every symbol include/ffwm_b200.h declares, the ctypes binding covers
each of them, and argument validation fails cleanly (no kernel is launched)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "ffwm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffwm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ffwm_b200 import _lib
    lib = _lib.lib()
    names = _header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "libffwm_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert lib.ffwm_abi_version() == _lib.ABI_VERSION


def test_header_cites_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "ffwm_b200.h")).read()
    for cite in ("resample2d_cuda.cc", "block_extractor_cuda.cc", "local_attn_reshape_cuda.cc", "base_networks.py:168-173"):
        assert cite in text


def test_argument_validation_without_gpu():
    from ffwm_b200 import _lib
    lib = _lib.lib()
    a = torch.zeros(1, 3, 4, 4)
    flow3 = torch.zeros(1, 3, 4, 4)
    out = torch.zeros(1, 3, 4, 4)
    null = ctypes.c_void_p(0)
    # wrong channel count on the flow
    rc = lib.ffwm_block_extractor_forward(_lib.t4(a), _lib.t4(flow3), _lib.t4(out), 3, 0, null)
    assert rc == -2 and b"2 channels" in lib.ffwm_last_error()
    # bad dtype code
    rc = lib.ffwm_resample2d_forward(_lib.t4(a), _lib.t4(flow3), _lib.t4(out), 2, 1, 7, null)
    assert rc == -3
    # negative kernel size
    rc = lib.ffwm_resample2d_forward(_lib.t4(a), _lib.t4(flow3), _lib.t4(out), -2, 1, 0, null)
    assert rc == -3
    # k*k channel rule of local_attn_reshape (external_function.py:77)
    rc = lib.ffwm_local_attn_reshape_forward(_lib.t4(a), _lib.t4(torch.zeros(1, 1, 12, 12)), 3, 0, null)
    assert rc == -2
    # null data pointer
    d = _lib.t4(a)
    d.data = None
    rc = lib.ffwm_grid_warp_forward(d, _lib.t4(torch.zeros(1, 2, 4, 4)), _lib.t4(out), 0, null)
    assert rc == -1
    # empty tensors are a successful no-op (no launch)
    e = torch.zeros(0, 3, 4, 4)
    rc = lib.ffwm_grid_warp_forward(_lib.t4(e), _lib.t4(torch.zeros(0, 2, 4, 4)), _lib.t4(e), 0, null)
    assert rc == 0


def test_conv_entry_points_validate_without_gpu():
    """The tensor-core convolution entry points reject bad tiles / shapes before any launch."""
    from ffwm_b200 import _lib
    lib = _lib.lib()
    null = ctypes.c_void_p(0)
    assert _lib.get_option("DISABLE_TILED") == 0 and _lib.get_option("NO_SUCH") == -1
    assert _lib.set_option("FORCE_TILED", 1) == 0 and _lib.get_option("FFWM_FORCE_TILED") == 1 and _lib.set_option("FORCE_TILED", 0) == 1
    with pytest.raises(ValueError):
        _lib.set_option("NO_SUCH", 1)
    TF, BF = _lib.MATH_TF32X3, _lib.MATH_BF16X3
    assert lib.ffwm_conv3x3_packed_floats(195, 195, 64, TF) == 4 * 25 * 9216           # 3xTF32: K blocks of 8 channels
    assert lib.ffwm_conv3x3_packed_floats(195, 195, 128, TF) == 2 * 25 * 18432
    assert lib.ffwm_conv3x3_packed_floats(195, 195, 64, BF) == 4 * 13 * 9216           # 3xBF16: K blocks of 16, same bytes each
    assert lib.ffwm_conv3x3_packed_floats(195, 195, 96, BF) == 0 and lib.ffwm_conv3x3_packed_floats(0, 8, 64, BF) == 0
    assert lib.ffwm_conv3x3_packed_floats(195, 195, 64, 2) == 0
    x64, o64 = torch.zeros(1, 8, 4, 64), torch.zeros(1, 128, 4, 64)
    fake = ctypes.c_void_p(16)                                      # non-null "packed" pointer, never dereferenced
    rc = lib.ffwm_conv3x3_forward(_lib.t4(x64), fake, null, _lib.t4(o64), 128, BF, null)
    assert rc == -3 and b"nt must be 64" in lib.ffwm_last_error()  # the 128-channel tile exists for W = 128 only
    rc = lib.ffwm_conv3x3_forward(_lib.t4(x64), fake, null, _lib.t4(o64), 32, BF, null)
    assert rc == -3
    rc = lib.ffwm_conv3x3_forward(_lib.t4(torch.zeros(1, 8, 4, 48)), fake, null, _lib.t4(torch.zeros(1, 8, 4, 48)), 64, TF, null)
    assert rc == -2                                                 # width not in {128, 64, 32, 16}
    w = torch.zeros(16, 8, 3, 3)
    rc = lib.ffwm_conv3x3_pack_weights(_lib.t4(w), 0, fake, ctypes.c_int64(10), 64, BF, null)
    assert rc == -2 and b"too small" in lib.ffwm_last_error()
    rc = lib.ffwm_conv3x3_pack_weights(_lib.t4(w), 0, fake, ctypes.c_int64(1 << 20), 100, BF, null)
    assert rc == -3
    # general convolution family: geometry is validated before any launch
    assert lib.ffwm_conv_packed_bytes(195, 195, 3, 3, BF) == 13 * 9 * 208 * 64 and lib.ffwm_conv_packed_bytes(195, 195, 3, 3, TF) == 25 * 9 * 208 * 64
    assert lib.ffwm_conv_packed_bytes(8, 8, 8, 8, BF) == 0                                   # more than 49 taps
    x, o = torch.zeros(1, 8, 8, 8), torch.zeros(1, 16, 5, 4)
    rc = lib.ffwm_conv_forward(_lib.t4(x), fake, null, _lib.t4(o), 3, 3, 2, 1, 0, BF, null)
    assert rc == -2 and b"does not match" in lib.ffwm_last_error()
    rc = lib.ffwm_conv_forward(_lib.t4(x), fake, null, _lib.t4(torch.zeros(1, 16, 4, 4)), 3, 3, 3, 1, 0, BF, null)
    assert rc == -3                                                                          # stride 3
    assert lib.ffwm_conv_wgrad_workspace_bytes(1, 16, 8, 4, 4, 8, 8, 3, 3, 2, 1) > 0
    assert lib.ffwm_conv_wgrad_workspace_bytes(1, 16, 8, 4, 4, 9, 9, 3, 3, 1, 1) == 0        # sides do not match
    rc = lib.ffwm_conv_wgrad(_lib.t4(torch.zeros(1, 16, 4, 4)), _lib.t4(x), _lib.t4(torch.zeros(16, 9, 3, 3)), 2, 1, fake, ctypes.c_int64(1 << 30), null)
    assert rc == -2                                                                          # grad_weight channels
    assert lib.ffwm_corr_max_workspace_bytes(2, 64, 1024) > 0 and lib.ffwm_corr_max_workspace_bytes(2, 96, 1024) == 0
    rc = lib.ffwm_affine_reg_forward(_lib.t4(torch.zeros(1, 1, 8, 8)), fake, _lib.t4(torch.zeros(1, 1, 5, 5)), 4, 0, null)
    assert rc == -3
    rc = lib.ffwm_ingest_u8(fake, null, fake, 1, 4, 4, 2, null)
    assert rc == -2
    # max-feature-map: sizes and pointers
    assert lib.ffwm_mfm_forward(null, null, ctypes.c_int64(0), ctypes.c_int64(5), null) == 0
    assert lib.ffwm_mfm_forward(null, null, ctypes.c_int64(2), ctypes.c_int64(5), null) == -1
    assert lib.ffwm_mfm_backward(null, null, null, ctypes.c_int64(-1), ctypes.c_int64(5), null) == -2
    # guided filter: the reference's H, W > 2r+1 assert, null pointers, empty input
    f = ctypes.c_float(1e-8)
    assert lib.ffwm_guided_filter_forward(fake, fake, fake, fake, fake, ctypes.c_int64(3), 16, 16, 8, f, null) == -3
    assert lib.ffwm_guided_filter_forward(null, fake, fake, fake, fake, ctypes.c_int64(3), 32, 32, 8, f, null) == -1
    assert lib.ffwm_guided_filter_forward(null, null, null, null, null, ctypes.c_int64(0), 32, 32, 8, f, null) == 0
    assert lib.ffwm_guided_filter_backward(fake, fake, fake, fake, fake, null, ctypes.c_int64(3), 32, 32, 8, null) == -1


def test_python_surface_mirrors_reference_errors():
    from ffwm_b200 import external_function as E
    with pytest.raises(NotImplementedError):                       # external_function.py:37-38
        E.BlockExtractor(3)(torch.rand(1, 1, 8, 8), torch.zeros(1, 2, 6, 6))
    with pytest.raises(NotImplementedError):                       # :84-85
        E.LocalAttnReshape()(torch.rand(1, 9, 4, 4), 3)
    with pytest.raises(AssertionError):                            # :77  ds == k*k
        E.LocalAttnReshape()(torch.rand(1, 8, 4, 4), 3)
    with pytest.raises(AssertionError):                            # :31  df == 2
        E.BlockExtractor(3)(torch.rand(1, 1, 8, 8), torch.zeros(1, 3, 6, 6))
    with pytest.raises(TypeError):
        from ffwm_b200 import _lib
        _lib.dtype_code(torch.zeros(1, dtype=torch.float16))


def test_dropin_modules_expose_reference_surface():
    import sys
    import ffwm_b200
    ffwm_b200.install_dropin()
    for n in ("resample2d_cuda", "block_extractor_cuda", "local_attn_reshape_cuda"):
        m = sys.modules[n]
        assert callable(m.forward) and callable(m.backward)
    import inspect
    import resample2d_cuda
    assert list(inspect.signature(resample2d_cuda.backward).parameters) == [
        "input1", "input2", "gradOutput", "gradInput1", "gradInput2", "kernel_size", "dilation"]


def test_guided_filter_box_sums():
    from ffwm_b200.external_function import BoxFilter
    x = torch.rand(2, 3, 40, 37, dtype=torch.float64)
    r = 5
    ref = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(x, (r, r, r, r)), 2 * r + 1, 1, divisor_override=1)
    assert torch.allclose(BoxFilter(r)(x), ref, atol=1e-9)


def test_packed_weight_cache_invalidation(monkeypatch):
    """conv._packed (FFWM_CACHE_PACKED, default on): frozen leaves are packed once per (direction, tile) and re-packed
    after any in-place update; trainable or derived weights are never cached.  (Packing itself is CUDA-only: mocked.)"""
    from ffwm_b200 import conv
    calls = []
    monkeypatch.setattr(conv.ops, "conv3x3_pack_weights", lambda w, dgrad=False, nt=64, math=1: calls.append((dgrad, nt)) or object())
    monkeypatch.setattr(conv, "CACHE_PACKED", True)
    w = torch.nn.Parameter(torch.zeros(8, 4, 3, 3), requires_grad=False)
    a = conv._packed(w, False, 64, 1)
    assert conv._packed(w, False, 64, 1) is a and len(calls) == 1
    assert conv._packed(w, True, 64, 1) is not a and conv._packed(w, False, 128, 1) is not a and len(calls) == 3
    assert conv._packed(w, False, 64, 0) is not a and len(calls) == 4         # another operand math: another image
    with torch.no_grad():
        w.add_(1.0)                                       # optimizer step / load_state_dict: version counter moves
    assert conv._packed(w, False, 64, 1) is not a and len(calls) == 5
    w.data = torch.ones(8, 4, 3, 3)                       # storage swapped
    conv._packed(w, False, 64, 1)
    assert len(calls) == 6
    t = torch.nn.Parameter(torch.zeros(8, 4, 3, 3))       # trainable: never cached
    conv._packed(t, False, 64, 1), conv._packed(t, False, 64, 1)
    conv._packed(w * 2, False, 64, 1), conv._packed(w * 2, False, 64, 1)   # derived (spectral norm): a new tensor per call
    assert len(calls) == 10
    monkeypatch.setattr(conv, "CACHE_PACKED", False)
    conv._packed(w, False, 64, 1), conv._packed(w, False, 64, 1)
    assert len(calls) == 12
