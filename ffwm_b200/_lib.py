"""ctypes binding of libffwm_b200.so (include/ffwm_b200.h).

PyTorch is only plumbing here: it owns device memory and streams; every byte
of arithmetic happens in the hand-written sm_100a kernels behind the C ABI.
There is no CPU fallback: a missing library or a non-CUDA tensor raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libffwm_b200.so")

FFWM_F32, FFWM_F64 = 0, 1
MATH_TF32X3, MATH_BF16X3 = 0, 1          # enum FFWM_MATH_* (operand split of the tcgen05 convolutions)
ABI_VERSION = 2


class Tensor4(ctypes.Structure):
    """struct ffwm_tensor4"""
    _fields_ = [("data", ctypes.c_void_p),
                ("size", ctypes.c_int64 * 4),
                ("stride", ctypes.c_int64 * 4)]


_T4P = ctypes.POINTER(Tensor4)
_I, _VP = ctypes.c_int, ctypes.c_void_p

# name -> argtypes; mirrors include/ffwm_b200.h one to one (tests check it)
SIGNATURES = {
    "ffwm_abi_version": [],
    "ffwm_last_error": [],
    "ffwm_kernel_launches": [],
    "ffwm_set_option": [ctypes.c_char_p, _I],
    "ffwm_get_option": [ctypes.c_char_p],
    "ffwm_resample2d_forward": [_T4P, _T4P, _T4P, _I, _I, _I, _VP],
    "ffwm_resample2d_backward": [_T4P, _T4P, _T4P, _T4P, _T4P, _I, _I, _I, _VP],
    "ffwm_block_extractor_forward": [_T4P, _T4P, _T4P, _I, _I, _VP],
    "ffwm_block_extractor_backward": [_T4P, _T4P, _T4P, _T4P, _T4P, _I, _I, _VP],
    "ffwm_local_attn_reshape_forward": [_T4P, _T4P, _I, _I, _VP],
    "ffwm_local_attn_reshape_backward": [_T4P, _T4P, _I, _I, _VP],
    "ffwm_grid_warp_forward": [_T4P, _T4P, _T4P, _I, _VP],
    "ffwm_grid_warp_backward": [_T4P, _T4P, _T4P, _T4P, _T4P, _I, _VP],
    "ffwm_conv3x3_packed_floats": [_I, _I, _I, _I],
    "ffwm_conv3x3_pack_weights": [_T4P, _I, _VP, ctypes.c_int64, _I, _I, _VP],
    "ffwm_conv3x3_forward": [_T4P, _VP, _VP, _T4P, _I, _I, _VP],
    "ffwm_conv_packed_bytes": [_I, _I, _I, _I, _I],
    "ffwm_conv_pack_weights": [_T4P, _I, _I, _I, _I, _I, _VP, ctypes.c_int64, _VP],
    "ffwm_conv_forward": [_T4P, _VP, _VP, _T4P, _I, _I, _I, _I, _I, _I, _VP],
    "ffwm_conv_wgrad_workspace_bytes": [_I] * 11,
    "ffwm_conv_wgrad": [_T4P, _T4P, _T4P, _I, _I, _VP, ctypes.c_int64, _VP],
    "ffwm_affine_reg_forward": [_T4P, _VP, _T4P, _I, _I, _VP],
    "ffwm_affine_reg_backward": [_T4P, _VP, _T4P, _T4P, _I, _I, _VP],
    "ffwm_batch_norm_workspace_bytes": [_I, _I, ctypes.c_int64],
    "ffwm_batch_norm_forward": [_VP] * 6 + [ctypes.c_float] * 3 + [_VP] * 3 + [_I, _I, ctypes.c_int64, _VP, ctypes.c_int64, _VP],
    "ffwm_batch_norm_backward": [_VP] * 7 + [ctypes.c_float] + [_VP] * 4 + [_I, _I, ctypes.c_int64, _VP, ctypes.c_int64, _VP],
    "ffwm_conv_few": [_T4P, _T4P, _I, _I, _VP, _T4P, _I, _I, _VP],
    "ffwm_max_pool2x2_forward": [_VP, _VP, ctypes.c_int64, _I, _I, _I, _I, _VP],
    "ffwm_max_pool2x2_backward": [_VP, _VP, _VP, ctypes.c_int64, _I, _I, _I, _I, _VP],
    "ffwm_spectral_norm_forward": [_VP, _I, _I, ctypes.c_float] + [_VP] * 6 + [_I, _I, _I, _VP],
    "ffwm_spectral_norm_backward": [_VP, _I] + [_VP] * 6 + [_I, _VP],
    "ffwm_channel_sum": [_VP, _VP, _I, _I, ctypes.c_int64, _VP, ctypes.c_int64, _VP],
    "ffwm_corr_max_workspace_bytes": [_I, _I, _I],
    "ffwm_corr_max": [_T4P, _T4P, ctypes.c_float, _VP, _VP, ctypes.c_int64, _VP],
    "ffwm_ingest_u8": [_VP, _VP, _VP, _I, _I, _I, _I, _VP],
    "ffwm_mfm_forward": [_VP, _VP, ctypes.c_int64, ctypes.c_int64, _VP],
    "ffwm_mfm_backward": [_VP, _VP, _VP, ctypes.c_int64, ctypes.c_int64, _VP],
    "ffwm_guided_filter_forward": [_VP, _VP, _VP, _VP, _VP, ctypes.c_int64, _I, _I, _I, ctypes.c_float, _VP],
    "ffwm_guided_filter_backward": [_VP, _VP, _VP, _VP, _VP, _VP, ctypes.c_int64, _I, _I, _I, _VP],
}

_lib = None
LAUNCHES = 0   # C-ABI entry-point calls made by this process (kernel launches: kernel_launches())


def kernel_launches():
    """Kernels enqueued by libffwm_b200 in this process (an entry point may launch more than one)."""
    return int(lib().ffwm_kernel_launches())


def lib():
    """Load the product library once; fail loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "ffwm_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C ffwm_b200/csrc`. There is no CPU or PyTorch fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = {"ffwm_last_error": ctypes.c_char_p, "ffwm_kernel_launches": ctypes.c_ulonglong,
                          "ffwm_conv3x3_packed_floats": ctypes.c_int64,
                          "ffwm_conv_packed_bytes": ctypes.c_int64,
                          "ffwm_conv_wgrad_workspace_bytes": ctypes.c_int64,
                          "ffwm_batch_norm_workspace_bytes": ctypes.c_int64,
                          "ffwm_corr_max_workspace_bytes": ctypes.c_int64}.get(name, ctypes.c_int)
        if l.ffwm_abi_version() != ABI_VERSION:
            raise ImportError("ffwm_b200: ABI version mismatch (library %d, binding %d)"
                              % (l.ffwm_abi_version(), ABI_VERSION))
        _lib = l
    return _lib


def set_option(name, value):
    """Process-wide runtime option of the library (include/ffwm_b200.h: ffwm_set_option); returns the old value."""
    l = lib()
    old = l.ffwm_get_option(name.encode())
    if l.ffwm_set_option(name.encode(), int(value)) != 0:
        raise ValueError(l.ffwm_last_error().decode())
    return old


def get_option(name):
    return int(lib().ffwm_get_option(name.encode()))


def dtype_code(t):
    if t.dtype == torch.float32:
        return FFWM_F32
    if t.dtype == torch.float64:
        return FFWM_F64
    raise TypeError("ffwm_b200 kernels take float32/float64 (as the reference's "
                    "AT_DISPATCH_FLOATING_TYPES), got %s" % t.dtype)


def t4(t):
    """Describe a 4-D CUDA tensor (or None -> null data) for the C ABI."""
    d = Tensor4()
    if t is None:
        d.data = None
        return d
    if t.dim() != 4:
        raise ValueError("expected a 4-D tensor, got %d-D" % t.dim())
    d.data = t.data_ptr()
    d.size[:] = t.shape
    d.stride[:] = t.stride()
    return d


def require_cuda(*tensors):
    """The reference raises NotImplementedError for CPU tensors
    (models/external_function.py:37-38,84-85); so does this path.  Also checks that all tensors of the call live on
    one device and share one dtype."""
    dev = dtype = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NotImplementedError("ffwm_b200 ops are CUDA-only (no CPU path, as in the reference)")
        if dev is None:
            dev, dtype = t.device, t.dtype
        elif t.device != dev:
            raise RuntimeError("ffwm_b200: tensors live on different devices (%s vs %s)" % (dev, t.device))
        elif t.dtype != dtype:
            # the C ABI takes ONE dtype code per call and reinterprets every pointer with it; the reference's
            # `.data<scalar_t>()` raises on a mismatch, so does this (a silent mismatch would read out of bounds)
            raise TypeError("ffwm_b200: tensors of one call must share a dtype (%s vs %s)" % (dtype, t.dtype))
    return dev


def call(name, dev, *args):
    """Invoke one entry point on `dev`'s current stream and raise on failure."""
    global LAUNCHES
    l = lib()
    LAUNCHES += 1
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        rc = getattr(l, name)(*args, ctypes.c_void_p(stream))
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, l.ffwm_last_error().decode()))
