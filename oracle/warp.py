"""ORACLE (test infrastructure): ctypes front-end of oracle/liboracle.so.

Each function takes CPU torch tensors (float32 or float64, any strides) and
mirrors one reference entry point; the C side cites the reference file:line.
Output allocation follows models/external_function.py (zero-filled, caller
owned): :35 (block extractor), :82 (local attn reshape), :124 (resample2d).
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("warp_ops.c", "warp_ops.inc", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def _suffix(t):
    if t.dtype == torch.float32:
        return "_f32"
    if t.dtype == torch.float64:
        return "_f64"
    raise TypeError("oracle supports float32/float64 only (AT_DISPATCH_FLOATING_TYPES), got %s" % t.dtype)


def _i64x4(vals):
    return (ctypes.c_int64 * 4)(*[int(v) for v in vals])


def _t(t):
    """(data pointer, sizes[4], strides[4]) of a 4-D CPU tensor."""
    assert t.dim() == 4 and t.device.type == "cpu"
    return ctypes.c_void_p(t.data_ptr()), _i64x4(t.shape), _i64x4(t.stride())


def _call(name, like, *args):
    fn = getattr(lib(), name + _suffix(like))
    fn.restype = None
    flat = []
    for a in args:
        if isinstance(a, torch.Tensor):
            flat.extend(_t(a))
        elif isinstance(a, tuple):       # (tensor, 'nosize') -> ptr, strides only
            p, _, st = _t(a[0])
            flat.extend([p, st])
        else:
            flat.append(ctypes.c_int(int(a)))
    fn(*flat)


def resample2d_forward(input1, input2, kernel_size=2, dilation=1):
    b, _, h, w = input2.shape
    out = input1.new_zeros(b, input1.shape[1], h, w)
    _call("oracle_resample2d_fwd", input1, input1, input2, out, kernel_size, dilation)
    return out


def resample2d_backward(input1, input2, grad_output, kernel_size=2, dilation=1):
    g1 = torch.zeros_like(input1, memory_format=torch.contiguous_format)
    g2 = torch.zeros_like(input2, memory_format=torch.contiguous_format)
    _call("oracle_resample2d_bwd_input1", input1, input1, input2, grad_output, g1, kernel_size, dilation)
    _call("oracle_resample2d_bwd_input2", input1, input1, input2, grad_output, g2, kernel_size, dilation)
    return g1, g2


def block_extractor_forward(source, flow, kernel_size):
    bs, ds, _, _ = source.shape
    _, df, hf, wf = flow.shape
    assert df == 2
    out = flow.new_zeros(bs, ds, kernel_size * hf, kernel_size * wf)
    _call("oracle_block_extractor_fwd", source, source, flow, out, kernel_size)
    return out


def block_extractor_backward(source, flow, grad_output, kernel_size):
    gs = torch.zeros_like(source, memory_format=torch.contiguous_format)
    gf = torch.zeros_like(flow, memory_format=torch.contiguous_format)
    _call("oracle_block_extractor_bwd", source, source, flow, grad_output, gs, gf, kernel_size)
    return gs, gf


def local_attn_reshape_forward(inputs, kernel_size):
    bs, ds, hs, ws = inputs.shape
    assert ds == kernel_size * kernel_size
    out = inputs.new_zeros(bs, 1, kernel_size * hs, kernel_size * ws)
    _call("oracle_local_attn_reshape_fwd", inputs, inputs, out, kernel_size)
    return out


def local_attn_reshape_backward(inputs, grad_output, kernel_size):
    gi = torch.zeros_like(inputs, memory_format=torch.contiguous_format)
    _call("oracle_local_attn_reshape_bwd", inputs, grad_output, gi, kernel_size)
    return gi


def grid_warp_forward(images, flow):
    b, _, h, w = flow.shape
    out = images.new_zeros(b, images.shape[1], h, w)
    _call("oracle_grid_warp_fwd", images, images, flow, out)
    return out


def grid_warp_backward(images, flow, grad_output):
    gi = torch.zeros_like(images, memory_format=torch.contiguous_format)
    gf = torch.zeros_like(flow, memory_format=torch.contiguous_format)
    _call("oracle_grid_warp_bwd", images, images, flow, grad_output, (gi, "nosize"), (gf, "nosize"))
    return gi, gf
