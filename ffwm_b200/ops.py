"""Thin tensor-level wrappers over the C ABI (one function per entry point).

These take and fill caller-allocated CUDA tensors exactly like the reference's
pybind functions do; autograd lives in external_function.py.
"""
from . import _lib as L


def resample2d_forward(input1, input2, output, kernel_size, dilation):
    dev = L.require_cuda(input1, input2, output)
    L.call("ffwm_resample2d_forward", dev, L.t4(input1), L.t4(input2), L.t4(output),
           int(kernel_size), int(dilation), L.dtype_code(input1))


def resample2d_backward(input1, input2, grad_output, grad_input1, grad_input2, kernel_size, dilation):
    dev = L.require_cuda(input1, input2, grad_output, grad_input1, grad_input2)
    L.call("ffwm_resample2d_backward", dev, L.t4(input1), L.t4(input2), L.t4(grad_output),
           L.t4(grad_input1), L.t4(grad_input2), int(kernel_size), int(dilation), L.dtype_code(input1))


def block_extractor_forward(source, flow_field, output, kernel_size):
    dev = L.require_cuda(source, flow_field, output)
    L.call("ffwm_block_extractor_forward", dev, L.t4(source), L.t4(flow_field), L.t4(output),
           int(kernel_size), L.dtype_code(source))


def block_extractor_backward(source, flow_field, grad_output, grad_source, grad_flow_field, kernel_size):
    dev = L.require_cuda(source, flow_field, grad_output, grad_source, grad_flow_field)
    L.call("ffwm_block_extractor_backward", dev, L.t4(source), L.t4(flow_field), L.t4(grad_output),
           L.t4(grad_source), L.t4(grad_flow_field), int(kernel_size), L.dtype_code(source))


def local_attn_reshape_forward(inputs, output, kernel_size):
    dev = L.require_cuda(inputs, output)
    L.call("ffwm_local_attn_reshape_forward", dev, L.t4(inputs), L.t4(output),
           int(kernel_size), L.dtype_code(inputs))


def local_attn_reshape_backward(grad_output, grad_inputs, kernel_size):
    dev = L.require_cuda(grad_output, grad_inputs)
    L.call("ffwm_local_attn_reshape_backward", dev, L.t4(grad_output), L.t4(grad_inputs),
           int(kernel_size), L.dtype_code(grad_inputs))


def grid_warp_forward(images, flow, output):
    dev = L.require_cuda(images, flow, output)
    L.call("ffwm_grid_warp_forward", dev, L.t4(images), L.t4(flow), L.t4(output), L.dtype_code(images))


def grid_warp_backward(images, flow, grad_output, grad_images, grad_flow):
    dev = L.require_cuda(images, flow, grad_output, grad_images, grad_flow)
    L.call("ffwm_grid_warp_backward", dev, L.t4(images), L.t4(flow), L.t4(grad_output),
           L.t4(grad_images), L.t4(grad_flow), L.dtype_code(images))


def conv3x3_pack_weights(weight, dgrad=False, nt=64):
    """(Cout,Cin,3,3) fp32 CUDA weight -> packed image for conv3x3_forward (dgrad: for the data gradient).
    nt: output channels per CTA the image is laid out for (64; 128 is experimental, W = 128 only)."""
    import ctypes
    import torch
    dev = L.require_cuda(weight)
    cout, cin = weight.shape[:2]
    n = L.lib().ffwm_conv3x3_packed_floats_nt(int(cin if dgrad else cout), int(cout if dgrad else cin), int(nt))
    if n <= 0:
        raise ValueError("conv3x3_pack_weights: bad shape or nt (cout %d, cin %d, nt %d)" % (cout, cin, nt))
    packed = torch.empty(n, dtype=torch.float32, device=weight.device)
    L.call("ffwm_conv3x3_pack_weights_nt", dev, L.t4(weight), int(bool(dgrad)), ctypes.c_void_p(packed.data_ptr()),
           ctypes.c_int64(n), int(nt))
    return packed


def conv3x3_forward(x, packed, bias, out, nt=64):
    import ctypes
    dev = L.require_cuda(x, packed, out)
    L.call("ffwm_conv3x3_forward_nt", dev, L.t4(x), ctypes.c_void_p(packed.data_ptr()),
           ctypes.c_void_p(bias.data_ptr() if bias is not None else None), L.t4(out), int(nt))


def conv3x3_wgrad(x, grad_out, grad_weight, grad_bias=None):
    """grad_weight (Cout,Cin,3,3, zero-filled by the caller) += weight gradient of the 3x3/s1/p1 convolution;
    grad_bias (Cout, contiguous fp32, zero-filled) += grad_out.sum((0,2,3)) if given.
    EXPERIMENTAL (not yet run on a B200): see csrc/conv3x3_wgrad_tc.cu."""
    import ctypes
    dev = L.require_cuda(x, grad_out, grad_weight, grad_bias)
    if grad_bias is not None and not (grad_bias.is_contiguous() and grad_bias.numel() == grad_out.size(1)
                                      and grad_bias.dtype == grad_out.dtype):
        raise ValueError("conv3x3_wgrad: grad_bias must be a contiguous fp32 tensor of Cout elements")
    L.call("ffwm_conv3x3_wgrad", dev, L.t4(x), L.t4(grad_out), L.t4(grad_weight),
           ctypes.c_void_p(grad_bias.data_ptr() if grad_bias is not None else None))


def mfm_forward(x, out):
    """out (N,C,...) = elementwise max of the two channel halves of x (N,2C,...); contiguous fp32 CUDA tensors.
    EXPERIMENTAL (not yet run on a B200): see csrc/mfm.cu."""
    import ctypes
    dev = L.require_cuda(x, out)
    n, chw = x.size(0), out.numel() // max(x.size(0), 1)
    if not (x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype and L.dtype_code(x) == L.FFWM_F32
            and x.numel() == 2 * out.numel()):
        raise ValueError("mfm_forward: contiguous fp32 x (N,2C,...) and out (N,C,...) expected")
    L.call("ffwm_mfm_forward", dev, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()),
           ctypes.c_int64(n), ctypes.c_int64(chw))


def mfm_backward(x, grad_out, grad_x):
    import ctypes
    dev = L.require_cuda(x, grad_out, grad_x)
    n, chw = x.size(0), grad_out.numel() // max(x.size(0), 1)
    if not (x.is_contiguous() and grad_out.is_contiguous() and grad_x.is_contiguous() and L.dtype_code(x) == L.FFWM_F32
            and grad_out.dtype == x.dtype and grad_x.dtype == x.dtype and x.numel() == 2 * grad_out.numel()
            and grad_x.numel() == x.numel()):
        raise ValueError("mfm_backward: contiguous fp32 x / grad_x (N,2C,...) and grad_out (N,C,...) expected")
    L.call("ffwm_mfm_backward", dev, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(grad_out.data_ptr()),
           ctypes.c_void_p(grad_x.data_ptr()), ctypes.c_int64(n), ctypes.c_int64(chw))


def _gf_ptrs(*tensors):
    import ctypes
    for t in tensors:
        if not (t.is_contiguous() and L.dtype_code(t) == L.FFWM_F32):
            raise ValueError("guided_filter: contiguous fp32 CUDA tensors expected")
    return [ctypes.c_void_p(t.data_ptr()) for t in tensors]


def guided_filter_forward(x, y, q, save, scratch, r, eps):
    """q = GuidedFilter(r, eps)(x, y) for (B,C,H,W) x, y; save (5,B,C,H,W), scratch (5,B,C,H,W).
    EXPERIMENTAL (not yet run on a B200): see csrc/guided_filter.cu."""
    import ctypes
    dev = L.require_cuda(x, y, q, save, scratch)
    b, c, h, w = x.shape
    if y.shape != x.shape or q.shape != x.shape or save.numel() < 5 * x.numel() or scratch.numel() < 5 * x.numel():
        raise ValueError("guided_filter_forward: shape mismatch")
    L.call("ffwm_guided_filter_forward", dev, *_gf_ptrs(x, y, q, save, scratch), ctypes.c_int64(b * c), int(h), int(w),
           int(r), ctypes.c_float(eps))


def guided_filter_backward(x, y, grad_q, save, grad_x, scratch, r):
    import ctypes
    dev = L.require_cuda(x, y, grad_q, save, grad_x, scratch)
    b, c, h, w = x.shape
    if (y.shape != x.shape or grad_q.shape != x.shape or grad_x.shape != x.shape or save.numel() < 5 * x.numel()
            or scratch.numel() < 6 * x.numel()):
        raise ValueError("guided_filter_backward: shape mismatch")
    L.call("ffwm_guided_filter_backward", dev, *_gf_ptrs(x, y, grad_q, save, grad_x, scratch), ctypes.c_int64(b * c),
           int(h), int(w), int(r))
