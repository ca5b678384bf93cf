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


def conv3x3_pack_weights(weight, dgrad=False, nt=64, math=L.MATH_BF16X3):
    """(Cout,Cin,3,3) fp32 CUDA weight -> packed image for conv3x3_forward (dgrad: for the data gradient).
    nt: output channels per CTA the image is laid out for (64, or 128 for W = 128); math: L.MATH_TF32X3 / L.MATH_BF16X3."""
    import ctypes
    import torch
    dev = L.require_cuda(weight)
    cout, cin = weight.shape[:2]
    n = L.lib().ffwm_conv3x3_packed_floats(int(cin if dgrad else cout), int(cout if dgrad else cin), int(nt), int(math))
    if n <= 0:
        raise ValueError("conv3x3_pack_weights: bad shape, nt or math (cout %d, cin %d, nt %d, math %d)" % (cout, cin, nt, math))
    packed = torch.empty(n, dtype=torch.float32, device=weight.device)
    L.call("ffwm_conv3x3_pack_weights", dev, L.t4(weight), int(bool(dgrad)), ctypes.c_void_p(packed.data_ptr()),
           ctypes.c_int64(n), int(nt), int(math))
    return packed


def conv3x3_forward(x, packed, bias, out, nt=64, math=L.MATH_BF16X3):
    import ctypes
    dev = L.require_cuda(x, packed, out)
    L.call("ffwm_conv3x3_forward", dev, L.t4(x), ctypes.c_void_p(packed.data_ptr()),
           ctypes.c_void_p(bias.data_ptr() if bias is not None else None), L.t4(out), int(nt), int(math))


def conv_pack_weights(weight, in_major, stride, pad, transposed, math=L.MATH_BF16X3):
    """Packed image of a (d0, d1, kh, kw) fp32 CUDA weight for conv_forward (csrc/conv_gen_tc.cu).  in_major: dim 0 is
    the INPUT channel of the operation (ConvTranspose2d forward, Conv2d data gradient)."""
    import ctypes
    import torch
    dev = L.require_cuda(weight)
    n_out, n_in = (weight.size(1), weight.size(0)) if in_major else (weight.size(0), weight.size(1))
    n = L.lib().ffwm_conv_packed_bytes(int(n_out), int(n_in), int(weight.size(2)), int(weight.size(3)), int(math))
    if n <= 0:
        raise ValueError("conv_pack_weights: unsupported weight shape %s" % (tuple(weight.shape),))
    packed = torch.empty(n, dtype=torch.uint8, device=weight.device)
    L.call("ffwm_conv_pack_weights", dev, L.t4(weight), int(bool(in_major)), int(stride), int(pad), int(bool(transposed)), int(math),
           ctypes.c_void_p(packed.data_ptr()), ctypes.c_int64(n))
    return packed


def conv_forward(x, packed, bias, out, kh, kw, stride, pad, transposed, math=L.MATH_BF16X3):
    """out = conv2d / conv_transpose2d(x, W, bias, stride, pad) on tcgen05, W packed by conv_pack_weights."""
    import ctypes
    dev = L.require_cuda(x, out)
    L.require_cuda(packed)
    if bias is not None and not (bias.is_cuda and bias.is_contiguous() and bias.dtype == x.dtype and bias.numel() == out.size(1)):
        raise ValueError("conv_forward: bias must be a contiguous fp32 CUDA tensor of Cout elements")
    L.call("ffwm_conv_forward", dev, L.t4(x), ctypes.c_void_p(packed.data_ptr()),
           ctypes.c_void_p(bias.data_ptr() if bias is not None else None), L.t4(out), int(kh), int(kw), int(stride), int(pad),
           int(bool(transposed)), int(math))


def conv_wgrad(small, large, grad_weight, stride, pad):
    """grad_weight (Ca,Cb,kh,kw; overwritten) = weight gradient of a convolution (csrc/conv_gen_wgrad_tc.cu).
    nn.Conv2d: small = grad_out, large = input; nn.ConvTranspose2d: small = input, large = grad_out."""
    import ctypes
    import torch
    dev = L.require_cuda(small, large, grad_weight)
    n = L.lib().ffwm_conv_wgrad_workspace_bytes(int(small.size(0)), int(small.size(1)), int(large.size(1)), int(small.size(2)),
                                                int(small.size(3)), int(large.size(2)), int(large.size(3)),
                                                int(grad_weight.size(2)), int(grad_weight.size(3)), int(stride), int(pad))
    if n <= 0:
        raise RuntimeError("conv_wgrad: unsupported shapes: %s" % L.lib().ffwm_last_error().decode())
    ws = torch.empty(n, dtype=torch.uint8, device=small.device)
    L.call("ffwm_conv_wgrad", dev, L.t4(small), L.t4(large), L.t4(grad_weight), int(stride), int(pad),
           ctypes.c_void_p(ws.data_ptr()), ctypes.c_int64(n))


def affine_reg_forward(grid, q, out, kz):
    """out (B,1,H-kz+1,W-kz+1) = w^T Q w / kz^2 per kz x kz window of grid (B,1,H,W); q: (kz^2, kz^2) contiguous, same dtype."""
    import ctypes
    dev = L.require_cuda(grid, q, out)
    if not (q.is_contiguous() and q.numel() == kz ** 4):
        raise ValueError("affine_reg: Q must be a contiguous (kz^2, kz^2) matrix")
    L.call("ffwm_affine_reg_forward", dev, L.t4(grid), ctypes.c_void_p(q.data_ptr()), L.t4(out), int(kz), L.dtype_code(grid))


def affine_reg_backward(grid, q, grad_out, grad_grid, kz):
    import ctypes
    dev = L.require_cuda(grid, q, grad_out, grad_grid)
    if not (q.is_contiguous() and q.numel() == kz ** 4):
        raise ValueError("affine_reg: Q must be a contiguous (kz^2, kz^2) matrix")
    L.call("ffwm_affine_reg_backward", dev, L.t4(grid), ctypes.c_void_p(q.data_ptr()), L.t4(grad_out), L.t4(grad_grid), int(kz),
           L.dtype_code(grid))


def corr_max_supported(x):
    return x.is_cuda and x.dtype == __import__("torch").float32 and x.dim() == 4 and x.size(1) % 64 == 0 and x.size(1) <= 256


def corr_max(source, target, eps):
    """(B, H*W) column max of the cosine correlation between every source and target pixel (csrc/corr_max.cu)."""
    import ctypes
    import torch
    dev = L.require_cuda(source, target)
    source, target = source.contiguous(), target.contiguous()
    b, c, h, w = source.shape
    n = L.lib().ffwm_corr_max_workspace_bytes(int(b), int(c), int(h * w))
    if n <= 0:
        raise ValueError("corr_max: unsupported shape %s (C must be a multiple of 64 up to 256)" % (tuple(source.shape),))
    ws = torch.empty(n, dtype=torch.uint8, device=source.device)
    out = torch.empty((b, h * w), dtype=torch.float32, device=source.device)
    L.call("ffwm_corr_max", dev, L.t4(source), L.t4(target), ctypes.c_float(eps), ctypes.c_void_p(out.data_ptr()),
           ctypes.c_void_p(ws.data_ptr()), ctypes.c_int64(n))
    return out


def ingest_u8(src, flip, out):
    """out (B,C,H,W) fp32 = src (B,H,W,C) uint8 transposed, flipped left-right where flip[b] (uint8, or None), / 255."""
    import ctypes
    import torch
    dev = L.require_cuda(out)
    if not (src.is_cuda and src.dtype == torch.uint8 and src.is_contiguous() and out.is_contiguous() and out.dtype == torch.float32
            and src.dim() == 4 and out.shape == (src.size(0), src.size(3), src.size(1), src.size(2))):
        raise ValueError("ingest_u8: src (B,H,W,C) uint8 and out (B,C,H,W) float32, both dense CUDA tensors")
    if flip is not None and not (flip.is_cuda and flip.dtype == torch.uint8 and flip.is_contiguous() and flip.numel() == src.size(0)):
        raise ValueError("ingest_u8: flip must be B uint8 flags on the device")
    L.call("ffwm_ingest_u8", dev, ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(flip.data_ptr() if flip is not None else None),
           ctypes.c_void_p(out.data_ptr()), int(src.size(0)), int(src.size(1)), int(src.size(2)), int(src.size(3)))


def mfm_forward(x, out):
    """out (N,C,...) = elementwise max of the two channel halves of x (N,2C,...); contiguous fp32 CUDA tensors.
    See csrc/mfm.cu."""
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
    See csrc/guided_filter.cu."""
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


def _bn_ws(x):
    import torch
    n, c, hw = int(x.size(0)), int(x.size(1)), int(x.size(2) * x.size(3))
    nbytes = L.lib().ffwm_batch_norm_workspace_bytes(n, c, hw)
    if nbytes <= 0:
        raise ValueError("batch_norm: unsupported shape %s: %s" % (tuple(x.shape), L.lib().ffwm_last_error().decode()))
    return n, c, hw, torch.empty(nbytes // 8, dtype=torch.float64, device=x.device), nbytes


def _ptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


def _bn_check(what, maps, vectors, c):
    import torch
    for t in maps:
        if t is not None and not (t.is_contiguous() and t.dtype == torch.float32 and t.shape == maps[0].shape):
            raise ValueError("%s: contiguous fp32 (N,C,H,W) maps of one shape expected" % what)
    for v in vectors:
        if v is not None and not (v.is_cuda and v.is_contiguous() and v.dtype == torch.float32 and v.numel() == c):
            raise ValueError("%s: per-channel vectors must be contiguous fp32 CUDA tensors of C elements" % what)


def batch_norm_forward(x, residual, gamma, beta, running_mean, running_var, momentum, eps, act_slope, y, save_mean, save_invstd):
    """Training-mode BatchNorm2d (+ residual add, + LeakyReLU with `act_slope`; 1.0 = none) of a contiguous fp32 (N,C,H,W)
    map into y; updates running_mean / running_var (None to skip) like torch (csrc/batch_norm.cu)."""
    import ctypes
    dev = L.require_cuda(x, residual, y)
    n, c, hw, ws, nbytes = _bn_ws(x)
    _bn_check("batch_norm_forward", (x, residual, y), (gamma, beta, running_mean, running_var, save_mean, save_invstd), c)
    L.call("ffwm_batch_norm_forward", dev, _ptr(x), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
           ctypes.c_float(momentum), ctypes.c_float(eps), ctypes.c_float(act_slope), _ptr(y), _ptr(save_mean), _ptr(save_invstd),
           n, c, ctypes.c_int64(hw), _ptr(ws), ctypes.c_int64(nbytes))


def batch_norm_backward(x, grad_out, y_out, gamma, beta, save_mean, save_invstd, act_slope, grad_x, grad_residual, grad_gamma, grad_beta):
    """Gradient of batch_norm_forward: grad_x, grad_gamma, grad_beta (and grad_residual, with y_out = the forward's output,
    when a residual was added) are overwritten."""
    import ctypes
    dev = L.require_cuda(x, grad_out, y_out, grad_x, grad_residual)
    n, c, hw, ws, nbytes = _bn_ws(x)
    _bn_check("batch_norm_backward", (x, grad_out, y_out, grad_x, grad_residual), (gamma, beta, save_mean, save_invstd, grad_gamma, grad_beta), c)
    L.call("ffwm_batch_norm_backward", dev, _ptr(x), _ptr(grad_out), _ptr(y_out), _ptr(gamma), _ptr(beta), _ptr(save_mean), _ptr(save_invstd),
           ctypes.c_float(act_slope), _ptr(grad_x), _ptr(grad_residual), _ptr(grad_gamma), _ptr(grad_beta), n, c, ctypes.c_int64(hw),
           _ptr(ws), ctypes.c_int64(nbytes))


def channel_sum(x):
    """(C,) = x.sum((0, 2, 3)) of a contiguous fp32 CUDA (N,C,H,W) map (csrc/batch_norm.cu: the bias gradient of a convolution)."""
    import ctypes
    import torch
    dev = L.require_cuda(x)
    if not (x.dim() == 4 and x.is_contiguous() and x.dtype == torch.float32 and x.numel() > 0):
        raise ValueError("channel_sum: a non-empty contiguous fp32 (N,C,H,W) map expected")
    n, c, hw = int(x.size(0)), int(x.size(1)), int(x.size(2) * x.size(3))
    nbytes = L.lib().ffwm_batch_norm_workspace_bytes(max(n, 2), c, hw)
    ws = torch.empty(nbytes // 8, dtype=torch.float64, device=x.device)
    out = torch.empty(c, dtype=torch.float32, device=x.device)
    L.call("ffwm_channel_sum", dev, _ptr(x), _ptr(out), n, c, ctypes.c_int64(hw), _ptr(ws), ctypes.c_int64(nbytes))
    return out


def max_pool2x2_forward(x, out):
    """out (N,C,ho,wo) = 2x2 / stride-2 max pooling of a contiguous fp32 CUDA (N,C,h,w) map (csrc/pool.cu)."""
    import ctypes
    import torch
    dev = L.require_cuda(x, out)
    if not (x.dim() == 4 and x.is_contiguous() and out.is_contiguous() and x.dtype == torch.float32 and out.shape[:2] == x.shape[:2]):
        raise ValueError("max_pool2x2: contiguous fp32 (N,C,H,W) tensors expected")
    L.call("ffwm_max_pool2x2_forward", dev, _ptr(x), _ptr(out), ctypes.c_int64(x.size(0) * x.size(1)), int(x.size(2)), int(x.size(3)),
           int(out.size(2)), int(out.size(3)))


def max_pool2x2_backward(x, grad_out, grad_x):
    import ctypes
    import torch
    dev = L.require_cuda(x, grad_out, grad_x)
    if not (x.dim() == 4 and x.is_contiguous() and grad_out.is_contiguous() and grad_x.is_contiguous() and x.dtype == torch.float32
            and grad_x.shape == x.shape and grad_out.shape[:2] == x.shape[:2]):
        raise ValueError("max_pool2x2: contiguous fp32 (N,C,H,W) tensors expected")
    L.call("ffwm_max_pool2x2_backward", dev, _ptr(x), _ptr(grad_out), _ptr(grad_x), ctypes.c_int64(x.size(0) * x.size(1)), int(x.size(2)),
           int(x.size(3)), int(grad_out.size(2)), int(grad_out.size(3)))


def conv_few(x, weight, in_major, flip, bias, out, stride, pad):
    """Direct fp32 convolution for <= 4 input channels, or <= 4 output channels at stride 1 (csrc/conv_few.cu); `weight` is read
    through (in_major, flip), see include/ffwm_b200.h."""
    import ctypes
    dev = L.require_cuda(x, weight, out)
    if bias is not None and not (bias.is_cuda and bias.is_contiguous() and bias.dtype == x.dtype and bias.numel() == out.size(1)):
        raise ValueError("conv_few: bias must be a contiguous fp32 CUDA tensor of Cout elements")
    L.call("ffwm_conv_few", dev, L.t4(x), L.t4(weight), int(bool(in_major)), int(bool(flip)),
           ctypes.c_void_p(bias.data_ptr() if bias is not None else None), L.t4(out), int(stride), int(pad))
