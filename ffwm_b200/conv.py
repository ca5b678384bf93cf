"""Convolution layers of the mirrored networks: `nn.Conv2d` / `nn.ConvTranspose2d` whose eligible calls run on the
hand-written tcgen05 kernels (ffwm_b200/csrc/conv3x3_tc.cu for the dominant 3x3 / stride-1 layers described first,
csrc/conv_gen_tc.cu for every other kernel size / stride / map size: see GENERAL below) instead of cuDNN.

Eligible: CUDA fp32, 3x3, stride 1, padding 1, dilation 1, groups 1, zero padding, map width 128, 64
or 32 — the generator's residual / attention / reconstruction / pixel-shuffle convolutions at all
three decoder scales, FlowNet's 3x3 layers down to 32x32, VGG19's blocks 1-3 and LightCNN's 3x3
layers at 64 and 32.  Everything else (strided 4x4 / 7x7 / 5x5 / 1x1 layers,
maps of 16x16 and below) stays on cuDNN (library).

    forward        conv3x3_tc (implicit GEMM, 3xTF32 split: library-grade fp32 accuracy)
    grad input     the same kernel with the weights packed transposed + flipped (3xBF16 split)
    grad weight    csrc/conv_gen_wgrad_tc.cu (MN-major tcgen05 GEMM over pixels, deterministic split-K), for every layer
FFWM_CONV_NT128=0 selects 64 (instead of 128) output channels per CTA for the W = 128 layers with more than 64 of them.

The module is a drop-in `nn.Conv2d`: same parameters, same state_dict keys, spectral norm hooks work
unchanged (the weight is re-packed on every call: 0.02 ms).
"""
import os

import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops

ENABLED = os.environ.get("FFWM_CONV_TC", "1") == "1"     # False / FFWM_CONV_TC=0: cuDNN everywhere (A/B measurements)
# Operand split per direction (include/ffwm_b200.h FFWM_MATH_*).  Forward passes run 3xTF32: their rounding errors decide
# which side of zero a pre-activation lands on, and with 3xBF16 forwards the GRADIENTS of the deep BatchNorm / LeakyReLU
# stacks come out 1000x less accurate than cuDNN's strict fp32 (FlowNet: 5e-3 vs 8e-6 of max|grad| against float64;
# with 3xTF32 forwards 1e-5; torch's default TF32: 8e-2 — profiles/r02h_network_accuracy.txt).  Gradients are linear in
# their operands' errors, so the data and weight gradients run 3xBF16 (twice the tensor rate) at no measurable loss.
MATH_FWD = int(os.environ.get("FFWM_CONV_MATH_FWD", ops.L.MATH_TF32X3))
MATH_BWD = int(os.environ.get("FFWM_CONV_MATH_BWD", ops.L.MATH_BF16X3))
# The kernel also handles width 16, but there a map is one or two CTAs per (image, channel tile) with a
# long serial K loop: measured slower than cuDNN inside the train step (98.4 -> 101.8 ms), so it is off.
WIDTHS = (128, 64, 32)
if os.environ.get("FFWM_CONV3X3_WIDTHS"):     # A/B: route some widths to the general kernel instead
    WIDTHS = tuple(int(v) for v in os.environ["FFWM_CONV3X3_WIDTHS"].split(","))
# 128 output channels per CTA for the W = 128 layers with more than 64 of them (measured: profiles/r02a_conv_nt128.txt;
# FFWM_CONV_NT128=0 for the A/B)
NT128 = os.environ.get("FFWM_CONV_NT128", "1") == "1"


def _nt(width, n_out):
    return 128 if (NT128 and width == 128 and n_out > 64) else 64


# Keep the packed image of FROZEN weights (VGG19, LightCNN: ~280 of the step's 380 packing launches at the start of round 2)
# instead of re-packing on every call (default on since round 2: profiles/r02b_switches.txt).  A weight qualifies while it is a leaf that does not require grad;
# the image is tagged with the tensor's version counter and data pointer, so any in-place update (optimizer step,
# load_state_dict, .data swap) re-packs.  Spectral-normed weights are new tensors on every call and never qualify.
CACHE_PACKED = os.environ.get("FFWM_CACHE_PACKED", "1") == "1"


def _cached(weight, key, pack):
    """Packed image of a FROZEN weight, kept on the tensor and re-made when its version counter or storage changes.
    A weight that has ever been seen with requires_grad=True is never served from the cache again: fused optimizers
    (torch.optim.Adam(fused=True)) update parameters without bumping `_version`, and the reference toggles
    requires_grad on the discriminator around the generator's backward pass (models/ffwm_model.py:206-207), so
    "does not require grad right now" does not mean "frozen"."""
    if weight.requires_grad:
        try:
            weight._ffwm_trainable = True
        except AttributeError:
            pass
    if not (CACHE_PACKED and weight.is_leaf and not weight.requires_grad and not getattr(weight, "_ffwm_trainable", False)):
        return pack()
    cache = getattr(weight, "_ffwm_packed", None)
    if cache is None:
        cache = weight._ffwm_packed = {}
    tag = (weight._version, weight.data_ptr())
    hit = cache.get(key)
    if hit is None or hit[0] != tag:
        hit = cache[key] = (tag, pack())
    return hit[1]


def _packed(weight, dgrad, nt, math):
    return _cached(weight, (dgrad, nt, math), lambda: ops.conv3x3_pack_weights(weight, dgrad=dgrad, nt=nt, math=math))


def _packed_gen(weight, in_major, stride, pad, transposed, math):
    return _cached(weight, ("gen", in_major, stride, pad, transposed, math),
                   lambda: ops.conv_pack_weights(weight, in_major, stride, pad, transposed, math))


def _bias_ok(bias, weight, x, out_dim=0):
    return bias is None or (bias.is_cuda and bias.device == x.device and bias.dtype == torch.float32 and bias.is_contiguous()
                            and bias.numel() == weight.size(out_dim))


def eligible(x, weight, stride, padding, dilation, groups, padding_mode="zeros", bias=None):
    """3x3 / stride 1 / pad 1 at W in {128, 64, 32} -> conv3x3_tc.  Anything the kernel would mis-read (channel mismatch,
    foreign bias, batch beyond the grid limit) is NOT eligible and reaches F.conv2d, which raises / handles it as torch does."""
    return (ENABLED and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 4
            and weight.device == x.device and x.size(1) == weight.size(1) and 0 < x.size(0) <= 65535
            and x.size(3) in WIDTHS and tuple(weight.shape[2:]) == (3, 3) and tuple(stride) == (1, 1)
            and tuple(padding) == (1, 1) and tuple(dilation) == (1, 1) and groups == 1 and padding_mode == "zeros"
            and _bias_ok(bias, weight, x))


_DIAG_FWD_LIB = _DIAG_DGRAD_LIB = False      # scripts/diag_nets.py only: one direction on the library

# Degenerate channel counts (csrc/conv_few.cu): <= 4 input channels, or ONE output channel at stride 1 on maps of at least 32x32
# with at most 512 input channels, run as direct fp32 FFMA kernels instead of tensor-core tiles that are 75-97 % padding.
# Measured in the train step (gpurun call 80): the few-input kernel replaces 2.4 ms of tensor-core launches with 1.0 ms, the
# one-output kernel (LightCNN stem's data gradient) 331 us with 190 us; with 2-4 output channels the direct kernel is
# instruction-bound and LOSES to the tensor cores (233 vs ~40 us on VGG conv1_1's data gradient), so those stay where they were.
# FFWM_CONV_FEW=0: tensor cores for everything.
FEW = os.environ.get("FFWM_CONV_FEW", "1") == "1"


def _few_forward(x, weight, bias, out, stride, pad):
    """conv2d(x, weight, bias, stride, pad) -> out on the direct kernels when the shape qualifies; False otherwise."""
    if not FEW:
        return False
    if weight.size(1) <= 4 or (weight.size(0) == 1 and stride == 1 and weight.size(1) <= 512 and out.size(2) * out.size(3) >= 1024):
        ops.conv_few(x, weight, False, False, bias, out, stride, pad)
        return True
    return False


def _few_dgrad(grad_out, weight, gx, stride, pad):
    """grad_input of conv2d(x, weight, stride 1, pad) -> gx: a convolution of grad_out with the transposed, flipped weight."""
    kh, kw = weight.shape[2:]
    if not FEW or stride != 1 or kh != kw or kh - 1 - pad < 0:
        return False
    if weight.size(0) <= 4 or (weight.size(1) == 1 and weight.size(0) <= 512 and gx.size(2) * gx.size(3) >= 1024):
        ops.conv_few(grad_out, weight, True, True, None, gx, 1, kh - 1 - pad)
        return True
    return False


class Conv3x3TCFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, math_fwd=None):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        if _DIAG_FWD_LIB:
            return torch.nn.functional.conv2d(x, weight, bias, padding=1)
        out = x.new_empty((x.size(0), weight.size(0), x.size(2), x.size(3)))
        if _few_forward(x, weight, bias, out, 1, 1):
            return out
        nt = _nt(x.size(3), weight.size(0))
        math = MATH_FWD if math_fwd is None else math_fwd
        ops.conv3x3_forward(x, _packed(weight, False, nt, math), bias, out, nt=nt, math=math)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable          # double backward (gradient penalties) is not provided: fails loudly
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0] and _DIAG_DGRAD_LIB:
            gx = torch.ops.aten.convolution_backward(grad_out, x, weight, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, False, False])[0]
        elif ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            if not _few_dgrad(grad_out, weight, gx, 1, 1):
                nt = _nt(grad_out.size(3), weight.size(1))
                ops.conv3x3_forward(grad_out, _packed(weight, True, nt, MATH_BWD), None, gx, nt=nt, math=MATH_BWD)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gw, gb = _wgrad(grad_out, x, weight, ctx.has_bias, 1, 1, False, 0, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw, gb, None


# ---------------------------------------------------------------------------------------------------------------
# Every other dense convolution (csrc/conv_gen_tc.cu): kernels up to 7x7, stride 1 or 2, any padding / map size,
# nn.Conv2d and nn.ConvTranspose2d, forward and data gradient on tcgen05 (3xBF16).  Weight gradients: cuDNN.
GENERAL = os.environ.get("FFWM_CONV_GENERAL", "1") == "1"


def _sym(v):
    v = tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    return v[0] if len(v) == 2 and v[0] == v[1] else None


def eligible_general(x, weight, stride, padding, dilation, groups, padding_mode="zeros", output_padding=(0, 0), transposed=False,
                     bias=None):
    if not (ENABLED and GENERAL and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 4
            and weight.device == x.device and groups == 1 and padding_mode == "zeros" and isinstance(padding, (tuple, list, int))
            and _bias_ok(bias, weight, x, 1 if transposed else 0)):
        return False
    s, p, d, op = _sym(stride), _sym(padding), _sym(dilation), _sym(output_padding)
    kh, kw = weight.shape[2:]
    return (s in (1, 2) and p is not None and 0 <= p <= 7 and d == 1 and op is not None and op < s and kh <= 7 and kw <= 7
            and min(kh, kw) >= (s if transposed else 1) and x.size(1) == weight.size(0 if transposed else 1)
            and x.size(0) > 0 and x.size(2) + 2 * p >= kh and x.size(3) + 2 * p >= kw)


# weight gradient of the general layers on tcgen05 (csrc/conv_gen_wgrad_tc.cu); FFWM_WGRAD_GEN=0: cuDNN through aten
WGRAD_GEN = os.environ.get("FFWM_WGRAD_GEN", "1") == "1"


def _wgrad(grad_out, x, weight, has_bias, stride, pad, transposed, out_pad, need_w, need_b):
    if not WGRAD_GEN:
        return _wgrad_library(grad_out, x, weight, has_bias, stride, pad, transposed, out_pad, need_w, need_b)
    gw = gb = None
    if need_w:
        gw = torch.empty_like(weight, memory_format=torch.contiguous_format)
        small, large = (x, grad_out) if transposed else (grad_out, x)
        ops.conv_wgrad(small, large, gw, stride, pad)
    if has_bias and need_b:
        gb = ops.channel_sum(grad_out) if (grad_out.is_contiguous() and grad_out.numel() > 0) else grad_out.sum((0, 2, 3))
    return gw, gb


def _wgrad_library(grad_out, x, weight, has_bias, stride, pad, transposed, out_pad, need_w, need_b):
    """Weight / bias gradient of the general layers: cuDNN through aten (grad_input is NOT requested)."""
    _, gw, gb = torch.ops.aten.convolution_backward(
        grad_out, x, weight, [weight.size(1 if transposed else 0)] if has_bias else None, [stride, stride], [pad, pad], [1, 1],
        transposed, [out_pad, out_pad], 1, [False, bool(need_w), bool(has_bias and need_b)])
    return gw, gb


class ConvGenFunction(Function):
    """conv2d(x, W, b, stride, pad) — forward: conv_forward; grad_input: the transposed mode on the same weights."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, math_fwd=None):
        ctx.save_for_backward(x, weight)
        ctx.cfg = (bias is not None, stride, pad)
        kh, kw = weight.shape[2:]
        out = x.new_empty((x.size(0), weight.size(0), (x.size(2) + 2 * pad - kh) // stride + 1, (x.size(3) + 2 * pad - kw) // stride + 1))
        if _few_forward(x, weight, bias, out, stride, pad):
            return out
        math = MATH_FWD if math_fwd is None else math_fwd
        ops.conv_forward(x, _packed_gen(weight, False, stride, pad, False, math), bias, out, kh, kw, stride, pad, False, math)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        has_bias, stride, pad = ctx.cfg
        kh, kw = weight.shape[2:]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x, memory_format=torch.contiguous_format)
            if not _few_dgrad(grad_out, weight, gx, stride, pad):
                ops.conv_forward(grad_out, _packed_gen(weight, True, stride, pad, True, MATH_BWD), None, gx, kh, kw, stride, pad, True, MATH_BWD)
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            gw, gb = _wgrad(grad_out, x, weight, has_bias, stride, pad, False, 0, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw, gb, None, None, None


class ConvTGenFunction(Function):
    """conv_transpose2d(x, W, b, stride, pad, output_padding) — forward: the transposed mode; grad_input: a convolution."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, out_pad):
        ctx.save_for_backward(x, weight)
        ctx.cfg = (bias is not None, stride, pad, out_pad)
        kh, kw = weight.shape[2:]
        out = x.new_empty((x.size(0), weight.size(1), (x.size(2) - 1) * stride - 2 * pad + kh + out_pad,
                           (x.size(3) - 1) * stride - 2 * pad + kw + out_pad))
        ops.conv_forward(x, _packed_gen(weight, True, stride, pad, True, MATH_FWD), bias, out, kh, kw, stride, pad, True, MATH_FWD)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        has_bias, stride, pad, out_pad = ctx.cfg
        kh, kw = weight.shape[2:]
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x, memory_format=torch.contiguous_format)
            # a convolution over grad_out; with output_padding < stride its floor division still yields x's size
            ops.conv_forward(grad_out, _packed_gen(weight, False, stride, pad, False, MATH_BWD), None, gx, kh, kw, stride, pad, False, MATH_BWD)
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            gw, gb = _wgrad(grad_out, x, weight, has_bias, stride, pad, True, out_pad, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return gx, gw, gb, None, None, None


def conv2d(x, weight, bias, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1):
    """F.conv2d with the tcgen05 paths for eligible calls."""
    if eligible(x, weight, stride, padding, dilation, groups, bias=bias):
        return Conv3x3TCFunction.apply(x, weight, bias)
    if eligible_general(x, weight, stride, padding, dilation, groups, bias=bias):
        return ConvGenFunction.apply(x, weight, bias, _sym(stride), _sym(padding))
    return torch.nn.functional.conv2d(x, weight, bias, stride, padding, dilation, groups)


class Conv2d(nn.Conv2d):
    math_fwd = None          # operand split of THIS layer's forward pass (None: the module-wide MATH_FWD); see set_forward_math

    def _conv_forward(self, input, weight, bias):
        if isinstance(self.padding, tuple):
            if eligible(input, weight, self.stride, self.padding, self.dilation, self.groups, self.padding_mode, bias=bias):
                return Conv3x3TCFunction.apply(input, weight, bias, self.math_fwd)
            if eligible_general(input, weight, self.stride, self.padding, self.dilation, self.groups, self.padding_mode, bias=bias):
                return ConvGenFunction.apply(input, weight, bias, _sym(self.stride), _sym(self.padding), self.math_fwd)
        return super()._conv_forward(input, weight, bias)


# The frozen feature extractors of the losses (VGG19, LightCNN-29) run their forward passes in the 3xBF16 split: their outputs
# are as close to float64 as with 3xTF32 (LightCNN fc / pool 1.4e-5 / 2.6e-5 vs 3.2e-5 / 3.7e-5, profiles/r02h_network_accuracy.txt)
# — what 3xBF16 forwards cost is gradient accuracy of the deep BatchNorm / LeakyReLU stacks being TRAINED (FlowNet 1.3e-2 vs
# 1.9e-5, the discriminator 7.6e-3 vs 1.2e-5), and those keep 3xTF32.  FFWM_LOSSNET_MATH_FWD=0 puts them back on 3xTF32.
LOSSNET_MATH_FWD = int(os.environ.get("FFWM_LOSSNET_MATH_FWD", ops.L.MATH_BF16X3))


def set_forward_math(net, math):
    """Operand split of the forward pass of every Conv2d of `net` (ops.L.MATH_TF32X3 / MATH_BF16X3; None = module default)."""
    for m in net.modules():
        if isinstance(m, Conv2d):
            m.math_fwd = math
    return net


class ConvTranspose2d(nn.ConvTranspose2d):
    """Drop-in nn.ConvTranspose2d (same parameters / state_dict keys) on the tcgen05 kernel for eligible calls."""

    def forward(self, input, output_size=None):
        if (output_size is None and isinstance(self.padding, tuple)
                and eligible_general(input, self.weight, self.stride, self.padding, self.dilation, self.groups, self.padding_mode,
                                     self.output_padding, transposed=True, bias=self.bias)):
            return ConvTGenFunction.apply(input, self.weight, self.bias, _sym(self.stride), _sym(self.padding), _sym(self.output_padding))
        return super().forward(input, output_size)
