"""Convolution layer of the mirrored networks: `nn.Conv2d` whose eligible calls run on the
hand-written tcgen05 kernel (ffwm_b200/csrc/conv3x3_tc.cu) instead of cuDNN.

Eligible: CUDA fp32, 3x3, stride 1, padding 1, dilation 1, groups 1, zero padding, map width 128, 64
or 32 — the generator's residual / attention / reconstruction / pixel-shuffle convolutions at all
three decoder scales, FlowNet's 3x3 layers down to 32x32, VGG19's blocks 1-3 and LightCNN's 3x3
layers at 64 and 32.  Everything else (strided 4x4 / 7x7 / 5x5 / 1x1 layers,
maps of 16x16 and below) stays on cuDNN (library).

    forward        conv3x3_tc (implicit GEMM, 3xTF32 split: fp32-level accuracy)
    grad input     the same kernel with the weights packed transposed + flipped
    grad weight    cuDNN (`aten.convolution_backward`, weight/bias outputs only) by default;
                   FFWM_WGRAD_TC=1 opts in to the tcgen05 weight-gradient kernel (csrc/conv3x3_wgrad_tc.cu),
                   which is EXPERIMENTAL: written after the round-1 GPU budget was spent, checked by a CPU
                   emulation of its indexing only, not yet run or measured on a B200
FFWM_CONV_NT128=1 selects 128 (instead of 64) output channels per CTA for the W = 128 layers with more than
64 output channels — also experimental and unmeasured; the validated kernels are bit-identical either way.

The module is a drop-in `nn.Conv2d`: same parameters, same state_dict keys, spectral norm hooks work
unchanged (the weight is re-packed on every call: 0.02 ms).
"""
import os

import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops

ENABLED = True          # set False to force cuDNN everywhere (A/B measurements)
# The kernel also handles width 16, but there a map is one or two CTAs per (image, channel tile) with a
# long serial K loop: measured slower than cuDNN inside the train step (98.4 -> 101.8 ms), so it is off.
WIDTHS = (128, 64, 32)
WGRAD_TC = os.environ.get("FFWM_WGRAD_TC", "0") == "1"     # experimental, unmeasured: off unless asked for
# its CTA tile is 128 output x 48 input channels: layers far below that (flow heads, RGB reconstructions,
# the first convolutions on 3 channels: 0.7 % of the weight-gradient FLOPs of the step) stay on cuDNN
WGRAD_MIN_COUT, WGRAD_MIN_CIN = 32, 16
# experimental, unmeasured: 128 output channels per CTA for the W = 128 layers with more than 64 of them
NT128 = os.environ.get("FFWM_CONV_NT128", "1") == "1"


def _nt(width, n_out):
    return 128 if (NT128 and width == 128 and n_out > 64) else 64


# experimental, unmeasured: keep the packed image of FROZEN weights (VGG19, LightCNN: ~280 of the step's 380 packing
# launches) instead of re-packing on every call.  A weight qualifies while it is a leaf that does not require grad;
# the image is tagged with the tensor's version counter and data pointer, so any in-place update (optimizer step,
# load_state_dict, .data swap) re-packs.  Spectral-normed weights are new tensors on every call and never qualify.
CACHE_PACKED = os.environ.get("FFWM_CACHE_PACKED", "1") == "1"


def _packed(weight, dgrad, nt):
    if not (CACHE_PACKED and weight.is_leaf and not weight.requires_grad):
        return ops.conv3x3_pack_weights(weight, dgrad=dgrad, nt=nt)
    cache = getattr(weight, "_ffwm_packed", None)
    if cache is None:
        cache = weight._ffwm_packed = {}
    tag = (weight._version, weight.data_ptr())
    hit = cache.get((dgrad, nt))
    if hit is None or hit[0] != tag:
        hit = cache[(dgrad, nt)] = (tag, ops.conv3x3_pack_weights(weight, dgrad=dgrad, nt=nt))
    return hit[1]


def eligible(x, weight, stride, padding, dilation, groups, padding_mode="zeros"):
    return (ENABLED and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.dim() == 4
            and x.size(3) in WIDTHS and tuple(weight.shape[2:]) == (3, 3) and tuple(stride) == (1, 1)
            and tuple(padding) == (1, 1) and tuple(dilation) == (1, 1) and groups == 1 and padding_mode == "zeros")


class Conv3x3TCFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        out = x.new_empty((x.size(0), weight.size(0), x.size(2), x.size(3)))
        nt = _nt(x.size(3), weight.size(0))
        ops.conv3x3_forward(x, _packed(weight, False, nt), bias, out, nt=nt)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(x)
            nt = _nt(grad_out.size(3), weight.size(1))
            ops.conv3x3_forward(grad_out, _packed(weight, True, nt), None, gx, nt=nt)
        if WGRAD_TC and weight.size(0) >= WGRAD_MIN_COUT and weight.size(1) >= WGRAD_MIN_CIN:
            want_b = ctx.has_bias and ctx.needs_input_grad[2]
            if ctx.needs_input_grad[1]:
                gw = torch.zeros_like(weight, memory_format=torch.contiguous_format)
                gb = grad_out.new_zeros(weight.size(0)) if want_b else None    # falls out of staging grad_out
                ops.conv3x3_wgrad(x, grad_out, gw, gb)
            elif want_b:
                gb = grad_out.sum((0, 2, 3))
        elif ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            _, gw, gb = torch.ops.aten.convolution_backward(
                grad_out, x, weight, [weight.size(0)] if ctx.has_bias else None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                [False, bool(ctx.needs_input_grad[1]), bool(ctx.has_bias and ctx.needs_input_grad[2])])
        return gx, gw, gb


def conv2d(x, weight, bias, stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1):
    """F.conv2d with the tcgen05 path for eligible calls."""
    if eligible(x, weight, stride, padding, dilation, groups):
        return Conv3x3TCFunction.apply(x, weight, bias)
    return torch.nn.functional.conv2d(x, weight, bias, stride, padding, dilation, groups)


class Conv2d(nn.Conv2d):
    def _conv_forward(self, input, weight, bias):
        if isinstance(self.padding, tuple) and eligible(input, weight, self.stride, self.padding, self.dilation,
                                                        self.groups, self.padding_mode):
            return Conv3x3TCFunction.apply(input, weight, bias)
        return super()._conv_forward(input, weight, bias)
