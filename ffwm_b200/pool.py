"""nn.MaxPool2d(2, 2[, ceil_mode]) / F.max_pool2d(x, 2) of LightCNN-29 and VGG19 (lightcnn/light_cnn.py:38-42,96-124,
models/losses.py:430-470) on csrc/pool.cu: no int64 index map, the backward pass recomputes the argmax from the input.
Anything else (other kernel sizes, CPU tensors, other dtypes) takes torch's path.  FFWM_FUSED_POOL=0: torch everywhere."""
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd.function import once_differentiable

from . import ops

ENABLED = os.environ.get("FFWM_FUSED_POOL", "1") == "1"


class MaxPool2x2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ceil_mode):
        x = x.contiguous()
        h, w = x.shape[2:]
        ho, wo = ((h + 1) // 2, (w + 1) // 2) if ceil_mode else (h // 2, w // 2)
        out = x.new_empty((x.size(0), x.size(1), ho, wo))
        ops.max_pool2x2_forward(x, out)
        ctx.save_for_backward(x)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        grad_x = torch.empty_like(x)
        ops.max_pool2x2_backward(x, grad_out.contiguous(), grad_x)
        return grad_x, None


def eligible(x):
    return (ENABLED and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.size(2) >= 2 and x.size(3) >= 2 and x.numel() > 0)


def max_pool2x2(x, ceil_mode=False):
    """F.max_pool2d(x, 2, 2, ceil_mode=ceil_mode)"""
    if eligible(x):
        return MaxPool2x2Function.apply(x, bool(ceil_mode))
    return F.max_pool2d(x, 2, 2, ceil_mode=ceil_mode)


class MaxPool2d(nn.MaxPool2d):
    def forward(self, x):
        def two(v):
            return v == 2 or v == (2, 2)
        if (two(self.kernel_size) and two(self.stride) and self.padding in (0, (0, 0)) and self.dilation in (1, (1, 1))
                and not self.return_indices and eligible(x)):
            return MaxPool2x2Function.apply(x, bool(self.ceil_mode))
        return super().forward(x)
