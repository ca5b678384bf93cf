"""nn.BatchNorm2d on the streaming kernels of csrc/batch_norm.cu.

Every conv unit of the reference's networks is conv -> nn.BatchNorm2d -> nn.LeakyReLU (models/base_networks.py:15-45,
:179-206, :397-410), and a ResidualBlock ends in `activ(blocks(x) + input(x))` with a batch norm as the last block
(:208-233).  `BatchNorm2d` below is a drop-in nn.BatchNorm2d (same parameters, buffers and state_dict keys) whose
training-mode forward on a CUDA fp32 map runs two hand-written kernels per direction and can absorb the LeakyReLU that
follows it (`act_slope`) and the residual add (`forward(x, residual=...)`).  `fuse_activations()` rewrites a module list
accordingly, leaving an `AbsorbedLeakyReLU` placeholder where the activation was so that the indices inside the
nn.Sequential — and with them the reference's checkpoint keys — do not move.

Everything else (eval mode, CPU tensors, other dtypes, affine=False, momentum=None) takes torch's own path, with the
absorbed activation applied afterwards, so the module computes the same function everywhere.
FFWM_FUSED_BN=0 routes training mode through torch as well (the A/B switch).
"""
import copy
import functools
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd.function import once_differentiable

from . import ops

ENABLED = os.environ.get("FFWM_FUSED_BN", "1") == "1"


class BatchNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, momentum, eps, slope):
        x = x.contiguous()
        residual = residual.contiguous() if residual is not None else None
        y = torch.empty_like(x)
        save_mean = torch.empty(x.size(1), dtype=torch.float32, device=x.device)
        save_invstd = torch.empty_like(save_mean)
        ops.batch_norm_forward(x, residual, weight.detach(), bias.detach(), running_mean, running_var, float(momentum), float(eps),
                               float(slope), y, save_mean, save_invstd)
        ctx.slope = float(slope)
        ctx.has_residual = residual is not None
        # the activation's sign: recomputed from x in backward, except after a residual add (then the output carries it)
        ctx.save_for_backward(x, weight, bias, save_mean, save_invstd, y if (ctx.has_residual and ctx.slope != 1.0) else None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x, weight, bias, save_mean, save_invstd, y = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_x = torch.empty_like(x)
        grad_res = torch.empty_like(x) if ctx.has_residual else None
        grad_w, grad_b = torch.empty_like(weight), torch.empty_like(bias)
        ops.batch_norm_backward(x, grad_out, y, weight.detach(), bias.detach(), save_mean, save_invstd, ctx.slope, grad_x, grad_res,
                                grad_w, grad_b)
        need = ctx.needs_input_grad
        return (grad_x if need[0] else None, grad_res if need[1] else None, grad_w if need[2] else None, grad_b if need[3] else None,
                None, None, None, None, None)


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d; `act_slope` (None = no activation) is the LeakyReLU this layer applies to its own output."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.act_slope = None
        self._deferred = None                           # [pending calls] once a DeferredCounters owns this layer

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        new.__dict__ = copy.deepcopy(self.__dict__, memo)
        new._deferred = None            # a copy is not owned by the original's DeferredCounters: it counts by itself again
        return new

    def _fast(self, x):
        return (ENABLED and self.training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and self.affine
                and self.track_running_stats and self.momentum is not None and self.weight.dtype == torch.float32
                and x.size(0) * x.size(2) * x.size(3) > 1 and x.size(1) == self.num_features)

    def forward(self, x, residual=None, act_slope=None):
        slope = act_slope if act_slope is not None else self.act_slope
        if self._fast(x):
            if self.num_batches_tracked is not None:
                if self._deferred is not None:
                    self._deferred[0] += 1              # counted; DeferredCounters.flush() adds them in one launch
                else:
                    self.num_batches_tracked.add_(1)
            return BatchNormFunction.apply(x, residual, self.weight, self.bias, self.running_mean, self.running_var, self.momentum,
                                           self.eps, 1.0 if slope is None else slope)
        out = super().forward(x)
        if residual is not None:
            out = out + residual
        return out if slope is None else F.leaky_relu(out, slope)

    def extra_repr(self):
        s = super().extra_repr()
        return s if self.act_slope is None else s + ", act_slope=%g" % self.act_slope


class DeferredCounters:
    """`num_batches_tracked += 1` of every batch norm of `nets` as ONE multi-tensor add per `flush()` (a trainer calls it once
    per step) instead of one tiny kernel per layer call: 112 launches per FFWM train step.  The state_dict sees the same
    counts after every flush.  Layers outside a DeferredCounters (the reference's own orchestrators) keep torch's behaviour."""

    def __init__(self, nets):
        self.layers = [m for net in nets for m in net.modules() if isinstance(m, BatchNorm2d) and m.num_batches_tracked is not None]
        for m in self.layers:
            m._deferred = [0]

    def flush(self):
        live = [m for m in self.layers if m._deferred[0]]
        if live:
            torch._foreach_add_([m.num_batches_tracked for m in live], [int(m._deferred[0]) for m in live])
            for m in live:
                m._deferred[0] = 0


class AbsorbedLeakyReLU(nn.LeakyReLU):
    """Keeps the reference's position in an nn.Sequential; the BatchNorm2d in front of it applies the activation."""

    def forward(self, x):
        return x


def fuse_activations(modules):
    """[..., BatchNorm2d, nn.LeakyReLU, ...] -> [..., BatchNorm2d(act_slope), AbsorbedLeakyReLU, ...] (same length)."""
    out = list(modules)
    for i in range(len(out) - 1):
        bn, act = out[i], out[i + 1]
        if isinstance(bn, BatchNorm2d) and type(act) is nn.LeakyReLU and bn.act_slope is None:
            bn.act_slope = float(act.negative_slope)
            out[i + 1] = AbsorbedLeakyReLU(act.negative_slope, act.inplace)
    return out


def as_product_norm(norm):
    """The reference hands FlowNet `nn.BatchNorm2d` itself, or `functools.partial(nn.BatchNorm2d, affine=True,
    track_running_stats=True)` from models/networks.py:18-33 get_norm_layer, as the norm-layer factory."""
    if norm is nn.BatchNorm2d:
        return BatchNorm2d
    if isinstance(norm, functools.partial) and norm.func is nn.BatchNorm2d:
        return functools.partial(BatchNorm2d, *norm.args, **norm.keywords)
    return norm
