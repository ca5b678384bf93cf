"""Host-side mirror of the reference's `models/external_function.py`.

Same class names, constructor/forward signatures and error behaviour as the
reference (file:line cited per class) so that losses/models written against
it run unchanged; the arithmetic is the hand-written sm_100a kernels behind
the C ABI (ffwm_b200/csrc, include/ffwm_b200.h).  CUDA only, like the
reference: a CPU tensor raises NotImplementedError.

Differences that are deliberate and visible:
  * outputs that the kernels fully overwrite are allocated with `empty`
    instead of `zeros` (the reference pays a memset per call and then
    atomicAdds into it); scatter targets are still zero-filled;
  * gradients that autograd does not need are not computed;
  * `GridWarpFunction` is new: it is WarpNet's `F.grid_sample` as a kernel
    that reads the (B,2,H,W) flow directly;
  * `Resample2d` builds its sigma plane on the input's device (the reference
    keeps it on the CPU and would fail on GPU input).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import ops

# the guided filter as four kernels per direction (csrc/guided_filter.cu): parity-green on a B200 (tests/test_guided_filter_gpu.py),
# train step 95.2 -> 90.7 ms (profiles/r02a_switches.txt).  FFWM_FUSED_GF=0 restores the cumsum formulation for A/B runs.
FUSED_GF = os.environ.get("FFWM_FUSED_GF", "1") == "1"


def _cuda_only(t):
    if not t.is_cuda:
        raise NotImplementedError  # models/external_function.py:37-38


class AffineResidualFunction(Function):
    """Per-window affine-fit residual w^T Q w / kz^2 of one grid plane (csrc/affine_reg.cu): the fused form of the
    reference's conv2d -> LocalAttnReshape -> BlockExtractor -> multiply -> avg_pool2d chain (models/losses.py:211-219)."""

    @staticmethod
    def forward(ctx, grid, q, kz):
        _cuda_only(grid)
        grid = grid.contiguous()
        ctx.save_for_backward(grid, q)
        ctx.kz = kz
        out = grid.new_empty((grid.size(0), 1, grid.size(2) - kz + 1, grid.size(3) - kz + 1))
        ops.affine_reg_forward(grid, q, out, kz)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        grid, q = ctx.saved_tensors
        grad_grid = torch.zeros_like(grid)                   # scatter target
        ops.affine_reg_backward(grid, q, grad_out.contiguous(), grad_grid, ctx.kz)
        return grad_grid, None, None


class BlockExtractorFunction(Function):
    """models/external_function.py:19-56."""

    @staticmethod
    def forward(ctx, source, flow_field, kernel_size):
        assert source.is_contiguous()
        assert flow_field.is_contiguous()
        bs, ds, hs, ws = source.size()
        bf, df, hf, wf = flow_field.size()
        assert df == 2
        _cuda_only(source)
        ctx.save_for_backward(source, flow_field)
        ctx.kernel_size = kernel_size
        # dtype/device follow flow_field, as in the reference (`flow_field.new`)
        output = flow_field.new_empty((bs, ds, kernel_size * hf, kernel_size * wf))
        ops.block_extractor_forward(source, flow_field, output, kernel_size)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        source, flow_field = ctx.saved_tensors
        need_src, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_source = torch.zeros_like(source) if need_src else None          # scatter target
        grad_flow = torch.empty_like(flow_field) if need_flow else None       # overwritten
        if need_src or need_flow:
            # grad_output may be non-contiguous: the kernels honour strides
            ops.block_extractor_backward(source, flow_field, grad_output, grad_source, grad_flow,
                                         ctx.kernel_size)
        return grad_source, grad_flow, None


class BlockExtractor(nn.Module):
    """models/external_function.py:58-67."""

    def __init__(self, kernel_size=3):
        super().__init__()
        self.kernel_size = kernel_size

    def forward(self, source, flow_field):
        return BlockExtractorFunction.apply(source.contiguous(), flow_field.contiguous(), self.kernel_size)


class LocalAttnReshapeFunction(Function):
    """models/external_function.py:69-100."""

    @staticmethod
    def forward(ctx, inputs, kernel_size):
        assert inputs.is_contiguous()
        bs, ds, hs, ws = inputs.size()
        assert ds == kernel_size * kernel_size
        _cuda_only(inputs)
        ctx.kernel_size = kernel_size
        ctx.in_shape = inputs.shape
        output = inputs.new_empty((bs, 1, kernel_size * hs, kernel_size * ws))
        ops.local_attn_reshape_forward(inputs, output, kernel_size)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        grad_inputs = grad_output.new_empty(ctx.in_shape)
        ops.local_attn_reshape_backward(grad_output, grad_inputs, ctx.kernel_size)
        return grad_inputs, None


class LocalAttnReshape(nn.Module):
    """models/external_function.py:102-109 (kernel size is a forward argument)."""

    def __init__(self):
        super().__init__()

    def forward(self, inputs, kernel_size=3):
        return LocalAttnReshapeFunction.apply(inputs.contiguous(), kernel_size)


class Resample2dFunction(Function):
    """models/external_function.py:111-144.  input2 = (dx, dy, sigma)."""

    @staticmethod
    def forward(ctx, input1, input2, kernel_size=2, dilation=1):
        assert input1.is_contiguous()
        assert input2.is_contiguous()
        _cuda_only(input1)
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size = kernel_size
        ctx.dilation = dilation
        _, d, _, _ = input1.size()
        b, _, h, w = input2.size()
        output = input1.new_empty((b, d, h, w))
        ops.resample2d_forward(input1, input2, output, kernel_size, dilation)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_input1 = torch.zeros_like(input1) if need1 else None                    # scatter target
        grad_input2 = input1.new_empty(input2.size()) if need2 else None             # overwritten
        if need1 or need2:
            ops.resample2d_backward(input1, input2, grad_output, grad_input1, grad_input2,
                                    ctx.kernel_size, ctx.dilation)
        return grad_input1, grad_input2, None, None


class Resample2d(nn.Module):
    """models/external_function.py:146-158."""

    def __init__(self, kernel_size=2, dilation=1, sigma=5):
        super().__init__()
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.sigma = torch.tensor(sigma, dtype=torch.float)

    def forward(self, input1, input2):
        sigma = self.sigma.to(device=input2.device, dtype=input2.dtype).expand(
            input2.size(0), 1, input2.size(2), input2.size(3))
        return Resample2dFunction.apply(input1.contiguous(), torch.cat((input2, sigma), 1),
                                        self.kernel_size, self.dilation)


class GridWarpFunction(Function):
    """WarpNet's sampler (models/base_networks.py:173): bilinear grid_sample,
    zeros padding, align_corners=False, flow kept as (B,2,H,W)."""

    @staticmethod
    def forward(ctx, images, flow):
        _cuda_only(images)
        ctx.save_for_backward(images, flow)
        b, _, h, w = flow.size()
        output = images.new_empty((b, images.size(1), h, w))
        ops.grid_warp_forward(images, flow, output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        images, flow = ctx.saved_tensors
        need_img, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_images = torch.zeros_like(images) if need_img else None     # scatter target
        grad_flow = torch.empty_like(flow) if need_flow else None        # overwritten
        if need_img or need_flow:
            ops.grid_warp_backward(images, flow, grad_output, grad_images, grad_flow)
        return grad_images, grad_flow


def grid_warp(images, flow):
    """Functional form used by base_networks.WarpNet."""
    if flow.dtype != images.dtype:
        flow = flow.to(images.dtype)
    return GridWarpFunction.apply(images, flow)


# --------------------------------------------------------------------------
# Guided filter (models/external_function.py:164-277), torch ops for now.
# Box sums are window sums truncated at the border, taken as differences of
# one running sum per axis — the same cumsum entries the reference subtracts,
# hence bit-identical results.
# --------------------------------------------------------------------------
def _window_sum(x, r, dim):
    n = x.size(dim)
    c = x.cumsum(dim=dim)
    idx = torch.arange(n, device=x.device)
    upper = c.index_select(dim, (idx + r).clamp(max=n - 1))
    lo_idx = idx - r - 1
    lower = c.index_select(dim, lo_idx.clamp(min=0))
    shape = [1] * x.dim()
    shape[dim] = n
    lower = lower * (lo_idx >= 0).to(x.dtype).view(shape)
    return upper - lower


class BoxFilter(nn.Module):
    """models/external_function.py:187-195."""

    def __init__(self, r):
        super().__init__()
        self.r = r

    def forward(self, x):
        assert x.dim() == 4
        return _window_sum(_window_sum(x, self.r, 2), self.r, 3)


class GuidedFilterFunction(Function):
    """GuidedFilter.forward (models/external_function.py:239-277) as four kernels per direction
    (csrc/guided_filter.cu); gradient with respect to x only (y is data for every caller in the reference)."""

    @staticmethod
    def forward(ctx, x, y, r, eps):
        _cuda_only(x)
        x, y = x.contiguous(), y.contiguous()
        q = torch.empty_like(x)
        save = x.new_empty((5,) + tuple(x.shape))
        ops.guided_filter_forward(x, y, q, save, x.new_empty((5,) + tuple(x.shape)), r, eps)
        ctx.save_for_backward(x, y, save)
        ctx.r = r
        return q

    @staticmethod
    def backward(ctx, grad_q):
        x, y, save = ctx.saved_tensors
        grad_x = torch.empty_like(x)
        ops.guided_filter_backward(x, y, grad_q.contiguous(), save, grad_x, x.new_empty((6,) + tuple(x.shape)), ctx.r)
        return grad_x, None, None, None


class GuidedFilter(nn.Module):
    """models/external_function.py:239-277: q = mean(A) * x + mean(b)."""

    def __init__(self, r, eps=1e-8):
        super().__init__()
        self.r = r
        self.eps = eps
        self.boxfilter = BoxFilter(r)

    def forward(self, x, y):
        n_x, c_x, h_x, w_x = x.size()
        n_y, c_y, h_y, w_y = y.size()
        assert n_x == n_y
        assert c_x == 1 or c_x == c_y
        assert h_x == h_y and w_x == w_y
        assert h_x > 2 * self.r + 1 and w_x > 2 * self.r + 1

        if (FUSED_GF and x.is_cuda and x.dtype == torch.float32 and y.dtype == torch.float32 and c_x == c_y
                and not y.requires_grad):
            return GuidedFilterFunction.apply(x, y, self.r, self.eps)
        count = self.boxfilter(x.new_ones((1, 1, h_x, w_x)))
        mean_x = self.boxfilter(x) / count
        mean_y = self.boxfilter(y) / count
        cov_xy = self.boxfilter(x * y) / count - mean_x * mean_y
        var_x = self.boxfilter(x * x) / count - mean_x * mean_x
        A = cov_xy / (var_x + self.eps)
        b = mean_y - A * mean_x
        mean_A = self.boxfilter(A) / count
        mean_b = self.boxfilter(b) / count
        return mean_A * x + mean_b


class FastGuidedFilter(nn.Module):
    """models/external_function.py:197-237: coefficients at low resolution,
    bilinearly upsampled (align_corners=True) and applied to hr_x."""

    def __init__(self, r, eps=1e-8):
        super().__init__()
        self.r = r
        self.eps = eps
        self.boxfilter = BoxFilter(r)

    def forward(self, lr_x, lr_y, hr_x):
        n_lrx, c_lrx, h_lrx, w_lrx = lr_x.size()
        n_lry, c_lry, h_lry, w_lry = lr_y.size()
        n_hrx, c_hrx, h_hrx, w_hrx = hr_x.size()
        assert n_lrx == n_lry and n_lry == n_hrx
        assert c_lrx == c_hrx and (c_lrx == 1 or c_lrx == c_lry)
        assert h_lrx == h_lry and w_lrx == w_lry
        assert h_lrx > 2 * self.r + 1 and w_lrx > 2 * self.r + 1

        count = self.boxfilter(lr_x.new_ones((1, 1, h_lrx, w_lrx)))
        mean_x = self.boxfilter(lr_x) / count
        mean_y = self.boxfilter(lr_y) / count
        cov_xy = self.boxfilter(lr_x * lr_y) / count - mean_x * mean_y
        var_x = self.boxfilter(lr_x * lr_x) / count - mean_x * mean_x
        A = cov_xy / (var_x + self.eps)
        b = mean_y - A * mean_x
        mean_A = F.interpolate(A, (h_hrx, w_hrx), mode='bilinear', align_corners=True)
        mean_b = F.interpolate(b, (h_hrx, w_hrx), mode='bilinear', align_corners=True)
        return mean_A * hr_x + mean_b
