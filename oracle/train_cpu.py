"""ORACLE (test infrastructure, not product code): the FFWM / FlowNet training steps on the CPU.

The reference's networks and losses are compositions of PyTorch ops, and its only CPU-capable
warp is `F.grid_sample` (models/base_networks.py:173); its three custom ops have no CPU path at
all.  This module therefore runs the train step the way the reference's CPU path does — PyTorch
CPU kernels (oneDNN) for the convolutions, `F.grid_sample` for every warp — by building the
host-side mirror classes with the warp swapped for the torch op.  For FlowNet pre-training the
custom ops are served by the C oracle (oracle/liboracle.so) wrapped in autograd Functions.

Used by tests/ (parity of the CUDA train step) and by bench.py's cpu_baseline / --impl reference
legs only.
"""
import contextlib

import torch
import torch.nn.functional as F

from . import warp as W


def _torch_grid_warp(images, flow):
    return F.grid_sample(images, flow.permute(0, 2, 3, 1), mode='bilinear', padding_mode='zeros', align_corners=False)


class _BlockExtractorCPU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, source, flow, k):
        ctx.save_for_backward(source, flow)
        ctx.k = k
        return W.block_extractor_forward(source.contiguous(), flow.contiguous(), k)

    @staticmethod
    def backward(ctx, g):
        s, f = ctx.saved_tensors
        gs, gf = W.block_extractor_backward(s.contiguous(), f.contiguous(), g.contiguous(), ctx.k)
        return gs, gf, None


class _LocalAttnReshapeCPU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        ctx.k, ctx.ref = k, x
        return W.local_attn_reshape_forward(x.contiguous(), k)

    @staticmethod
    def backward(ctx, g):
        return W.local_attn_reshape_backward(ctx.ref, g.contiguous(), ctx.k), None


@contextlib.contextmanager
def cpu_ops():
    """Temporarily route the mirror classes' warps to the CPU restatements."""
    from ffwm_b200 import base_networks, external_function as EF, losses
    saved = (EF.grid_warp, losses.grid_warp, EF.BlockExtractor.forward, EF.LocalAttnReshape.forward)
    EF.grid_warp = losses.grid_warp = _torch_grid_warp
    EF.BlockExtractor.forward = lambda self, s, f: _BlockExtractorCPU.apply(s, f, self.kernel_size)
    EF.LocalAttnReshape.forward = lambda self, x, kernel_size=3: _LocalAttnReshapeCPU.apply(x, kernel_size)
    try:
        yield
    finally:
        EF.grid_warp, losses.grid_warp, EF.BlockExtractor.forward, EF.LocalAttnReshape.forward = saved


def synthetic_batch(b, seed, titers=30000):
    """SURVEY 8(d) cfg3 batch."""
    g = torch.Generator().manual_seed(seed)
    return {
        'img_S': torch.rand(b, 3, 128, 128, generator=g), 'img_F': torch.rand(b, 3, 128, 128, generator=g),
        'mask_F': (torch.rand(b, 1, 128, 128, generator=g) > 0.3).float(),
        'mask_S': (torch.rand(b, 1, 128, 128, generator=g) > 0.3).float(),
        'lm_F': torch.randint(20, 108, (b, 1000, 2), generator=g), 'lm_S': torch.randint(20, 108, (b, 1000, 2), generator=g),
        'gate': (torch.rand(b, 1000, 1, generator=g) > 0.2).float(),
        'titers': titers, 'epoch': 0, 'input_path': [''] * b,
    }
