#!/usr/bin/env python
"""A few launches of the batch-norm kernels on netG's dres2 map (8x195x128x128) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ffwm_b200 import ops
dev = "cuda:0"
shape = (8, 195, 128, 128)
x, go, res = (torch.randn(shape, device=dev) for _ in range(3))
y, gx, gr = (torch.empty(shape, device=dev) for _ in range(3))
c = shape[1]
w, b = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
rm, rv, sm, si, gw, gb = (torch.zeros(c, device=dev) for _ in range(6))
for _ in range(3):
    ops.batch_norm_forward(x, res, w, b, rm, rv, 0.1, 1e-5, 0.2, y, sm, si)
    ops.batch_norm_backward(x, go, y, w, b, sm, si, 0.2, gx, gr, gw, gb)
    ops.channel_sum(go)
torch.cuda.synchronize()
print("ok")
