#!/usr/bin/env python
"""Run selected warp ops a few times at the cfg5 primary point (for ncu).  python scripts/roll_one.py op [op...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks.warp import make_inputs  # noqa: E402
from ffwm_b200 import ops  # noqa: E402

c, r = 128, 128
B = 64
t = make_inputs(dict(B=B, Bb=16, B2=8, C=c, R=r), torch.device("cuda", 0), seed=1234)
e = torch.empty_like
out, g1, g2, gf = e(t["feat"]), torch.zeros_like(t["feat"]), e(t["disp"]), e(t["grid"])
ob, gs, gfl = e(t["be_gout"]), torch.zeros_like(t["be_src"]), e(t["be_flow"])
fns = {
    "rs4_fwd": lambda: ops.resample2d_forward(t["feat"], t["disp"], out, 4, 1),
    "rs2_fwd": lambda: ops.resample2d_forward(t["feat"], t["disp"], out, 2, 1),
    "rs4_gflow": lambda: ops.resample2d_backward(t["feat"], t["disp"], t["gout"], None, g2, 4, 1),
    "rs4_gin1": lambda: ops.resample2d_backward(t["feat"], t["disp"], t["gout"], g1, None, 4, 1),
    "gw_fwd": lambda: ops.grid_warp_forward(t["feat"], t["grid"], out),
    "gw_gflow": lambda: ops.grid_warp_backward(t["feat"], t["grid"], t["gout"], None, gf),
    "gw_gimg": lambda: ops.grid_warp_backward(t["feat"], t["grid"], t["gout"], g1, None),
    "be_fwd": lambda: ops.block_extractor_forward(t["be_src"], t["be_flow"], ob, 3),
    "be_bwd": lambda: ops.block_extractor_backward(t["be_src"], t["be_flow"], t["be_gout"], gs, gfl, 3),
}
for name in sys.argv[1:]:
    for _ in range(3):
        fns[name]()
torch.cuda.synchronize()
