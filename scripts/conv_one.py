#!/usr/bin/env python
"""One shape of benchmarks/conv.py's general list, a few launches (for ncu):  python scripts/conv_one.py "<name substring>" [math]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmarks.conv import GEN_SHAPES
from ffwm_b200 import ops
pat, math = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
name, b, cin, cout, r, k, s, p, tr = [g for g in GEN_SHAPES if pat in g[0]][0]
dev = torch.device("cuda", 0)
x = torch.randn(b, cin, r, r, device=dev)
w = torch.randn((cin, cout, k, k) if tr else (cout, cin, k, k), device=dev) * 0.05
ho = (r - 1) * s - 2 * p + k if tr else (r + 2 * p - k) // s + 1
out = torch.empty(b, cout, ho, ho, device=dev)
packed = ops.conv_pack_weights(w, int(tr), s, p, tr, math)
for _ in range(6):
    ops.conv_forward(x, packed, None, out, k, k, s, p, tr, math)
torch.cuda.synchronize()
print(name, "ok")
