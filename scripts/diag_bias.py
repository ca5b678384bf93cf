"""Is the tensor core's fp32 accumulation biased?  conv(x) + conv(-x) is exactly zero for a symmetric rounding and a
coherent negative number for truncation toward -inf.  Also: mean signed error against float64."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffwm_b200 import _lib, ops
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
g = torch.Generator().manual_seed(0)
for name, cin, cout, r, relu in (("195->195 @128 relu-x", 195, 195, 128, True), ("195->195 @128 randn-x", 195, 195, 128, False), ("384->384 @32 relu-x", 384, 384, 32, True)):
    x = torch.randn(2, cin, r, r, generator=g)
    if relu:
        x = x.clamp_min(0)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, padding=1)
    xd, wd = x.to(dev), w.to(dev)
    for math in (1, 0):
        nt = 128 if r == 128 else 64
        p = ops.conv3x3_pack_weights(wd, nt=nt, math=math)
        o1, o2 = torch.empty(ref.shape, device=dev), torch.empty(ref.shape, device=dev)
        ops.conv3x3_forward(xd, p, None, o1, nt=nt, math=math)
        ops.conv3x3_forward(-xd, p, None, o2, nt=nt, math=math)
        e = (o1.double().cpu() - ref)
        s = (o1.double() + o2.double()).cpu() / 2
        print("%-24s %s  max|e|/max %.1e  rms(e)/rms %.1e  mean(e)/mean|ref| %+.2e  mean((o(x)+o(-x))/2)/mean|ref| %+.2e  rms of that %.1e" % (
            name, "bf16x3" if math else "tf32x3", e.abs().max() / ref.abs().max(), e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt(),
            e.mean() / ref.abs().mean(), s.mean() / ref.abs().mean(), s.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))
    ol = F.conv2d(xd, wd, None, padding=1)
    e = ol.double().cpu() - ref
    print("%-24s cudnn32 max|e|/max %.1e  rms(e)/rms %.1e  mean(e)/mean|ref| %+.2e" % (name, e.abs().max() / ref.abs().max(), e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt(), e.mean() / ref.abs().mean()))
