"""Microbenchmark of csrc/batch_norm.cu against torch's batch norm + LeakyReLU (cuDNN kernels) on the train step's shapes.
Inputs rotate over enough copies to exceed the 126 MB L2; times are CUDA-event averages.  Algorithmic bytes: forward
2 reads + 1 write of the map, backward 4 reads + 1 write (the second read of x may hit L2)."""
import json
import sys

import torch
import torch.nn.functional as F

SHAPES = [(8, 195, 128, 128), (8, 64, 128, 128), (8, 128, 64, 64), (8, 16, 128, 128), (8, 387, 32, 32), (8, 256, 16, 16), (8, 128, 4, 4)]


def timed(fn, n_iter):
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_iter):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n_iter


def main():
    from ffwm_b200 import ops
    dev = "cuda:0"
    rows = []
    for shape in SHAPES:
        n, c, h, w = shape
        nbytes = n * c * h * w * 4
        copies = max(2, int(300e6 // nbytes) + 1)
        copies = min(copies, 64)
        xs = [torch.randn(shape, device=dev) for _ in range(copies)]
        gos = [torch.randn(shape, device=dev) for _ in range(copies)]
        y, gx = torch.empty(shape, device=dev), torch.empty(shape, device=dev)
        wt, bs = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        sm, si, gw, gb = (torch.empty(c, device=dev) for _ in range(4))
        iters = 20

        def mine_fwd(i):
            ops.batch_norm_forward(xs[i % copies], None, wt, bs, rm, rv, 0.1, 1e-5, 0.2, y, sm, si)

        def mine_bwd(i):
            ops.batch_norm_backward(xs[i % copies], gos[i % copies], None, wt, bs, sm, si, 0.2, gx, None, gw, gb)

        def lib_fwd(i):
            F.leaky_relu(F.batch_norm(xs[i % copies], rm, rv, wt, bs, True, 0.1, 1e-5), 0.2, inplace=True)

        t_mf, t_mb, t_lf = timed(mine_fwd, iters), timed(mine_bwd, iters), timed(lib_fwd, iters)
        xg = [x.clone().requires_grad_() for x in xs[:2]]
        wg, bg = wt.clone().requires_grad_(), bs.clone().requires_grad_()

        def lib_fb(i):
            out = F.leaky_relu(F.batch_norm(xg[i % 2], rm, rv, wg, bg, True, 0.1, 1e-5), 0.2, inplace=True)
            out.backward(gos[i % copies])

        t_lfb = timed(lib_fb, iters)
        rows.append({"shape": shape, "MB": round(nbytes / 1e6, 1), "fwd_us": round(t_mf * 1e3, 1), "fwd_GBps": round(3 * nbytes / t_mf / 1e6),
                     "bwd_us": round(t_mb * 1e3, 1), "bwd_GBps": round(5 * nbytes / t_mb / 1e6),
                     "torch_fwd_us": round(t_lf * 1e3, 1), "torch_fwd_bwd_us": round(t_lfb * 1e3, 1),
                     "speedup_fwd": round(t_lf / t_mf, 2), "speedup_fwd_bwd": round(t_lfb / (t_mf + t_mb), 2)})
        print(json.dumps(rows[-1]), flush=True)
    return rows


if __name__ == "__main__":
    sys.path.insert(0, __file__.rsplit("/", 2)[0])
    main()
