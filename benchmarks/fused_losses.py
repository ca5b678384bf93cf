"""SURVEY 8f rows f1 / f3 at the sizes of FlowNet pre-training (BASELINE config 2: batch 6): the fused kernels against
the reference's chains of library ops / validated kernels on the same inputs, CUDA events, plus roofline figures.

    python -m benchmarks.fused_losses [--out gpurun_out/fused_losses.json]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fused_losses.json"))
    args = ap.parse_args()
    from ffwm_b200 import losses, ops
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    rows = []
    g = torch.Generator().manual_seed(0)
    # ---- f1: correlation column-max at the three VGG stages FlowNet pre-training uses (models/flownet_model.py:67)
    for name, c, hw in (("relu1_1", 64, 128), ("relu2_1", 128, 64), ("relu3_1", 256, 32)):
        b = 6
        src = torch.randn(b, c, hw, hw, generator=g).clamp_min(0).to(dev)
        tgt = torch.randn(b, c, hw, hw, generator=g).clamp_min(0).to(dev)
        t_fused = timeit(lambda: ops.corr_max(src, tgt, 1e-8))

        def chain():       # the reference's formulation (models/losses.py:347-353), row-blocked so that it fits at all
            sa = src.view(b, c, -1).transpose(1, 2)
            ta = tgt.view(b, c, -1)
            sn = sa / (sa.norm(dim=2, keepdim=True) + 1e-8)
            tn = ta / (ta.norm(dim=1, keepdim=True) + 1e-8)
            return losses.PerceptualCorrectness._column_max(sn, tn)
        t_chain = timeit(chain, iters=3)
        err = float((ops.corr_max(src, tgt, 1e-8) - chain()).abs().max())
        n2 = hw * hw
        flop = 2.0 * b * n2 * n2 * c
        tf = flop / t_fused / 1e9
        rows.append({"op": "corr_max " + name, "shape": [b, c, hw, hw], "ms_fused": t_fused, "ms_reference_chain": t_chain,
                     "speedup": t_chain / t_fused, "GFLOP": flop / 1e9, "TFLOP/s": tf, "tensor_issued_TFLOP/s": 3 * tf,
                     "frac_bf16_peak_issued": 3 * tf / peaks["bf16_tflops"] if peaks else None,
                     "intermediate_avoided_GB": 4.0 * b * n2 * n2 / 1e9, "max_abs_diff_vs_chain": err})
    # ---- f3: affine regularisation at the three flow scales (kz 3 / 5 / 7 on 32 / 64 / 128 grids)
    for kz, s in ((3, 32), (5, 64), (7, 128)):
        reg = losses.AffineRegularizationLoss(kz)
        lin = torch.linspace(-1, 1, s)
        gy, gx = torch.meshgrid(lin, lin, indexing="ij")
        flow = (torch.stack((gx, gy), 0).unsqueeze(0).repeat(6, 1, 1, 1) + 0.05 * torch.randn(6, 2, s, s, generator=g)).to(dev)

        def fb(fused):
            losses.FUSED_AFFINE = fused
            f = flow.clone().requires_grad_(True)
            reg(f).backward()
            losses.FUSED_AFFINE = True
            return f.grad
        t_fused = timeit(lambda: fb(True))
        t_chain = timeit(lambda: fb(False))
        diff = float((fb(True) - fb(False)).abs().max() / fb(False).abs().max())
        hp = s - kz + 1
        alg_bytes = 2 * 4.0 * 6 * (s * s + hp * hp) * 2          # fwd + bwd, two planes: read the grid, write/read the window map
        rows.append({"op": "affine_reg kz%d" % kz, "grid": [6, 2, s, s], "ms_fused_fwd_bwd": t_fused, "ms_reference_chain_fwd_bwd": t_chain,
                     "speedup": t_chain / t_fused, "algorithmic_MB": alg_bytes / 1e6, "intermediates_avoided_MB": 2 * 2 * 4.0 * 6 * (kz * hp) ** 2 / 1e6,
                     "rel_diff_grad_vs_chain": diff})
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    for r in rows:
        print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()})


if __name__ == "__main__":
    main()
