"""cfg5 sweep: every warp kernel at R in {64,128,256} x C in {64,128,256} (+ the model-shape points),
ours through the C ABI and — when oracle/_ref is present — the reference's own CUDA kernels rebuilt
for sm_100a, both timed with CUDA events on working sets far larger than L2.

    python -m benchmarks.sweep [--iters 10] [--out gpurun_out/sweep.json] [--quick]

Measurement harness only; prints one table and writes JSON.
"""
import argparse
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks.warp import alg_bytes, _identity_grid  # noqa: E402


def load_ref():
    d = os.path.join(ROOT, "oracle", "_ref")
    mods = {}
    for n in ("resample2d_cuda", "block_extractor_cuda", "local_attn_reshape_cuda"):
        p = os.path.join(d, n + ".so")
        if not os.path.exists(p):
            return None
        spec = importlib.util.spec_from_file_location(n, p)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[n] = m
    return mods


def timeit(fn, iters, pre=None):
    for _ in range(2):
        if pre:
            pre()
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        if pre:
            pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--target-mib", type=int, default=512)
    args = ap.parse_args()
    from ffwm_b200 import ops
    ref = load_ref()
    dev = torch.device("cuda", 0)
    peak = 6548.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    points = [(c, r, None) for r in (64, 128, 256) for c in (64, 128, 256)]
    if args.quick:
        points = [(128, 128, None)]
    # model-shape points: tiny, L2 resident — reported for latency, not roofline
    points += [(128, 32, 8), (64, 64, 8), (64, 128, 8)]
    rows = []
    torch.manual_seed(0)
    for c, r, bfix in points:
        B = bfix or max(8, -(-(args.target_mib << 20) // (4 * c * r * r)))
        Bb = bfix or max(2, B // 4)
        feat = torch.rand(B, c, r, r, device=dev) * 2 - 1
        gout = torch.randn(B, c, r, r, device=dev)
        out, g1 = torch.empty_like(feat), torch.empty_like(feat)
        grid = (_identity_grid(B, r).to(dev) + torch.randn(B, 2, r, r, device=dev) * (2 * 2.0 / r)).contiguous()
        gfl = torch.empty_like(grid)
        for ks in (2, 4):
            disp = torch.cat([torch.randn(B, 2, r, r, device=dev) * 2, torch.full((B, 1, r, r), 2.0, device=dev)], 1)
            g2 = torch.empty_like(disp)
            sh = dict(B=B, C=c, H=r, W=r, Hi=r, Wi=r)
            res = {"resample2d_ks%d_fwd" % ks: (alg_bytes("resample2d_fwd", **sh),
                                                timeit(lambda: ops.resample2d_forward(feat, disp, out, ks, 1), args.iters),
                                                timeit(lambda: ref["resample2d_cuda"].forward(feat, disp, out, ks, 1), args.iters, pre=out.zero_) if ref else None),
                   "resample2d_ks%d_bwd" % ks: (alg_bytes("resample2d_bwd", **sh),
                                                timeit(lambda: (g1.zero_(), ops.resample2d_backward(feat, disp, gout, g1, g2, ks, 1)), args.iters),
                                                timeit(lambda: (g1.zero_(), g2.zero_(), ref["resample2d_cuda"].backward(feat, disp, gout, g1, g2, ks, 1)), args.iters) if ref else None)}
            for k, v in res.items():
                rows.append(dict(op=k, C=c, R=r, B=B, alg_bytes=v[0], ms=v[1], ref_ms=v[2]))
            del disp, g2
        sh = dict(B=B, C=c, H=r, W=r)
        rows.append(dict(op="grid_warp_fwd", C=c, R=r, B=B, alg_bytes=alg_bytes("grid_warp_fwd", **sh),
                         ms=timeit(lambda: ops.grid_warp_forward(feat, grid, out), args.iters),
                         ref_ms=timeit(lambda: torch.nn.functional.grid_sample(feat, grid.permute(0, 2, 3, 1), align_corners=False), args.iters),
                         ref_kind="torch.grid_sample"))
        feat_g = feat.clone().requires_grad_(True)
        grid_g = grid.clone().requires_grad_(True)

        def torch_bwd():
            o = torch.nn.functional.grid_sample(feat_g, grid_g.permute(0, 2, 3, 1), align_corners=False)
            feat_g.grad = grid_g.grad = None
            o.backward(gout)
        t_fb = timeit(torch_bwd, args.iters)
        rows.append(dict(op="grid_warp_bwd", C=c, R=r, B=B, alg_bytes=alg_bytes("grid_warp_bwd", **sh),
                         ms=timeit(lambda: (g1.zero_(), ops.grid_warp_backward(feat, grid, gout, g1, gfl)), args.iters),
                         ref_ms=t_fb - rows[-1]["ref_ms"], ref_kind="torch.grid_sample fwd+bwd minus fwd"))
        del feat_g, grid_g, feat, gout, out, g1
        # block extractor, k=3
        k = 3
        src = torch.rand(Bb, c, r, r, device=dev)
        flow = torch.rand(Bb, 2, r, r, device=dev) * 1.8
        bo = torch.empty(Bb, c, k * r, k * r, device=dev)
        bg = torch.randn(Bb, c, k * r, k * r, device=dev)
        gs, gf = torch.empty_like(src), torch.empty_like(flow)
        sh = dict(B=Bb, C=c, Hs=r, Ws=r, Hf=r, Wf=r, k=k)
        rows.append(dict(op="block_extractor_fwd", C=c, R=r, B=Bb, alg_bytes=alg_bytes("block_extractor_fwd", **sh),
                         ms=timeit(lambda: ops.block_extractor_forward(src, flow, bo, k), args.iters),
                         ref_ms=timeit(lambda: ref["block_extractor_cuda"].forward(src, flow, bo, k), args.iters, pre=bo.zero_) if ref else None))
        rows.append(dict(op="block_extractor_bwd", C=c, R=r, B=Bb, alg_bytes=alg_bytes("block_extractor_bwd", **sh),
                         ms=timeit(lambda: (gs.zero_(), ops.block_extractor_backward(src, flow, bg, gs, gf, k)), args.iters),
                         ref_ms=timeit(lambda: (gs.zero_(), gf.zero_(), ref["block_extractor_cuda"].backward(src, flow, bg, gs, gf, k)), args.iters) if ref else None))
        del src, flow, bo, bg, gs, gf
        torch.cuda.empty_cache()
    # local_attn_reshape: k in 3,5,7 at ~512 MiB
    for k, h in ((3, 128), (5, 60), (7, 122)):
        B = max(6, (args.target_mib << 20) // (4 * k * k * h * h)) if not args.quick else 64
        x = torch.rand(B, k * k, h, h, device=dev)
        o = torch.empty(B, 1, k * h, k * h, device=dev)
        gi = torch.empty_like(x)
        nb = alg_bytes("local_attn_reshape_fwd", B=B, k=k, H=h, W=h)
        rows.append(dict(op="local_attn_reshape_k%d_fwd" % k, C=k * k, R=h, B=B, alg_bytes=nb,
                         ms=timeit(lambda: ops.local_attn_reshape_forward(x, o, k), args.iters),
                         ref_ms=timeit(lambda: ref["local_attn_reshape_cuda"].forward(x, o, k), args.iters, pre=o.zero_) if ref else None))
        rows.append(dict(op="local_attn_reshape_k%d_bwd" % k, C=k * k, R=h, B=B, alg_bytes=nb,
                         ms=timeit(lambda: ops.local_attn_reshape_backward(o, gi, k), args.iters),
                         ref_ms=timeit(lambda: (gi.zero_(), ref["local_attn_reshape_cuda"].backward(x, o, gi, k)), args.iters) if ref else None))
        del x, o, gi
        torch.cuda.empty_cache()
    for row in rows:
        row["GBps"] = row["alg_bytes"] / 1e9 / (row["ms"] * 1e-3)
        row["frac_hbm"] = row["GBps"] / peak
        row["speedup_vs_ref"] = (row["ref_ms"] / row["ms"]) if row.get("ref_ms") else None
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peak_hbm_gbs": peak, "iters": args.iters, "rows": rows,
               "note": "bwd timings include the zero fill of the scatter target(s); ref = reference .cu rebuilt for sm_100a "
                       "(oracle/_ref) including the zero fills its callers perform"}, open(args.out, "w"), indent=1)
    print("%-28s %4s %4s %5s %9s %9s %7s %9s %7s" % ("op", "C", "R", "B", "ms", "GB/s", "frac", "ref_ms", "x_ref"))
    for r_ in rows:
        print("%-28s %4d %4d %5d %9.3f %9.1f %7.3f %9s %7s" % (
            r_["op"], r_["C"], r_["R"], r_["B"], r_["ms"], r_["GBps"], r_["frac_hbm"],
            "%.3f" % r_["ref_ms"] if r_.get("ref_ms") else "-",
            "%.2f" % r_["speedup_vs_ref"] if r_["speedup_vs_ref"] else "-"))


if __name__ == "__main__":
    main()
