#!/usr/bin/env python
"""A/B of kernel generations at the cfg5 primary point (B=64, C=128, R=128; block_extractor B=16):
each op timed with CUDA events under a set of environment switches, and the results of the variants
compared with each other (max abs error / max |ref|).  Measurement harness, prints a table and
writes gpurun_out/roll_ab.json.

    python scripts/roll_ab.py [--iters 10] [--C 128] [--R 128]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchmarks.warp import alg_bytes, make_inputs  # noqa: E402


class env:
    def __init__(self, kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


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


def rel(a, b):
    return float((a - b).abs().max()) / max(1e-30, float(b.abs().max()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--C", type=int, default=128)
    ap.add_argument("--R", type=int, default=128)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "roll_ab.json"))
    a = ap.parse_args()
    from ffwm_b200 import ops
    dev = torch.device("cuda", 0)
    c, r = a.C, a.R
    B = max(8, -(-(512 << 20) // (4 * c * r * r)))
    shapes = dict(B=B, Bb=max(8, B // 4), B2=8, C=c, R=r)
    t = make_inputs(shapes, dev, seed=1234)
    peak = 6548.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    rs = dict(B=B, C=c, H=r, W=r, Hi=r, Wi=r)
    be = dict(B=shapes["Bb"], C=c, Hs=r, Ws=r, Hf=r, Wf=r, k=3)
    disp2 = t["disp"].clone()
    variants = [("new", {}), ("old", {"FFWM_DISABLE_ROLL": "1", "FFWM_SCATTER_TILED": "1", "FFWM_DISABLE_QUAD": "1"})]
    if os.environ.get("AB_DIRECT"):
        variants.append(("direct", {"FFWM_DISABLE_TILED": "1"}))
    e = torch.empty_like
    results = {}

    def run(name, nbytes, make_fn, outs, pre=None, only=None):
        ref = None
        for vn, kv in variants:
            if only and vn not in only:
                continue
            with env(kv):
                fn = make_fn()
                ms = timeit(fn, a.iters, pre)
                got = [o.clone() for o in outs()]
            err = None
            if ref is None:
                ref = got
            else:
                err = max(rel(x, y) for x, y in zip(got, ref))
            gbs = nbytes / ms / 1e6
            results.setdefault(name, {})[vn] = {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / peak, 4),
                                               "err_vs_new": err}
            print(f"{name:28s} {vn:8s} {ms:8.3f} ms {gbs:8.1f} GB/s {gbs / peak:6.3f}  err_vs_new={err}", flush=True)

    out = e(t["feat"])
    g1, g2 = e(t["feat"]), e(t["disp"])
    for ks in (4, 2):
        run(f"resample2d_ks{ks}_fwd", alg_bytes("resample2d_fwd", **rs),
            lambda: (lambda: ops.resample2d_forward(t["feat"], t["disp"], out, ks, 1)), lambda: [out])
        run(f"resample2d_ks{ks}_gflow", alg_bytes("resample2d_fwd", **rs),
            lambda: (lambda: ops.resample2d_backward(t["feat"], t["disp"], t["gout"], None, g2, ks, 1)), lambda: [g2])
        run(f"resample2d_ks{ks}_gin1", alg_bytes("resample2d_fwd", **rs),
            lambda: (lambda: ops.resample2d_backward(t["feat"], t["disp"], t["gout"], g1, None, ks, 1)), lambda: [g1],
            pre=lambda: g1.zero_())
        run(f"resample2d_ks{ks}_bwd", alg_bytes("resample2d_bwd", **rs),
            lambda: (lambda: ops.resample2d_backward(t["feat"], t["disp"], t["gout"], g1, g2, ks, 1)), lambda: [g1, g2],
            pre=lambda: g1.zero_())
    gf = e(t["grid"])
    run("grid_warp_fwd", alg_bytes("grid_warp_fwd", **rs),
        lambda: (lambda: ops.grid_warp_forward(t["feat"], t["grid"], out)), lambda: [out])
    run("grid_warp_gflow", alg_bytes("grid_warp_fwd", **rs),
        lambda: (lambda: ops.grid_warp_backward(t["feat"], t["grid"], t["gout"], None, gf)), lambda: [gf])
    run("grid_warp_gimg", alg_bytes("grid_warp_fwd", **rs),
        lambda: (lambda: ops.grid_warp_backward(t["feat"], t["grid"], t["gout"], g1, None)), lambda: [g1],
        pre=lambda: g1.zero_())
    run("grid_warp_bwd", alg_bytes("grid_warp_bwd", **rs),
        lambda: (lambda: ops.grid_warp_backward(t["feat"], t["grid"], t["gout"], g1, gf)), lambda: [g1, gf],
        pre=lambda: g1.zero_())
    del out, g1, g2
    ob = e(t["be_gout"])
    gs, gfl = e(t["be_src"]), e(t["be_flow"])
    for mode, flow in (("rand1.8", t["be_flow"]), ("randn2", torch.randn_like(t["be_flow"]) * 2)):
        run(f"block_extractor_fwd[{mode}]", alg_bytes("block_extractor_fwd", **be),
            lambda: (lambda: ops.block_extractor_forward(t["be_src"], flow, ob, 3)), lambda: [ob])
        run(f"block_extractor_bwd[{mode}]", alg_bytes("block_extractor_bwd", **be),
            lambda: (lambda: ops.block_extractor_backward(t["be_src"], flow, t["be_gout"], gs, gfl, 3)), lambda: [gs, gfl],
            pre=lambda: gs.zero_())
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"C": c, "R": r, "B": B, "peak": peak, "results": results}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
