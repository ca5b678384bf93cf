"""Generates tests/golden/ref_cuda_*.npz by running the REFERENCE's own CUDA
extensions (oracle/_ref, built by oracle/build_ref.py from /root/reference)
on a B200:

    gpurun -- python tests/golden/make_golden_gpu.py     # writes gpurun_out/golden/

then copy gpurun_out/golden/*.npz into tests/golden/.  The CPU test-suite
checks the C oracle against these vectors, which pins the oracle to outputs
of the reference itself (resample2d has no test in the reference).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import cases  # noqa: E402


def load(name):
    p = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    spec = importlib.util.spec_from_file_location(name, p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def sub(a, big):
    a = a.detach().cpu().numpy()
    return a.reshape(-1)[::cases.SUBSAMPLE].copy() if big else a


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rs, be, lar = load("resample2d_cuda"), load("block_extractor_cuda"), load("local_attn_reshape_cuda")
    dev = "cuda:0"
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        store = {}
        for name in cases.RESAMPLE2D_CASES:
            c = cases.resample2d_case(name)
            in1 = torch.tensor(c["in1"], dtype=dt, device=dev)
            in2 = torch.tensor(c["in2"], dtype=dt, device=dev)
            gout = torch.tensor(c["gout"], dtype=dt, device=dev)
            out = torch.zeros_like(gout)
            rs.forward(in1, in2, out, c["ks"], c["dil"])
            g1, g2 = torch.zeros_like(in1), torch.zeros_like(in2)
            rs.backward(in1, in2, gout, g1, g2, c["ks"], c["dil"])
            big = name.startswith("cfg1")
            store["resample2d/%s/out" % name] = sub(out, big)
            store["resample2d/%s/gin1" % name] = sub(g1, big)
            store["resample2d/%s/gin2" % name] = sub(g2, big)
        for name in cases.BLOCK_EXTRACTOR_CASES:
            c = cases.block_extractor_case(name)
            src = torch.tensor(c["src"], dtype=dt, device=dev)
            flow = torch.tensor(c["flow"], dtype=dt, device=dev)
            gout = torch.tensor(c["gout"], dtype=dt, device=dev)
            out = torch.zeros_like(gout)
            be.forward(src, flow, out, c["k"])
            gs, gf = torch.zeros_like(src), torch.zeros_like(flow)
            be.backward(src, flow, gout, gs, gf, c["k"])
            store["block_extractor/%s/out" % name] = sub(out, False)
            store["block_extractor/%s/gsrc" % name] = sub(gs, False)
            store["block_extractor/%s/gflow" % name] = sub(gf, False)
        for name in cases.LOCAL_ATTN_RESHAPE_CASES:
            c = cases.local_attn_reshape_case(name)
            x = torch.tensor(c["x"], dtype=dt, device=dev)
            gout = torch.tensor(c["gout"], dtype=dt, device=dev)
            out = torch.zeros_like(gout)
            lar.forward(x, out, c["k"])
            gi = torch.zeros_like(x)
            lar.backward(x, gout, gi, c["k"])
            store["local_attn_reshape/%s/out" % name] = sub(out, False)
            store["local_attn_reshape/%s/gin" % name] = sub(gi, False)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, "ref_cuda_%s.npz" % tag), **store)
        print("wrote", tag, len(store), "arrays")


if __name__ == "__main__":
    main()
