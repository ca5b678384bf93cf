"""3x3 convolution microbench: the hand-written tcgen05 kernel (3xTF32, fp32-level accuracy) vs cuDNN
in strict fp32 and in TF32, on the generator's dominant shapes, B = 8, 128x128.

    python -m benchmarks.conv [--out gpurun_out/conv.json]
    python -m benchmarks.conv --wgrad [--out gpurun_out/conv_wgrad.json]     (experimental weight-gradient kernel)
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


SHAPES = (("dres2 195->195 @128", 195, 195, 128), ("att2 128->128 @128", 128, 128, 128),
          ("e0 res 64->64 @128", 64, 64, 128), ("dres1 195->195 @64", 195, 195, 64),
          ("d2 195->256 @64", 195, 256, 64), ("vgg 128->128 @64", 128, 128, 64),
          ("dres0 384->384 @32", 384, 384, 32), ("vgg 256->256 @32", 256, 256, 32))


def wgrad_bench(out_path):
    """tcgen05 weight gradient (csrc/conv_gen_wgrad_tc.cu: split-K kernel + reduction) vs cuDNN's strict-fp32 and TF32
    weight gradients on the same shapes; errors against float64."""
    from ffwm_b200 import ops
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    rows = []

    def cudnn_wgrad(go, x, w):
        return torch.ops.aten.convolution_backward(go, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1,
                                                   [False, True, False])[1]

    for name, cin, cout, r in SHAPES:
        x = torch.randn(8, cin, r, r, device=dev)
        go = torch.randn(8, cout, r, r, device=dev)
        w = torch.zeros(cout, cin, 3, 3, device=dev)
        flop = 2.0 * 8 * r * r * cin * cout * 9
        gw2 = torch.empty_like(w)
        t_gen = timeit(lambda: ops.conv_wgrad(go, x, gw2, 1, 1))
        w64 = w.double().requires_grad_()
        F.conv2d(x.double(), w64, None, padding=1).backward(go.double())
        ref = w64.grad
        err_gen = float((gw2.double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = False
        t_fp32 = timeit(lambda: cudnn_wgrad(go, x, w))
        err_fp32 = float((cudnn_wgrad(go, x, w).double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = True
        t_tf32 = timeit(lambda: cudnn_wgrad(go, x, w))
        err_tf32 = float((cudnn_wgrad(go, x, w).double() - ref).abs().max() / ref.abs().max())
        rows.append(dict(shape=name, cin=cin, cout=cout, gflop=flop / 1e9, ms_cudnn_fp32=t_fp32, ms_tcgen05=t_gen,
                         tflops_tcgen05=flop / t_gen / 1e9, err_tcgen05=err_gen, ms_cudnn_tf32=t_tf32,
                         err_cudnn_fp32=err_fp32, err_cudnn_tf32=err_tf32))
        del ref, w64
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(rows, open(out_path, "w"), indent=1)
    print("%-22s %8s %8s %8s %8s %9s %9s %9s" % ("wgrad shape", "tc ms", "fp32 ms", "tf32 ms", "TF/s", "err tc", "err fp32", "err tf32"))
    for r in rows:
        print("%-22s %8.3f %8.3f %8.3f %8.1f %9.1e %9.1e %9.1e" % (
            r["shape"], r["ms_tcgen05"], r["ms_cudnn_fp32"], r["ms_cudnn_tf32"], r["tflops_tcgen05"],
            r["err_tcgen05"], r["err_cudnn_fp32"], r["err_cudnn_tf32"]))


def nt128_bench(out_path):
    """conv3x3_tc_kernel<128, 64> (validated) vs <128, 128> (experimental) on the W = 128 shapes with > 64 output channels."""
    from ffwm_b200 import ops
    dev = torch.device("cuda", 0)
    rows = []
    for name, cin, cout, r in SHAPES:
        if r != 128 or cout <= 64:
            continue
        x = torch.randn(8, cin, r, r, device=dev)
        w = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
        b = torch.randn(cout, device=dev)
        o64, o128 = torch.empty(8, cout, r, r, device=dev), torch.empty(8, cout, r, r, device=dev)
        p64, p128 = ops.conv3x3_pack_weights(w), ops.conv3x3_pack_weights(w, nt=128)
        t64 = timeit(lambda: ops.conv3x3_forward(x, p64, b, o64))
        t128 = timeit(lambda: ops.conv3x3_forward(x, p128, b, o128, nt=128))
        flop = 2.0 * 8 * r * r * cin * cout * 9
        rows.append(dict(shape=name, ms_nt64=t64, ms_nt128=t128, tflops_nt64=flop / t64 / 1e9, tflops_nt128=flop / t128 / 1e9,
                         max_abs_diff=float((o64 - o128).abs().max())))
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(rows, open(out_path, "w"), indent=1)
    for r in rows:
        print("%-22s nt64 %.3f ms (%.1f TF/s)   nt128 %.3f ms (%.1f TF/s)   max|diff| %.1e" % (
            r["shape"], r["ms_nt64"], r["tflops_nt64"], r["ms_nt128"], r["tflops_nt128"], r["max_abs_diff"]))


GEN_SHAPES = (  # name, b, cin, cout, h, k, stride, pad, transposed
    ("G stem 7x7 3->64 @128", 8, 3, 64, 128, 7, 1, 3, 0), ("G e1 4x4s2 64->128 @128", 8, 64, 128, 128, 4, 2, 1, 0),
    ("G e2 4x4s2 128->256 @64", 8, 128, 256, 64, 4, 2, 1, 0), ("G 1x1 195->195 @128", 8, 195, 195, 128, 1, 1, 0, 0),
    ("3x3 384->384 @16", 8, 384, 384, 16, 3, 1, 1, 0), ("F 3x3s2 64->128 @64", 8, 64, 128, 64, 3, 2, 1, 0),
    ("F 3x3s2 256->512 @16", 8, 256, 512, 16, 3, 2, 1, 0), ("F 3x3 512->512 @8", 8, 512, 512, 8, 3, 1, 1, 0),
    ("F 3x3 1024->1024 @2", 8, 1024, 1024, 2, 3, 1, 1, 0), ("F deconv 1024->512 @2", 8, 1024, 512, 2, 4, 2, 1, 1),
    ("F deconv 770->128 @8", 8, 770, 128, 8, 4, 2, 1, 1), ("F deconv 194->32 @32", 8, 194, 32, 32, 4, 2, 1, 1),
    ("L 5x5 1->96 @128", 8, 1, 96, 128, 5, 1, 2, 0), ("L 3x3 192->384 @16", 8, 192, 384, 16, 3, 1, 1, 0),
    ("VGG 3x3 512->512 @16", 8, 512, 512, 16, 3, 1, 1, 0), ("VGG 3x3 512->512 @8", 8, 512, 512, 8, 3, 1, 1, 0),
    ("D 3x3s2 3->64 @128", 8, 3, 64, 128, 3, 2, 1, 0))


def gen_bench(out_path):
    """csrc/conv_gen_tc.cu (incl. its weight packing) vs cuDNN strict fp32 and TF32 on the non-3x3-stride-1 shapes."""
    from ffwm_b200 import ops
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    rows = []
    for name, b, cin, cout, r, k, s, p, tr in GEN_SHAPES:
        x = torch.randn(b, cin, r, r, device=dev)
        if tr:
            w = torch.randn(cin, cout, k, k, device=dev) / (cin * k * k / s / s) ** 0.5
            lib = lambda: F.conv_transpose2d(x, w, None, stride=s, padding=p)
            ref = F.conv_transpose2d(x.double(), w.double(), None, stride=s, padding=p)
        else:
            w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
            lib = lambda: F.conv2d(x, w, None, stride=s, padding=p)
            ref = F.conv2d(x.double(), w.double(), None, stride=s, padding=p)
        out = torch.empty(ref.shape, device=dev)
        packed = ops.conv_pack_weights(w, bool(tr), s, p, bool(tr))
        flop = 2.0 * ref.numel() / cout * cout * cin * k * k / (s * s if tr else 1)
        t = timeit(lambda: ops.conv_forward(x, packed, None, out, k, k, s, p, bool(tr)))
        t_pack = timeit(lambda: ops.conv_pack_weights(w, bool(tr), s, p, bool(tr)))
        err = float((out.double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = False
        t32 = timeit(lib)
        e32 = float((lib().double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = True
        ttf = timeit(lib)
        rows.append(dict(shape=name, gflop=flop / 1e9, ms_tcgen05=t, ms_pack=t_pack, tflops=flop / t / 1e9, err=err,
                         ms_cudnn_fp32=t32, err_cudnn_fp32=e32, ms_cudnn_tf32=ttf))
        del ref
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(rows, open(out_path, "w"), indent=1)
    print("%-26s %7s | %8s %8s %7s %8s | %8s %8s | %8s" % ("shape", "GFLOP", "tc ms", "pack ms", "TF/s", "err", "cudnn32", "err", "cudnnTF"))
    for r in rows:
        print("%-26s %7.2f | %8.4f %8.4f %7.1f %8.1e | %8.4f %8.1e | %8.4f" % (
            r["shape"], r["gflop"], r["ms_tcgen05"], r["ms_pack"], r["tflops"], r["err"], r["ms_cudnn_fp32"], r["err_cudnn_fp32"],
            r["ms_cudnn_tf32"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--wgrad", action="store_true", help="measure the tcgen05 weight gradient instead")
    ap.add_argument("--nt128", action="store_true", help="A/B the experimental 128-channel CTA tile on the W = 128 shapes")
    ap.add_argument("--gen", action="store_true", help="the general kernel (conv_gen_tc.cu) on the other shapes of the path")
    args = ap.parse_args()
    if args.gen:
        return gen_bench(args.out or os.path.join(ROOT, "gpurun_out", "conv_gen.json"))
    if args.wgrad:
        return wgrad_bench(args.out or os.path.join(ROOT, "gpurun_out", "conv_wgrad.json"))
    if args.nt128:
        return nt128_bench(args.out or os.path.join(ROOT, "gpurun_out", "conv_nt128.json"))
    args.out = args.out or os.path.join(ROOT, "gpurun_out", "conv.json")
    from ffwm_b200 import _lib, ops
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    rows = []
    for name, cin, cout, r in SHAPES:
        x = torch.randn(8, cin, r, r, device=dev)
        w = torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5
        b = torch.randn(cout, device=dev)
        out = torch.empty(8, cout, r, r, device=dev)
        nt = 128 if (r == 128 and cout > 64) else 64            # what ffwm_b200/conv.py selects
        flop = 2.0 * 8 * r * r * cin * cout * 9
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
        row = dict(shape=name, cin=cin, cout=cout, gflop=flop / 1e9, nt=nt)
        for math, tag in ((1, "bf16x3"), (0, "tf32x3")):
            packed = ops.conv3x3_pack_weights(w, nt=nt, math=math)
            t = timeit(lambda: ops.conv3x3_forward(x, packed, b, out, nt=nt, math=math))
            row["ms_" + tag] = t
            row["tflops_" + tag] = flop / t / 1e9
            row["err_" + tag] = float((out.double() - ref).abs().max() / ref.abs().max())
            row["ms_pack_" + tag] = timeit(lambda: ops.conv3x3_pack_weights(w, nt=nt, math=math))
        torch.backends.cudnn.allow_tf32 = False
        row["ms_cudnn_fp32"] = timeit(lambda: F.conv2d(x, w, b, padding=1))
        row["err_cudnn_fp32"] = float((F.conv2d(x, w, b, padding=1).double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = True
        row["ms_cudnn_tf32"] = timeit(lambda: F.conv2d(x, w, b, padding=1))
        row["err_cudnn_tf32"] = float((F.conv2d(x, w, b, padding=1).double() - ref).abs().max() / ref.abs().max())
        rows.append(row)
        del ref
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    print("%-22s %3s | %8s %7s %8s | %8s %7s %8s | %8s %8s | %8s %8s" % (
        "shape", "nt", "bf16x3", "TF/s", "err", "tf32x3", "TF/s", "err", "cudnn32", "err", "cudnnTF", "err"))
    for r in rows:
        print("%-22s %3d | %8.3f %7.1f %8.1e | %8.3f %7.1f %8.1e | %8.3f %8.1e | %8.3f %8.1e" % (
            r["shape"], r["nt"], r["ms_bf16x3"], r["tflops_bf16x3"], r["err_bf16x3"], r["ms_tf32x3"], r["tflops_tf32x3"],
            r["err_tf32x3"], r["ms_cudnn_fp32"], r["err_cudnn_fp32"], r["ms_cudnn_tf32"], r["err_cudnn_tf32"]))


if __name__ == "__main__":
    main()
