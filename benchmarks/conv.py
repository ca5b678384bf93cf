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
    """tcgen05 weight gradient (csrc/conv3x3_wgrad_tc.cu, incl. the zero fill of its target) vs cuDNN's strict-fp32
    and TF32 weight gradients on the same shapes; errors against float64."""
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
        gw = torch.zeros_like(w)
        flop = 2.0 * 8 * r * r * cin * cout * 9

        def mine():
            gw.zero_()
            ops.conv3x3_wgrad(x, go, gw)

        t_mine = timeit(mine)
        w64 = w.double().requires_grad_()
        F.conv2d(x.double(), w64, None, padding=1).backward(go.double())
        ref = w64.grad
        err_mine = float((gw.double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = False
        t_fp32 = timeit(lambda: cudnn_wgrad(go, x, w))
        err_fp32 = float((cudnn_wgrad(go, x, w).double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = True
        t_tf32 = timeit(lambda: cudnn_wgrad(go, x, w))
        err_tf32 = float((cudnn_wgrad(go, x, w).double() - ref).abs().max() / ref.abs().max())
        rows.append(dict(shape=name, cin=cin, cout=cout, gflop=flop / 1e9, ms_tcgen05=t_mine, ms_cudnn_fp32=t_fp32,
                         ms_cudnn_tf32=t_tf32, tflops_tcgen05=flop / t_mine / 1e9, err_tcgen05=err_mine,
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--wgrad", action="store_true", help="measure the experimental tcgen05 weight gradient instead")
    ap.add_argument("--nt128", action="store_true", help="A/B the experimental 128-channel CTA tile on the W = 128 shapes")
    args = ap.parse_args()
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
            old = _lib.set_option("CONV_MATH", math)
            packed = ops.conv3x3_pack_weights(w, nt=nt)
            t = timeit(lambda: ops.conv3x3_forward(x, packed, b, out, nt=nt))
            row["ms_" + tag] = t
            row["tflops_" + tag] = flop / t / 1e9
            row["err_" + tag] = float((out.double() - ref).abs().max() / ref.abs().max())
            row["ms_pack_" + tag] = timeit(lambda: ops.conv3x3_pack_weights(w, nt=nt))
            _lib.set_option("CONV_MATH", old)
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
