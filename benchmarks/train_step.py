"""FFWM train step (BASELINE config 3 on one GPU, config 4 under torchrun: batch 8 per GPU).

One step = one `optimize_parameters()` of the full model — flowNetF + flowNetB + netG forward,
guided filter, 8 facial-part crops, discriminator step, generator/flow step with VGG19 perceptual,
L1, illumination, LightCNN identity, adversarial and facial-part losses, three Adam updates, and
(N>1) the NCCL gradient all-reduce — on the SURVEY 8(d) cfg3 synthetic batch: 8 images of
128x128 per rank, titers=30000 (steady-state branch), random-init weights of the reference
architectures (no checkpoints or ImageNet weights are reachable offline).

Algorithmic work: 447.0 GFLOP per image (SURVEY 6 [probe]: 461.5 as the reference executes it, conv +
matmul, forward + backward, minus the 14.5 GFLOP of LightCNN weight gradients that the reference
computes and never uses — they are skipped here, which changes no loss and no applied gradient).  Convolutions run in strict fp32 (`cudnn.allow_tf32 = False`) unless
FFWM_BENCH_TF32=1, because the parity target of the path is 1e-4 relative.
"""
import os
import time

import torch

GFLOP_PER_IMAGE = 447.0      # 461.5 as the reference executes it, minus LightCNN's never-used weight gradients (14.5)
BATCH = 8


def make_batch(b, seed, pin=False):
    g = torch.Generator().manual_seed(seed)
    batch = {
        'img_S': torch.rand(b, 3, 128, 128, generator=g), 'img_F': torch.rand(b, 3, 128, 128, generator=g),
        'mask_F': (torch.rand(b, 1, 128, 128, generator=g) > 0.3).float(),
        'mask_S': (torch.rand(b, 1, 128, 128, generator=g) > 0.3).float(),
        'lm_F': torch.randint(20, 108, (b, 1000, 2), generator=g),
    }
    if pin:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    batch.update(titers=30000, epoch=0)
    return batch


class TrainStepWorkload:
    METRIC = "FFWM train-step images/sec (128x128)"
    UNIT = "images/s"
    DTYPE = "f32"
    STEPS, WARMUP = 20, 5
    E2E_STEPS = 10
    REF_STEPS, REF_WARMUP = 2, 1

    def __init__(self, device, rank=0, world=1):
        self.dev, self.rank, self.world = device, rank, world
        self.tf32 = os.environ.get("FFWM_BENCH_TF32", "0") == "1"

    def setup(self):
        torch.backends.cudnn.allow_tf32 = self.tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32
        torch.backends.cudnn.benchmark = True            # as the reference (models/base_model.py:38-39)
        from ffwm_b200.train_step import FFWMTrainer
        from ffwm_b200.parallel import Distributed
        torch.manual_seed(0)                              # identical weights on every rank
        dist = Distributed() if self.world > 1 else None
        self.use_graph = os.environ.get("FFWM_BENCH_GRAPH", "1") == "1"
        self.trainer = FFWMTrainer(self.dev, distributed=dist, graph=self.use_graph)
        self.dev_batch = {k: (v.to(self.dev) if torch.is_tensor(v) else v)
                          for k, v in make_batch(BATCH, 1000 + self.rank).items()}
        self.host_batch = make_batch(BATCH, 2000 + self.rank, pin=True)
        if self.use_graph:
            # FFWM_BENCH_SEGMENTED=1: the three-graph (data-parallel) layout on one GPU too (diagnostic)
            seg = True if os.environ.get("FFWM_BENCH_SEGMENTED", "0") == "1" else None
            self.trainer.enable_cuda_graph(self.dev_batch, segmented=seg)
            self.dev_batch = self.trainer._static        # replay in place, no copy
        if os.environ.get("FFWM_BENCH_PHASES", "0") == "1":
            self.trainer.phase_events = []

    def phase_report(self):
        """Mean GPU milliseconds per phase of the segmented step (FFWM_BENCH_PHASES=1): graph k, then the all-reduce after it."""
        ev = getattr(self.trainer, "phase_events", None)
        if not ev:
            return None
        torch.cuda.synchronize()
        tot, cnt = {}, {}
        for (name, e0, _), (_, e1, _) in zip(ev[:-1], ev[1:]):
            if name == "end":
                continue
            tot[name] = tot.get(name, 0.0) + e0.elapsed_time(e1)
            cnt[name] = cnt.get(name, 0) + 1
        return {k: round(tot[k] / cnt[k], 3) for k in tot}

    def step(self, timed):
        self.trainer.step(self.dev_batch)

    def step_e2e(self):
        """What a user of the reference does per iteration (train_ffwm.py:72-83): set_input from host
        tensors (H2D inside), optimize_parameters, get_current_losses (8 float() reads, D2H)."""
        self.trainer.step(self.host_batch)
        self.losses = self.trainer.get_current_losses()

    def e2e_bytes(self):
        h2d = sum(v.numel() * v.element_size() for v in self.host_batch.values() if torch.is_tensor(v))
        return h2d, 8 * 4

    def extra_launches(self):
        # kernels inside a replayed graph do not pass through the ctypes counter: count them per replay
        t = self.trainer
        return getattr(t, "graph_replays", 0) * getattr(t, "graph_kernel_nodes", 0)

    def units_per_step(self):
        return BATCH

    def config(self):
        return {"workload": "ffwm_model train step (BASELINE cfg3; cfg4 when n_gpus>1)", "batch_per_gpu": BATCH,
                "global_batch": BATCH * self.world, "image": "128x128", "titers": 30000,
                "nets": "netG(FFWM sn) + flowNetF + flowNetB + netD(MSDiscriminator) + LightCNN-29 + VGG19",
                "parallelism": "dp%d (NCCL grad all-reduce, per-rank BN)" % self.world,
                "conv_math": "tf32" if self.tf32 else "fp32 (cudnn.allow_tf32=False)",
                "launch": ("eager launches" if not getattr(self, "use_graph", True) else
                           "whole step replayed as one CUDA graph" if self.world == 1 else
                           "step replayed as three CUDA graphs with the two NCCL all-reduces issued eagerly between them"),
                "skipped": "LightCNN weight gradients (computed but never used by the reference)",
                "l2": "activations of one step (several GB) exceed L2; weights 488 MB", "weights": "random init"}

    def roofline(self, pk):
        return None      # filled by bench.py from the measured step time (see step_roofline)

    def step_roofline(self, pk, ms_per_step):
        """`roofline` of the train-step line: the step's dominant hand-written kernel — conv3x3_tc on netG's dres2 layers
        (195->195 @128x128, batch 8: 89.7 GFLOP per launch, the largest single-layer share of the step's FLOPs) — timed
        live with CUDA events right after the timed region, in the operand math its forward pass uses (3xTF32: three
        tf32 MMAs per product), against the cuBLAS TF32 rate measured in this same run.  The whole-step figure
        (algorithmic FLOPs / step time) is reported next to it."""
        import json
        from ffwm_b200 import _lib, ops
        step_tflops = GFLOP_PER_IMAGE * BATCH / 1e3 / (ms_per_step * 1e-3)
        out = {"kernel": "conv3x3_tc_kernel<128,128> forward, netG dres2 195->195 @128x128 x batch 8 (3xTF32 operand split)",
               "bound": "tensor", "unit": "TFLOP/s", "peak_source": "cuBLAS tf32 GEMM 8192^3 measured in this run (bench.py tensor_peaks)",
               "whole_step_TFLOP/s": round(step_tflops, 2),
               "whole_step_note": "447.0 algorithmic GFLOP/image (conv+matmul fwd+bwd) x 8 / measured step time; "
                                  "%.4f of the measured sustained bf16 rate" % (step_tflops / pk["bf16_tflops_sustained"])}
        try:
            cin = cout = 195
            x = torch.randn(BATCH, cin, 128, 128, device=self.dev)
            w = torch.randn(cout, cin, 3, 3, device=self.dev) / (cin * 9) ** 0.5
            o = torch.empty(BATCH, cout, 128, 128, device=self.dev)
            packed = ops.conv3x3_pack_weights(w, nt=128, math=_lib.MATH_TF32X3)
            for _ in range(3):
                ops.conv3x3_forward(x, packed, None, o, nt=128, math=_lib.MATH_TF32X3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                ops.conv3x3_forward(x, packed, None, o, nt=128, math=_lib.MATH_TF32X3)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            gflop = 2.0 * BATCH * 128 * 128 * cin * cout * 9 / 1e9
            from bench import tensor_peaks
            peak = tensor_peaks(self.dev.index or 0, iters=5).get("tf32_tflops")
            issued = 3 * gflop / ms
            out.update({"achieved": round(issued, 1), "achieved_note": "issued tensor math = 3 x algorithmic (%.1f useful TFLOP/s, %.4f ms per launch)" % (gflop / ms, ms),
                        "algorithmic_GFLOP_per_launch": round(gflop, 2), "peak": peak, "frac": round(issued / peak, 4) if peak else None})
            tp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
            out["traffic"] = json.load(open(tp)).get("conv3x3_tc_dres2") if os.path.exists(tp) else None
            out["algorithmic_bytes_per_launch"] = 4 * BATCH * 128 * 128 * (cin + cout) + 4 * cin * cout * 9
        except Exception as e:          # noqa: BLE001 - the bench line must survive
            out.update({"achieved": round(step_tflops, 2), "peak": pk["bf16_tflops_sustained"], "frac": step_tflops / pk["bf16_tflops_sustained"],
                        "traffic": None, "error": repr(e)})
        return out

    # the step's dominant hand-written kernel (conv3x3_tc: 55 % of the step's convolution FLOPs), on its three
    # heaviest layer shapes, batch = the step's batch
    KERNEL_SHAPES = (("conv3x3_tc fwd 195->195 @128x128 (netG dres2)", 195, 195, 128),
                     ("conv3x3_tc fwd 128->128 @128x128 (netG att2)", 128, 128, 128),
                     ("conv3x3_tc fwd 384->384 @32x32 (netG dres0)", 384, 384, 32))

    def kernel_table(self, pk):
        """Those kernels timed alone, live, with CUDA events on the launching stream (operands 100+ MB at 128x128:
        larger than L2 together with the output), in both operand maths: 3xTF32 (what the step's forward passes run)
        and 3xBF16 (its data gradients).  `TFLOP/s` counts the algorithmic 2*B*H*W*Cin*Cout*9 once; the kernel issues
        three MMAs per product, `tensor_issued_TFLOP/s`, compared with the measured dense rate of that MMA kind.
        Never fatal: a failure here is reported in the table instead of losing the bench line."""
        try:
            from ffwm_b200 import _lib, ops
            rows = {}
            tf32_peak = None
            for name, cin, cout, r in self.KERNEL_SHAPES:
                x = torch.randn(BATCH, cin, r, r, device=self.dev)
                w = torch.randn(cout, cin, 3, 3, device=self.dev) / (cin * 9) ** 0.5
                out = torch.empty(BATCH, cout, r, r, device=self.dev)
                nt = 128 if (r == 128 and cout > 64) else 64
                row = {"GFLOP": round(2.0 * BATCH * r * r * cin * cout * 9 / 1e9, 2), "nt": nt}
                for math, tag in ((_lib.MATH_BF16X3, "bf16x3"), (_lib.MATH_TF32X3, "tf32x3")):
                    packed = ops.conv3x3_pack_weights(w, nt=nt, math=math)
                    for _ in range(3):
                        ops.conv3x3_forward(x, packed, None, out, nt=nt, math=math)
                    iters = 10
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(iters):
                        ops.conv3x3_forward(x, packed, None, out, nt=nt, math=math)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / iters
                    tf = row["GFLOP"] / ms
                    row[tag] = {"ms": round(ms, 4), "TFLOP/s": round(tf, 1), "tensor_issued_TFLOP/s": round(3 * tf, 1)}
                    if math == _lib.MATH_BF16X3 and pk.get("bf16_tflops"):
                        row[tag]["frac_bf16_peak_issued"] = round(3 * tf / pk["bf16_tflops"], 4)
                    del packed
                rows[name] = row
                del x, w, out
            rows["note"] = ("timed alone after the step (burst clocks); three MMAs are issued per product for fp32-level accuracy; "
                            "bf16x3 fractions are against the measured dense bf16 burst rate (MEASURED_PEAKS.json), the TF32 rate "
                            "measured in this run is in measured_tensor_peaks")
            rows.update(self._hbm_kernel_rows(pk))
            return rows
        except Exception as e:          # noqa: BLE001 - the bench line must survive
            return {"error": repr(e)}

    def _hbm_kernel_rows(self, pk):
        """The HBM-bound hand-written kernels of the step that are not warp ops, on the largest map of the step (netG dres2:
        8 x 195 x 128 x 128 = 102 MB, two rotating copies: 409 MB of operands, larger than L2): batch norm + LeakyReLU +
        residual forward (algorithmic 3 reads + 1 write), backward (5 reads + 2 writes incl. the residual's gradient), and the
        bias-gradient channel sum (1 read), against the measured HBM copy bandwidth."""
        from ffwm_b200 import ops
        shape = (BATCH, 195, 128, 128)
        nbytes = 4 * BATCH * 195 * 128 * 128
        xs = [torch.randn(shape, device=self.dev) for _ in range(2)]
        gos = [torch.randn(shape, device=self.dev) for _ in range(2)]
        res, y, gx, gr = (torch.empty(shape, device=self.dev).normal_() for _ in range(4))
        c = shape[1]
        w, b = torch.rand(c, device=self.dev) + 0.5, torch.randn(c, device=self.dev)
        rm, rv, sm, si, gw, gb = (torch.zeros(c, device=self.dev) for _ in range(6))
        peak = pk.get("hbm_gbs")

        def timed(fn, iters=10):
            for i in range(3):
                fn(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(iters):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        cases = {
            "batch_norm+lrelu+residual fwd 8x195x128x128": (4, lambda i: ops.batch_norm_forward(xs[i % 2], res, w, b, rm, rv, 0.1, 1e-5, 0.2, y, sm, si)),
            "batch_norm+lrelu+residual bwd 8x195x128x128": (7, lambda i: ops.batch_norm_backward(xs[i % 2], gos[i % 2], y, w, b, sm, si, 0.2, gx, gr, gw, gb)),
            "channel_sum (bias gradient) 8x195x128x128": (1, lambda i: ops.channel_sum(gos[i % 2])),
        }
        rows = {}
        for name, (passes, fn) in cases.items():
            ms = timed(fn)
            gbps = passes * nbytes / ms / 1e6
            rows[name] = {"ms": round(ms, 4), "algorithmic_GB": round(passes * nbytes / 1e9, 3), "GB/s": round(gbps, 1)}
            if peak:
                rows[name]["frac_hbm"] = round(gbps / peak, 4)
        return rows

    # ------------------------------------------------------------------ same-run GPU library baseline
    def gpu_library_baseline(self, steps=10, warmup=3):
        """The SAME trainer, batch and CUDA-graph replay with every hand-written kernel switched off: cuDNN for all
        convolutions and batch norms, ATen `grid_sample` for the warps, the torch-op guided filter and max-feature-map — i.e. what
        the reference's modules execute on this GPU through PyTorch — in strict fp32 and with TF32 allowed (torch's
        default for cuDNN, outside the path's 1e-4 parity target).  Reported next to `value`; never fatal."""
        import gc
        import torch.nn.functional as F
        from ffwm_b200 import conv, external_function as EF, light_cnn, losses, norm, pool, spectral
        from ffwm_b200.train_step import FFWMTrainer

        def lib_warp(images, flow):
            return F.grid_sample(images, flow.permute(0, 2, 3, 1), mode='bilinear', padding_mode='zeros', align_corners=False)

        saved = (conv.ENABLED, EF.grid_warp, losses.grid_warp, EF.FUSED_GF, light_cnn.FUSED_MFM, norm.ENABLED, pool.ENABLED, spectral.FUSED_SN,
                 losses.FUSED_AFFINE, losses.FUSED_CORRMAX, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        rows = {}
        try:
            conv.ENABLED, EF.FUSED_GF, light_cnn.FUSED_MFM, norm.ENABLED, pool.ENABLED, spectral.FUSED_SN = False, False, False, False, False, False
            losses.FUSED_AFFINE = losses.FUSED_CORRMAX = False
            EF.grid_warp = losses.grid_warp = lib_warp
            for name, tf32 in (("cudnn_fp32", False), ("cudnn_tf32", True)):
                torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.manual_seed(0)
                tr = FFWMTrainer(self.dev, graph=True)
                batch = {k: (v.to(self.dev) if torch.is_tensor(v) else v) for k, v in make_batch(BATCH, 1000).items()}
                tr.enable_cuda_graph(batch)
                for _ in range(warmup):
                    tr.step(tr._static)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(steps):
                    tr.step(tr._static)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                rows[name] = {"ms_per_step": round(ms, 3), "images_per_s": round(BATCH / (ms * 1e-3), 2), "steps": steps,
                              "ffwm_b200_kernels_in_graph": tr.graph_kernel_nodes}
                del tr, batch
                gc.collect()
                torch.cuda.empty_cache()
            rows["note"] = ("same FFWMTrainer / batch / graph replay with conv.ENABLED=False (cuDNN convolutions), cuDNN batch norm + "
                            "ATen LeakyReLU / max pooling, torch-op spectral norm (batched per shape), F.grid_sample warps, torch-op guided filter and MFM (ffwm_b200_kernels_in_graph must be 0); "
                            "host-side restructurings (batched spectral norm / loss networks, torch's fused Adam) unchanged")
        except Exception as e:          # noqa: BLE001 - the bench line must survive
            rows["error"] = repr(e)
        finally:
            (conv.ENABLED, EF.grid_warp, losses.grid_warp, EF.FUSED_GF, light_cnn.FUSED_MFM, norm.ENABLED, pool.ENABLED, spectral.FUSED_SN,
             losses.FUSED_AFFINE, losses.FUSED_CORRMAX, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32) = saved
        return rows

    # ------------------------------------------------------------------ the reference's own CPU path
    @classmethod
    def _cpu_time(cls, steps, warmup):
        """The UNMODIFIED reference `FFWMModel(gpu_ids=[])` (baseline/_ref: byte-identical staged copy of the
        reference's models/ + lightcnn/, baseline/stage_ref.py) at the benchmark's batch 8 on all host cores:
        `set_train_input` + `optimize_parameters()` + `get_current_losses()` per step, as train_ffwm.py:72-83.
        Harness shims only (baseline/ref_harness.py).  None of this repo's modules or kernels are on this path."""
        from baseline import ref_harness as H
        torch.set_num_threads(os.cpu_count())
        model = H.reference_ffwm_model("cpu")
        batch = H.synthetic_batch(BATCH, 3000)
        for _ in range(warmup):
            model.set_train_input(batch)
            model.optimize_parameters()
        t0 = time.perf_counter()
        for _ in range(steps):
            model.set_train_input(batch)
            model.optimize_parameters()
            model.get_current_losses()
        dt = (time.perf_counter() - t0) / steps
        sample = ("the reference's unmodified FFWMModel(gpu_ids=[]).optimize_parameters() on the host CPU, batch %d, %d timed "
                  "step(s) after %d warm-up: its own networks and losses on PyTorch CPU (oneDNN) kernels, F.grid_sample warps"
                  % (BATCH, steps, warmup))
        return BATCH / dt, dt, sample

    @classmethod
    def cpu_baseline(cls):
        v, dt, sample = cls._cpu_time(steps=1, warmup=1)
        return {"value": v, "unit": cls.UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": sample, "s_per_step": dt}

    @classmethod
    def run_reference(cls, steps, warmup, n_gpus):
        v, dt, sample = cls._cpu_time(steps=max(1, steps), warmup=warmup)
        cfg = {"workload": "ffwm_model train step (BASELINE cfg3)", "batch_per_gpu": BATCH, "global_batch": BATCH,
               "image": "128x128", "titers": 30000,
               "nets": "netG(FFWM sn) + flowNetF + flowNetB + netD(MSDiscriminator) + LightCNN-29 + VGG19",
               "parallelism": "cpu, %d threads (the reference has no multi-GPU path; one CPU process regardless of --gpus)" % os.cpu_count(),
               "conv_math": "fp32 (oneDNN)", "launch": "eager PyTorch CPU", "skipped": "nothing (LightCNN weight gradients included, as the reference computes them)",
               "weights": "random init", "code": "unmodified reference (baseline/_ref), harness shims only", "sample": sample}
        return {"impl": "reference", "metric": cls.METRIC, "value": v, "unit": cls.UNIT, "n_gpus": n_gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": cls.DTYPE, "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": cls.UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": sample},
                "e2e": {"value": v, "unit": cls.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
