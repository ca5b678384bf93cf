"""FlowNetF forward+backward (BASELINE config 2): batch 6, 128x128 synthetic pairs, one GPU.

One step = forward of `FlowNet(64)` in train mode + backward of sum(out_i * cotangent_i) over the
three flow scales (13.05 GFLOP per image).  Not the default bench line; run with
`python bench.py --workload flownet`.
"""
import os
import time

import torch

BATCH = 6
GFLOP_PER_IMAGE = 13.05


class FlowNetWorkload:
    METRIC = "FlowNetF forward+backward images/sec (128x128, batch 6)"
    UNIT = "images/s"
    DTYPE = "f32"
    STEPS, WARMUP = 30, 5
    E2E_STEPS = 10
    REF_STEPS, REF_WARMUP = 2, 1

    def __init__(self, device, rank=0, world=1):
        self.dev, self.rank, self.world = device, rank, world

    @staticmethod
    def _build_inputs():
        g = torch.Generator().manual_seed(1)
        x = torch.rand(BATCH, 3, 128, 128, generator=g)
        cots = [torch.randn(BATCH, 2, s, s, generator=g) for s in (128, 64, 32)]
        return None, x, cots

    def _build(self, dev):
        from ffwm_b200.base_networks import FlowNet
        torch.manual_seed(0)
        net = FlowNet(64).to(dev).train()
        _, x, cots = self._build_inputs()
        return net, x, cots

    def setup(self):
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        self.net, x, cots = self._build(self.dev)
        self.x, self.cots = x.to(self.dev), [c.to(self.dev) for c in cots]
        self.host_x = x.pin_memory()
        # the net is launch-bound (38 convolutions + 25 BatchNorms on maps down to 2x2, ~1500 launches per step):
        # forward + backward are captured once and replayed as one CUDA graph (FFWM_BENCH_GRAPH=0: eager launches)
        self.use_graph = os.environ.get("FFWM_BENCH_GRAPH", "1") == "1"
        self.graph = None
        if self.use_graph:
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._fb(self.x, set_to_none=False)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            from ffwm_b200 import _lib
            n0 = _lib.kernel_launches()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = self._fb(self.x, set_to_none=False)
            self.graph_kernel_nodes = _lib.kernel_launches() - n0
            self.graph_replays = 0

    def _fb(self, x, set_to_none=True):
        self.net.zero_grad(set_to_none=set_to_none)
        outs = self.net(x)
        loss = sum((o * c).sum() for o, c in zip(outs, self.cots))
        loss.backward()
        return loss

    def step(self, timed):
        if self.graph is not None:
            self.graph.replay()
            self.graph_replays += 1
        else:
            self._fb(self.x)

    def step_e2e(self):
        if self.graph is not None:
            self.x.copy_(self.host_x, non_blocking=True)
            self.graph.replay()
            self.graph_replays += 1
            self.loss_value = float(self.loss.detach())
        else:
            self.loss_value = float(self._fb(self.host_x.to(self.dev, non_blocking=True)).detach())

    def e2e_bytes(self):
        return self.host_x.numel() * 4, 4

    def extra_launches(self):
        return getattr(self, "graph_replays", 0) * getattr(self, "graph_kernel_nodes", 0)

    def units_per_step(self):
        return BATCH

    def config(self):
        return {"workload": "FlowNetF fwd+bwd (BASELINE cfg2)", "batch": BATCH, "image": "128x128",
                "conv_math": "fp32-accurate tcgen05 (3xTF32 forward, 3xBF16 gradients); library ops strict fp32",
                "launch": "forward+backward replayed as one CUDA graph" if getattr(self, "use_graph", True) else "eager launches",
                "weights": "random init (MSRA)"}

    def step_roofline(self, pk, ms_per_step):
        tflops = GFLOP_PER_IMAGE * BATCH / 1e3 / (ms_per_step * 1e-3)
        return {"kernel": "whole step (tcgen05 convolutions conv_gen_tc / conv3x3_tc / conv_gen_wgrad_tc; the net is launch / weight-bandwidth bound)", "bound": "tensor",
                "achieved": tflops, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": tflops / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": pk["source"]}

    def roofline(self, pk):
        return None

    def kernel_table(self, pk):
        return None

    @classmethod
    def _cpu_time(cls, steps, warmup):
        """The reference's own `models.base_networks.FlowNet(64)` (baseline/_ref: byte-identical staged copy) on the host
        CPU; none of this repo's modules on that path."""
        from baseline import ref_harness as H
        torch.set_num_threads(os.cpu_count())
        H.import_reference()
        import importlib
        RB = importlib.import_module("models.base_networks")
        w = cls(device=None)
        _, x, cots = w._build_inputs()
        torch.manual_seed(0)
        net = RB.FlowNet(64).train()
        w.net, w.cots = net, cots
        for _ in range(warmup):
            w._fb(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            w._fb(x)
        dt = (time.perf_counter() - t0) / steps
        return BATCH / dt, dt, "the reference's own FlowNet(64) (unmodified, baseline/_ref) fwd+bwd on the host CPU (PyTorch CPU kernels), batch %d, %d step(s)" % (BATCH, steps)

    @classmethod
    def cpu_baseline(cls):
        v, dt, sample = cls._cpu_time(2, 1)
        return {"value": v, "unit": cls.UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": sample}

    @classmethod
    def run_reference(cls, steps, warmup, n_gpus):
        v, dt, sample = cls._cpu_time(max(1, steps), warmup)
        return {"impl": "reference", "metric": cls.METRIC, "value": v, "unit": cls.UNIT, "n_gpus": n_gpus, "steps": steps,
                "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": cls.DTYPE, "data": "synthetic", "config": dict(cls(device=None).config(), sample=sample),
                "cpu_baseline": {"value": v, "unit": cls.UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": sample},
                "e2e": {"value": v, "unit": cls.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
