"""Flow-warp microbench (BASELINE config 5).

One step = forward + backward of the four warp-class ops of the hot path on
feature maps far larger than L2 (every big operand is >= 256 MiB, so each
launch streams from HBM; no explicit L2 flush is needed):

    resample2d (ks=4, per-pixel sigma plane = 2)   (B,C,R,R)  + (B,3,R,R)
    block_extractor (k=3)                          (Bb,C,R,R) + (Bb,2,R,R) -> (Bb,C,3R,3R)
    local_attn_reshape (k=3)                       (B2,9,H,W) -> (B2,1,3H,3W)
    grid_warp (WarpNet's bilinear sampler)         (B,C,R,R)  + (B,2,R,R)

The unit is ALGORITHMIC GB/s (SURVEY.md 8d "Algorithmic bytes"): bytes every
correct implementation has to move, summed over the step, divided by time.
Caller-side zero fills of the scatter targets are timed but are not counted
as algorithmic bytes.  Every kernel is also timed on its own with CUDA
events inside the timed region; `roofline` describes the kernel that takes
the largest share of the step.
"""
import os
import time

import torch

# (C, R) of the primary point; B follows SURVEY 8(d): operand >= 512 MiB.
C, R = 128, 128
K_BLOCK = 3
KS, DIL = 4, 1
# "iid" (default: SURVEY cfg5's independent per-pixel displacement) or "smooth" (see make_inputs.field)
FLOW_MODE = os.environ.get("FFWM_BENCH_FLOW", "iid")


def _mib(n):
    return n / (1 << 20)


def alg_bytes(op, **s):
    """SURVEY.md 8(d) algorithmic bytes, fp32."""
    if op == "resample2d_fwd":
        return 4 * (s["B"] * s["C"] * s["Hi"] * s["Wi"] + 3 * s["B"] * s["H"] * s["W"] + s["B"] * s["C"] * s["H"] * s["W"])
    if op == "resample2d_bwd":
        bchw, bhw = s["B"] * s["C"] * s["H"] * s["W"], s["B"] * s["H"] * s["W"]
        i1 = s["B"] * s["C"] * s["Hi"] * s["Wi"]
        return 4 * (i1 + 3 * bhw + bchw + i1 + 3 * bhw)
    if op == "block_extractor_fwd":
        return 4 * (s["B"] * s["C"] * s["Hs"] * s["Ws"] + 2 * s["B"] * s["Hf"] * s["Wf"]
                    + s["B"] * s["C"] * s["k"] ** 2 * s["Hf"] * s["Wf"])
    if op == "block_extractor_bwd":
        src, fl = s["B"] * s["C"] * s["Hs"] * s["Ws"], 2 * s["B"] * s["Hf"] * s["Wf"]
        return 4 * (src + fl + s["B"] * s["C"] * s["k"] ** 2 * s["Hf"] * s["Wf"] + src + fl)
    if op in ("local_attn_reshape_fwd", "local_attn_reshape_bwd"):
        return 4 * 2 * s["B"] * s["k"] ** 2 * s["H"] * s["W"]
    if op == "grid_warp_fwd":
        bchw, bhw = s["B"] * s["C"] * s["H"] * s["W"], s["B"] * s["H"] * s["W"]
        return 4 * (bchw + 2 * bhw + bchw)
    if op == "grid_warp_bwd":
        bchw, bhw = s["B"] * s["C"] * s["H"] * s["W"], s["B"] * s["H"] * s["W"]
        return 4 * (2 * bchw + 2 * bhw + bchw + 2 * bhw)
    raise KeyError(op)


def _identity_grid(b, r):
    lin = (torch.arange(r, dtype=torch.float32) * 2 + 1) / r - 1          # pixel centres, align_corners=False
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    return torch.stack((gx, gy), 0).unsqueeze(0).expand(b, 2, r, r)


def make_inputs(shapes, device, seed, pin=False):
    """Synthetic inputs of one step, seeded; `shapes` = dict(B, Bb, B2, C, R)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, Bb, B2, c, r = shapes["B"], shapes["Bb"], shapes["B2"], shapes["C"], shapes["R"]

    def rnd(*s, fn=torch.rand):
        t = fn(*s, generator=g)
        return t.pin_memory() if pin else t

    smooth = FLOW_MODE == "smooth"

    def field(b, fn=torch.randn):
        """(b, 2, r, r) displacement noise: independent per pixel (SURVEY cfg5), or — FFWM_BENCH_FLOW=smooth — noise
        drawn on an r/16 lattice and upsampled bilinearly: neighbouring pixels then move together, as the flow a
        network predicts does (same distribution at the lattice points, hence somewhat smaller between them)."""
        if not smooth:
            return rnd(b, 2, r, r, fn=fn)
        lo = fn(b, 2, max(r // 16, 2), max(r // 16, 2), generator=g)
        return torch.nn.functional.interpolate(lo, size=(r, r), mode="bilinear", align_corners=True)

    t = {
        "feat": rnd(B, c, r, r) * 2 - 1,
        "disp": torch.cat([field(B) * 2, torch.full((B, 1, r, r), 2.0)], 1),
        "gout": rnd(B, c, r, r, fn=torch.randn),
        # WarpNet's absolute sampling grid: identity + the same 2 px noise as resample2d
        "grid": _identity_grid(B, r) + field(B) * (2 * 2.0 / r),
        "be_src": rnd(Bb, c, r, r),
        "be_flow": field(Bb, fn=torch.rand) * 1.8,
        "be_gout": rnd(Bb, c, K_BLOCK * r, K_BLOCK * r, fn=torch.randn),
        "lar_in": rnd(B2, K_BLOCK * K_BLOCK, r, r),
        "lar_gout": rnd(B2, 1, K_BLOCK * r, K_BLOCK * r, fn=torch.randn),
    }
    if pin:
        t = {k: (v if v.is_pinned() else v.pin_memory()) for k, v in t.items()}
    if device is not None:
        t = {k: v.to(device) for k, v in t.items()}
    return t


OPS = ["resample2d_fwd", "resample2d_bwd", "block_extractor_fwd", "block_extractor_bwd",
       "local_attn_reshape_fwd", "local_attn_reshape_bwd", "grid_warp_fwd", "grid_warp_bwd"]


def op_bytes(shapes):
    B, Bb, B2, c, r = shapes["B"], shapes["Bb"], shapes["B2"], shapes["C"], shapes["R"]
    rs = dict(B=B, C=c, H=r, W=r, Hi=r, Wi=r)
    be = dict(B=Bb, C=c, Hs=r, Ws=r, Hf=r, Wf=r, k=K_BLOCK)
    la = dict(B=B2, k=K_BLOCK, H=r, W=r)
    return {
        "resample2d_fwd": alg_bytes("resample2d_fwd", **rs), "resample2d_bwd": alg_bytes("resample2d_bwd", **rs),
        "block_extractor_fwd": alg_bytes("block_extractor_fwd", **be),
        "block_extractor_bwd": alg_bytes("block_extractor_bwd", **be),
        "local_attn_reshape_fwd": alg_bytes("local_attn_reshape_fwd", **la),
        "local_attn_reshape_bwd": alg_bytes("local_attn_reshape_bwd", **la),
        "grid_warp_fwd": alg_bytes("grid_warp_fwd", **rs), "grid_warp_bwd": alg_bytes("grid_warp_bwd", **rs),
    }


class WarpWorkload:
    METRIC = "flow-warp (resample2d + block_extractor + local_attn_reshape + grid_warp) fwd+bwd algorithmic HBM GB/s"
    UNIT = "GB/s"
    DTYPE = "f32"
    STEPS, WARMUP = 20, 5
    E2E_STEPS = 3
    REF_STEPS, REF_WARMUP = 2, 1

    def __init__(self, device, rank=0, world=1):
        self.dev, self.rank, self.world = device, rank, world
        # B = max(8, ceil(512 MiB / (4 C R^2)))  
        B = max(8, -(-(512 << 20) // (4 * C * R * R)))
        self.shapes = dict(B=B, Bb=max(8, B // 4), B2=max(8, (512 << 20) // (4 * 9 * R * R)), C=C, R=R)
        self.bytes = op_bytes(self.shapes)
        self.events = []

    # ------------------------------------------------------------------ device leg
    def setup(self):
        from ffwm_b200 import ops
        self.ops = ops
        t = make_inputs(self.shapes, self.dev, seed=1234 + self.rank)
        self.t = t
        e = torch.empty_like
        self.out_rs, self.g1_rs, self.g2_rs = e(t["feat"]), e(t["feat"]), e(t["disp"])
        self.out_be = torch.empty_like(t["be_gout"])
        self.gs_be, self.gf_be = e(t["be_src"]), e(t["be_flow"])
        self.out_la, self.gi_la = torch.empty_like(t["lar_gout"]), e(t["lar_in"])
        self.out_gw, self.gi_gw, self.gf_gw = e(t["feat"]), e(t["feat"]), e(t["grid"])
        self._memsets = 0

    def _run(self, mark):
        o, t = self.ops, self.t
        mark()
        o.resample2d_forward(t["feat"], t["disp"], self.out_rs, KS, DIL)
        mark()
        self.g1_rs.zero_()
        o.resample2d_backward(t["feat"], t["disp"], t["gout"], self.g1_rs, self.g2_rs, KS, DIL)
        mark()
        o.block_extractor_forward(t["be_src"], t["be_flow"], self.out_be, K_BLOCK)
        mark()
        self.gs_be.zero_()
        o.block_extractor_backward(t["be_src"], t["be_flow"], t["be_gout"], self.gs_be, self.gf_be, K_BLOCK)
        mark()
        o.local_attn_reshape_forward(t["lar_in"], self.out_la, K_BLOCK)
        mark()
        o.local_attn_reshape_backward(t["lar_gout"], self.gi_la, K_BLOCK)
        mark()
        o.grid_warp_forward(t["feat"], t["grid"], self.out_gw)
        mark()
        self.gi_gw.zero_()
        o.grid_warp_backward(t["feat"], t["grid"], t["gout"], self.gi_gw, self.gf_gw)
        mark()
        self._memsets += 3

    def step(self, timed):
        if not timed:
            self._run(lambda: None)
            return
        evs = []

        def mark():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs.append(e)
        self._run(mark)
        self.events.append(evs)

    def extra_launches(self):
        return 0     # memsets are torch's, not ours: not counted in gpu_launches

    def units_per_step(self):
        return sum(self.bytes.values()) / 1e9

    def per_kernel_ms(self):
        """Mean device time of each op over the timed steps (call after a sync)."""
        acc = {k: 0.0 for k in OPS}
        for evs in self.events:
            for i, k in enumerate(OPS):
                acc[k] += evs[i].elapsed_time(evs[i + 1])
        n = max(1, len(self.events))
        return {k: v / n for k, v in acc.items()}

    def kernel_table(self, pk):
        ms = self.per_kernel_ms()
        tot = sum(ms.values())
        rows = {}
        for k in OPS:
            gbs = self.bytes[k] / 1e9 / (ms[k] * 1e-3) if ms[k] > 0 else None
            rows[k] = {"ms": round(ms[k], 4), "alg_GB": round(self.bytes[k] / 1e9, 4),
                       "GB/s": round(gbs, 1) if gbs else None,
                       "frac_hbm": round(gbs / pk["hbm_gbs"], 4) if gbs else None,
                       "share": round(ms[k] / tot, 4) if tot else None}
        return rows

    def roofline(self, pk):
        ms = self.per_kernel_ms()
        top = max(ms, key=ms.get)
        achieved = self.bytes[top] / 1e9 / (ms[top] * 1e-3)
        traffic = None
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        if os.path.exists(p):
            import json
            traffic = json.load(open(p)).get(top)
        return {"kernel": top, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"],
                "alg_bytes_per_launch": self.bytes[top], "ms_per_launch": ms[top],
                "note": "bwd times include the caller-side zero fill of the scatter target"}

    def config(self):
        s = self.shapes
        return {"workload": "warp (BASELINE cfg5 primary point)", "C": s["C"], "R": s["R"], "B": s["B"],
                "B_block_extractor": s["Bb"], "B_local_attn_reshape": s["B2"], "resample2d_ks": KS,
                "block_extractor_k": K_BLOCK, "l2": "inputs_exceed_l2 (every large operand >= 256 MiB)",
                "flow": ("independent per pixel: N(0, 2 px) / U(0, 1.8) px (SURVEY cfg5)" if FLOW_MODE != "smooth" else
                         "smooth: the same noise on an r/16 lattice, bilinearly upsampled (FFWM_BENCH_FLOW=smooth)"),
                "per_rank": "independent replicas, no collective"}

    # ------------------------------------------------------------------ end-to-end leg
    E2E_SHAPES = dict(B=8, Bb=8, B2=8, C=C, R=R)

    def _e2e_setup(self):
        from ffwm_b200 import external_function as E
        self.E = E
        self.host = make_inputs(self.E2E_SHAPES, None, seed=99 + self.rank, pin=True)
        self.res = E.Resample2d(KS, DIL, sigma=2).to(self.dev)
        self.be = E.BlockExtractor(K_BLOCK)
        self.la = E.LocalAttnReshape()
        self.host_out = None

    def step_e2e(self):
        """Public API (the reference-shaped nn.Modules), host buffers in, host result out."""
        if not hasattr(self, "host"):
            self._e2e_setup()
        h, E, dev = self.host, self.E, self.dev
        d = {k: v.to(dev, non_blocking=True) for k, v in h.items()}
        feat = d["feat"].requires_grad_(True)
        flow = d["disp"][:, :2].contiguous().requires_grad_(True)
        out = self.res(feat, flow)
        out.backward(d["gout"])
        src = d["be_src"].requires_grad_(True)
        bflow = d["be_flow"].requires_grad_(True)
        ob = self.be(src, bflow)
        ob.backward(d["be_gout"])
        li = d["lar_in"].requires_grad_(True)
        ol = self.la(li, K_BLOCK)
        ol.backward(d["lar_gout"])
        grid = d["grid"].requires_grad_(True)
        ow = E.grid_warp(feat, grid)
        ow.backward(d["gout"])
        results = [out, feat.grad, flow.grad, ob, src.grad, bflow.grad, ol, li.grad, ow, grid.grad]
        if self.host_out is None:
            self.host_out = [torch.empty(r.shape, dtype=r.dtype, pin_memory=True) for r in results]
        for ho, r in zip(self.host_out, results):
            ho.copy_(r, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def e2e_units(self):
        return sum(op_bytes(self.E2E_SHAPES).values()) / 1e9

    def e2e_bytes(self):
        h2d = sum(v.numel() * 4 for v in self.host.values())
        d2h = sum(v.numel() * 4 for v in self.host_out)
        return h2d, d2h

    # ------------------------------------------------------------------ CPU arm
    CPU_SHAPES = dict(B=8, Bb=4, B2=32, C=32, R=128)

    @classmethod
    def _cpu_step(cls, W, t):
        W.resample2d_forward(t["feat"], t["disp"], KS, DIL)
        W.resample2d_backward(t["feat"], t["disp"], t["gout"], KS, DIL)
        W.block_extractor_forward(t["be_src"], t["be_flow"], K_BLOCK)
        W.block_extractor_backward(t["be_src"], t["be_flow"], t["be_gout"], K_BLOCK)
        o = W.local_attn_reshape_forward(t["lar_in"], K_BLOCK)
        W.local_attn_reshape_backward(t["lar_in"], t["lar_gout"], K_BLOCK)
        W.grid_warp_forward(t["feat"], t["grid"])
        W.grid_warp_backward(t["feat"], t["grid"], t["gout"])
        return o

    @classmethod
    def _cpu_time(cls, steps, warmup):
        from oracle import warp as W     # the checker, timed as the CPU restatement of the reference kernels
        W.build()
        t = make_inputs(cls.CPU_SHAPES, None, seed=7)
        for _ in range(warmup):
            cls._cpu_step(W, t)
        t0 = time.perf_counter()
        for _ in range(steps):
            cls._cpu_step(W, t)
        dt = (time.perf_counter() - t0) / steps
        gb = sum(op_bytes(cls.CPU_SHAPES).values()) / 1e9
        s = cls.CPU_SHAPES
        sample = ("oracle/liboracle.so (C restatement of the reference .cu kernels, OpenMP where the kernel is a "
                  "gather; scatter-add kernels sequential) on the same 8 ops at B=%d/%d/%d C=%d R=%d, %d step(s)"
                  % (s["B"], s["Bb"], s["B2"], s["C"], s["R"], steps))
        return gb / dt, dt, sample

    @classmethod
    def cpu_baseline(cls):
        v, dt, sample = cls._cpu_time(steps=1, warmup=1)
        return {"value": v, "unit": cls.UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample,
                "s_per_sample": dt,
                "note": "the reference has no CPU implementation of these ops (NotImplementedError for CPU tensors)"}

    @classmethod
    def run_reference(cls, steps, warmup, n_gpus):
        v, dt, sample = cls._cpu_time(steps=max(1, steps), warmup=warmup)
        w = cls(device=None)
        return {"impl": "reference", "metric": cls.METRIC, "value": v, "unit": cls.UNIT, "n_gpus": n_gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": cls.DTYPE, "data": "synthetic",
                "config": dict(w.config(), sample=sample),
                "cpu_baseline": {"value": v, "unit": cls.UNIT, "cores": os.cpu_count(), "kind": "port",
                                 "sample": sample},
                "e2e": {"value": v, "unit": cls.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
