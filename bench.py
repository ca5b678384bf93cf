#!/usr/bin/env python
"""bench.py — the driver's measurement contract for ffwm_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One JSON line on stdout (rank 0).  Workloads live in benchmarks/*.py; each is
"one pass of the hot path over one batch of synthetic input":

    train_step   FFWM train step (BASELINE config 3; batch 8/GPU, 128x128; config 4 under torchrun)
                 [default].  At N=1 the flow-warp microbench below is attached as `warp_microbench`.
    warp         flow-warp microbench (BASELINE config 5): resample2d, block_extractor,
                 local_attn_reshape and the bilinear grid-warp, forward+backward, on feature
                 maps far larger than L2

    flownet      FlowNet pre-training forward+backward (BASELINE config 2; batch 6), CUDA-graph replay

`--impl reference` times the reference's own CPU implementation of the workload on all host cores:
for train_step the UNMODIFIED `FFWMModel(gpu_ids=[])` at batch 8 from `baseline/_ref` (a byte-identical,
git-ignored staging of the reference's `models/` + `lightcnn/`, see baseline/stage_ref.py), for flownet
the reference's `FlowNet(64)`, for warp the reference's formulation (C oracle / torch grid_sample) — the
only place besides tests/ and smoke() that executes oracle/.  Under torchrun rank 0 alone runs it.
Nothing here reads the reference checkout (absent on the GPU box).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed
    region runs (the profiling recipe's `nvidia-smi -lms` line, in-process)."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTED = {"sw_power_cap": 0x4}

    def __init__(self, cuda_index, period=0.02):
        self.period, self.samples, self.reasons = period, [], set()
        self._stop = threading.Event()
        self._thread = None
        self.sm_max = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.nv = pynvml
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for name, bit in list(self.BAD.items()) + list(self.NOTED.items()):
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0,
                    "note": getattr(self, "err", "no samples")}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------- helpers
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def tensor_peaks(local_rank, n=8192, iters=10):
    """cuBLAS GEMM rates measured in this run (burst, best of `iters`, CUDA events): the TF32 rate BASELINE.md 2 asks
    for before any TF32 fraction is quoted, and the bf16 rate next to MEASURED_PEAKS.json's.  Never fatal."""
    import torch
    out = {}
    try:
        dev = torch.device("cuda", local_rank)
        saved = torch.backends.cuda.matmul.allow_tf32
        for name, dtype, tf32 in (("tf32_tflops", torch.float32, True), ("bf16_tflops", torch.bfloat16, False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a = torch.randn(n, n, device=dev, dtype=dtype)
            b = torch.randn(n, n, device=dev, dtype=dtype)
            for _ in range(3):
                a @ b
            best = 1e30
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                a @ b
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name] = round(2.0 * n ** 3 / (best * 1e-3) / 1e12, 1)
            del a, b
        torch.backends.cuda.matmul.allow_tf32 = saved
        out["how"] = "torch.matmul %d^3 (cuBLAS), best of %d, CUDA events, this run" % (n, iters)
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)
    return out


def warp_microbench(local_rank, pk, steps=10, warmup=3):
    """BASELINE's second metric (resample2d / block_extractor HBM GB/s): the `warp` workload's
    device leg, attached to the train-step line at N=1."""
    import torch
    from benchmarks.warp import WarpWorkload
    from ffwm_b200 import _lib
    wl = WarpWorkload(device=torch.device("cuda", local_rank))
    wl.setup()
    for _ in range(warmup):
        wl.step(timed=False)
    torch.cuda.synchronize()
    n0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        wl.step(timed=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"metric": wl.METRIC, "value": wl.units_per_step() / (ms * 1e-3), "unit": wl.UNIT, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "config": wl.config(), "gpu_launches": _lib.kernel_launches() - n0,
            "roofline": wl.roofline(pk), "kernels": wl.kernel_table(pk)}


def note(msg):
    """Phase marker on stderr (rank-tagged), so a stalled multi-GPU run shows where it stopped."""
    sys.stderr.write("[bench r%s %6.1fs] %s\n" % (os.environ.get("RANK", "0"), time.time() - _T0, msg))
    sys.stderr.flush()


_T0 = time.time()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FFWM_BENCH_WORKLOAD", "default"))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-warp", action="store_true", help="train_step: skip the attached flow-warp microbench")
    ap.add_argument("--no-library-baseline", action="store_true",
                    help="train_step: skip the same-run cuDNN/ATen baseline (fp32 and TF32) of the same trainer")
    args = ap.parse_args()

    from benchmarks import get_workload
    wl_cls = get_workload(args.workload)
    rank, local_rank, world = dist_env()

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = args.steps if args.steps is not None else wl_cls.REF_STEPS
        warmup = args.warmup if args.warmup is not None else wl_cls.REF_WARMUP
        line = wl_cls.run_reference(steps=steps, warmup=max(0, warmup), n_gpus=args.gpus)
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — ffwm_b200 has no CPU path (use --impl reference for the CPU arm)")
    steps = args.steps if args.steps is not None else wl_cls.STEPS
    warmup = max(3, args.warmup if args.warmup is not None else wl_cls.WARMUP)
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import ffwm_b200
    from ffwm_b200 import _lib
    wl = wl_cls(device=torch.device("cuda", local_rank), rank=rank, world=world)
    note("setup")
    wl.setup()
    note("warm-up")

    # ---- device-resident leg: inputs already in HBM ------------------------------------
    for _ in range(warmup):
        wl.step(timed=False)
    barrier()
    note("timed region")
    sampler = ClockSampler(local_rank)
    launches0 = _lib.kernel_launches() + wl.extra_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    ncu_range = os.environ.get("FFWM_BENCH_NCU_RANGE") == "1"     # `ncu --profile-from-start off`
    if ncu_range:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(steps):
        wl.step(timed=True)
    ev1.record()
    barrier()
    if ncu_range:
        torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = _lib.kernel_launches() + wl.extra_launches() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / steps
    units = wl.units_per_step() * world
    value = units / (ms_per_step * 1e-3)

    # ---- end-to-end leg: host buffers through the public API --------------------------
    note("device leg done: %.2f ms/step" % ms_per_step)
    if hasattr(wl, "phase_report") and wl.phase_report() is not None:
        note("phases (ms): %s" % json.dumps(wl.phase_report()))
        wl.trainer.phase_events = None
    e2e = None
    if not args.no_e2e:
        for _ in range(3):
            wl.step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = max(3, min(steps, wl.E2E_STEPS))
        e0.record()
        for _ in range(e2e_steps):
            wl.step_e2e()
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d, d2h = wl.e2e_bytes()
        eunits = (wl.e2e_units() if hasattr(wl, "e2e_units") else wl.units_per_step()) * world
        e2e = {"value": eunits / (float(ems.item()) / e2e_steps * 1e-3), "unit": wl.UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps}

    note("e2e leg done")
    if rank == 0:
        pk = peaks()
        line = {
            "metric": wl.METRIC, "value": value, "unit": wl.UNIT, "n_gpus": n_gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": wl.DTYPE, "data": "synthetic", "config": wl.config(),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": (wl.step_roofline(pk, ms_per_step) if hasattr(wl, "step_roofline") else wl.roofline(pk)),
            "kernels": wl.kernel_table(pk),
        }
        if args.workload in ("default", "train_step") and world == 1 and not args.no_warp:
            line["warp_microbench"] = warp_microbench(local_rank, pk)
        if hasattr(wl, "gpu_library_baseline") and world == 1 and not args.no_library_baseline:
            note("library baseline (same trainer on cuDNN / ATen)")
            line["gpu_library_baseline"] = wl.gpu_library_baseline()
            line["measured_tensor_peaks"] = tensor_peaks(local_rank)
        if not args.no_cpu_baseline and world == 1:
            note("cpu baseline")
            line["cpu_baseline"] = wl_cls.cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
