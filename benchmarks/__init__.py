"""Benchmark workloads for bench.py (measurement harness, not product code)."""


def get_workload(name):
    if name in ("default",):
        name = DEFAULT
    if name == "warp":
        from .warp import WarpWorkload
        return WarpWorkload
    if name == "train_step":
        from .train_step import TrainStepWorkload
        return TrainStepWorkload
    if name == "flownet":
        from .flownet import FlowNetWorkload
        return FlowNetWorkload
    raise SystemExit("bench.py: unknown workload %r (train_step, warp, flownet)" % name)


# The workload BASELINE.json's metric is quoted on.
DEFAULT = "train_step"
