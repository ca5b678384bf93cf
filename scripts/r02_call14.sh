#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/dp_check.py > $O/r02k_dp_check.log 2>&1; echo "dp_check rc=$?"; grep -v Warn $O/r02k_dp_check.log | tail -8 | cut -c1-250
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline > $O/r02k_bench_2gpu.json 2> $O/r02k_bench_2gpu.err; echo "bench 2gpu overlap rc=$?"; tail -3 $O/r02k_bench_2gpu.err
FFWM_DP_OVERLAP=0 timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline > $O/r02k_bench_2gpu_nooverlap.json 2> $O/r02k_bench_2gpu_nooverlap.err; echo "bench 2gpu no-overlap rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-warp --no-library-baseline > $O/r02k_bench_1gpu.json 2> $O/r02k_bench_1gpu.err; echo "bench 1gpu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02k_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms/step", "e2e", round(d["e2e"]["value"],2))
    except Exception as e:
        print(f, "unreadable:", e)
PY
