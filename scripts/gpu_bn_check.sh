#!/bin/bash
# batch-norm kernels: parity tests, microbench, train-step A/B, then the whole GPU suite
cd "$(dirname "$0")/.."; O=gpurun_out; mkdir -p $O; TAG=${1:-bn}
timeout 600 python -m pytest tests/test_batch_norm_gpu.py -x -q 2>&1 | tail -15
timeout 300 python benchmarks/batch_norm.py 2>&1 | tee $O/${TAG}_batch_norm.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-warp > $O/${TAG}_bench_on.json 2> $O/${TAG}_bench_on.err; tail -c 600 $O/${TAG}_bench_on.json | head -c 300; echo
FFWM_FUSED_BN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-warp > $O/${TAG}_bench_off.json 2> $O/${TAG}_bench_off.err
python - $TAG <<'PY'
import json,sys
for t in ("on","off"):
    try:
        d=json.loads(open("gpurun_out/%s_bench_%s.json"%(sys.argv[1] if len(sys.argv)>1 else "bn",t)).read().strip().splitlines()[-1]); print(t, d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"))
    except Exception as e: print(t, "failed", e)
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
