#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m benchmarks.conv --wgrad --out $O/r02g_conv_wgrad.json > $O/r02g_conv_wgrad.txt 2>&1; tail -10 $O/r02g_conv_wgrad.txt
bench() { env $2 timeout 500 python bench.py --no-cpu-baseline --no-warp --no-library-baseline > $O/r02g_bench_$1.json 2> $O/r02g_bench_$1.err; echo "bench $1 rc=$?"; }
bench default "X=1"
bench wgen "FFWM_WGRAD_GEN=1"
bench wgen3 "FFWM_WGRAD_GEN=1 FFWM_WGRAD_GEN_3X3=1"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02g_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms/step", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
FFWM_WGRAD_GEN=1 timeout 900 python -m pytest tests/test_train_step.py tests/test_networks.py tests/test_orchestrators.py -m gpu -x -q > $O/r02g_pytest.log 2>&1; echo "goldens with general wgrad rc=$?"; tail -4 $O/r02g_pytest.log | cut -c1-200
