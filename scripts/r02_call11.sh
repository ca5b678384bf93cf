#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02h_pytest.log 2>&1; echo "pytest -m gpu rc=$?"; tail -12 $O/r02h_pytest.log | cut -c1-200
python scripts/diag_nets.py 2>&1 | grep -v Warn > $O/r02h_network_accuracy.txt; cat $O/r02h_network_accuracy.txt | cut -c1-300
bench() { env $2 timeout 500 python bench.py --no-cpu-baseline --no-warp --no-library-baseline > $O/r02h_bench_$1.json 2> $O/r02h_bench_$1.err; echo "bench $1 rc=$?"; }
bench default "X=1"
bench fwdbf16 "FFWM_CONV_MATH_FWD=1"
bench nowgen "FFWM_WGRAD_GEN=0"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02h_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms/step", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 300 python -m benchmarks.conv --gen --out $O/r02h_conv_gen.json > $O/r02h_conv_gen.txt 2>&1; tail -20 $O/r02h_conv_gen.txt
