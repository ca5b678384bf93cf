#!/bin/bash
# Round 2, GPU call 2: full GPU suite with the new defaults + 3xBF16 operand math, conv A/B of both maths, train-step benches.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest -m gpu rc=$?"; tail -15 $O/r02b_pytest.log
timeout 300 python -m benchmarks.conv --out $O/r02b_conv.json > $O/r02b_conv.txt 2>&1; echo "conv bench rc=$?"; cat $O/r02b_conv.txt | tail -12
bench() { env $2 timeout 500 python bench.py --no-cpu-baseline --no-warp $3 > $O/r02b_bench_$1.json 2> $O/r02b_bench_$1.err; echo "bench $1 rc=$?"; }
bench default "X=1" ""
bench tf32x3 "FFWM_CONV_MATH=0" "--no-library-baseline"
bench mfm "FFWM_FUSED_MFM=1" "--no-library-baseline"
bench wgrad "FFWM_WGRAD_TC=1" "--no-library-baseline"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02b_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms/step", d.get("gpu_launches"), d.get("gpu_library_baseline"), d.get("measured_tensor_peaks"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 $O/r02b_bench_default.err
