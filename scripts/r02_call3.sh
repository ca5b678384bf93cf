#!/bin/bash
# Round 2, GPU call 3: 3xBF16 weight gradient (parity + per-shape + train-step A/B), per-kernel launch list of one eager step.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_wgrad_gpu.py tests/test_mfm_gpu.py tests/test_conv_tc_gpu.py -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02c_pytest.log
timeout 300 python -m benchmarks.conv --wgrad --out $O/r02c_conv_wgrad.json > $O/r02c_conv_wgrad.txt 2>&1; tail -10 $O/r02c_conv_wgrad.txt
bench() { env $2 timeout 500 python bench.py --no-cpu-baseline --no-warp --no-library-baseline > $O/r02c_bench_$1.json 2> $O/r02c_bench_$1.err; echo "bench $1 rc=$?"; }
bench default "X=1"
bench wgrad "FFWM_WGRAD_TC=1"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms/step", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
FFWM_WGRAD_TC=1 FFWM_BENCH_GRAPH=0 FFWM_BENCH_NCU_RANGE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv --log-file $O/launches_train_r02c.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-warp --no-library-baseline > $O/ncu_launches_train_r02c.log 2>&1; echo "ncu train launches rc=$?"
[[ -f $O/launches_train_r02c.csv ]] && python scripts/launch_summary.py $O/launches_train_r02c.csv $O/r02c_launches_train_summary.txt --rm
head -60 $O/r02c_launches_train_summary.txt
