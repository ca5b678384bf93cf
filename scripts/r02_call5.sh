#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m benchmarks.conv --gen --out $O/r02e_conv_gen.json > $O/r02e_conv_gen.txt 2>&1; echo "gen bench rc=$?"; tail -22 $O/r02e_conv_gen.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02e_pytest.log 2>&1; echo "pytest -m gpu rc=$?"; tail -8 $O/r02e_pytest.log | cut -c1-200
bench() { env $2 timeout 500 python bench.py --no-cpu-baseline --no-warp --no-library-baseline > $O/r02e_bench_$1.json 2> $O/r02e_bench_$1.err; echo "bench $1 rc=$?"; }
bench default "X=1"
bench nogen "FFWM_CONV_GENERAL=0"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02e_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"],2), d["unit"], round(d["ms_per_step"],2), "ms/step", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 $O/r02e_bench_default.err
timeout 400 python bench.py --workload flownet --no-cpu-baseline > $O/r02e_bench_flownet.json 2> $O/r02e_bench_flownet.err; echo "flownet rc=$?"; cut -c1-400 $O/r02e_bench_flownet.json
FFWM_BENCH_GRAPH=0 FFWM_BENCH_NCU_RANGE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv --log-file $O/launches_train_r02e.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-warp --no-library-baseline > $O/ncu_launches_train_r02e.log 2>&1; echo "ncu train launches rc=$?"
[[ -f $O/launches_train_r02e.csv ]] && python scripts/launch_summary.py $O/launches_train_r02e.csv $O/r02e_launches_train_summary.txt --rm
head -45 $O/r02e_launches_train_summary.txt | cut -c1-150
