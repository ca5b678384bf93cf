#!/bin/bash
# A/B of train-step switches + launch list.  usage: gpu_step_ab.sh TAG "ENV1=.. ENV2=.." "ENV.." ...
cd "$(dirname "$0")/.."; O=gpurun_out; mkdir -p $O; TAG=$1; shift
timeout 900 python -m pytest tests/test_batch_norm_gpu.py tests/test_conv_gen_gpu.py tests/test_conv_tc_gpu.py -x -q 2>&1 | grep -v Warning | tail -30
run() {  # name, env...
    local name=$1; shift
    env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-warp > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
    python - $O/${TAG}_bench_$name.json "$name $*" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("%-40s %.2f ms  %.1f img/s  e2e %.1f" % (sys.argv[2], d["ms_per_step"], d["value"], (d.get("e2e") or {}).get("value") or 0))
except Exception as e: print(sys.argv[2], "failed", e)
PY
}
run default FFWM_NOP=1
i=0
for cfg in "$@"; do i=$((i+1)); run ab$i $cfg; done
run default2 FFWM_NOP=1
FFWM_BENCH_GRAPH=0 FFWM_BENCH_NCU_RANGE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv --log-file $O/launches_train_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-warp --no-library-baseline > $O/ncu_launches_train_$TAG.log 2>&1; echo "ncu train launches rc=$?"
[[ -f $O/launches_train_$TAG.csv ]] && python scripts/launch_summary.py $O/launches_train_$TAG.csv $O/${TAG}_launches_train_summary.txt && gzip -f $O/launches_train_$TAG.csv
head -45 $O/${TAG}_launches_train_summary.txt | cut -c1-150
