#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_|channel_sum" -s 10 -c 5 -f -o $O/prof_bn python scripts/bn_one.py > $O/ncu_bn.log 2>&1; echo "ncu bn rc=$?"
python scripts/ncu_summary.py $O/prof_bn.ncu-rep r02y_bn $O > /dev/null 2>&1; cat $O/r02y_bn_ncu_summary.txt | cut -c1-400
bash scripts/gpu_step_ab.sh c71 FFWM_BATCHED_L1=0 | grep -v "^ " | tail -8
grep -c . $O/c71_launches_train_summary.txt; head -3 $O/c71_launches_train_summary.txt
