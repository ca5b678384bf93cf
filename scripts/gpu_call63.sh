#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 600 python -m pytest tests/test_train_step.py tests/test_conv_tc_gpu.py tests/test_batch_norm_gpu.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed" | head -12
i=0
for shp in "F 3x3 512->512 @8" "F 3x3s2 64->128 @64" "3x3 384->384 @16"; do
  i=$((i+1))
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gen_tc_kernel -s 4 -c 1 -f -o $O/prof_cg$i python scripts/conv_one.py "$shp" > $O/ncu_cg$i.log 2>&1; echo "ncu $shp rc=$?"
  python scripts/ncu_hot.py $O/prof_cg$i.ncu-rep conv_gen_tc_kernel 0x200 > $O/cg${i}_hot.txt 2>&1
  ncu -i $O/prof_cg$i.ncu-rep --page details 2>/dev/null | grep -E "Duration|Elapsed Cycles|Registers Per|Theoretical Occ|Achieved Occ|Executed Ipc|No Eligible|DRAM Throughput|L2 Cache Throughput|Stall|Grid Size|Block Size" | head -20
done
