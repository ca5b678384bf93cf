#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_rows -c 2 -f -o gpurun_out/prof_rows \
    python scripts/roll_one.py rs4_gin1 > gpurun_out/ncu_rows.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_rows.log
python scripts/ncu_hot.py gpurun_out/prof_rows.ncu-rep scatter_rows 0x400 > gpurun_out/rows_hot.txt 2>&1
cat gpurun_out/rows_hot.txt
