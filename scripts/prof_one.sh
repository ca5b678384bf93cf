#!/bin/bash
# ncu --set full on one op of scripts/roll_one.py: bash scripts/prof_one.sh <op> <kernel regex>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -f -o gpurun_out/prof_one \
    python scripts/roll_one.py $1 > gpurun_out/ncu_one.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_hot.py gpurun_out/prof_one.ncu-rep $2 0x200 > gpurun_out/one_hot.txt 2>&1
head -60 gpurun_out/one_hot.txt
