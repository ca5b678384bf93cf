#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:roll -f -o gpurun_out/prof_roll \
    python scripts/roll_one.py gw_fwd rs4_fwd be_fwd rs4_gflow > gpurun_out/ncu_roll.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/ncu_roll.log
ls -la gpurun_out/prof_roll.ncu-rep
