#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out; TAG=c70
timeout 900 python -m pytest tests/test_conv_gen_gpu.py -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -10
FFWM_CONV_OCC2=2 timeout 900 python -m pytest tests/test_conv_gen_gpu.py -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -10
timeout 300 python -m benchmarks.conv --gen --out $O/${TAG}_conv_gen.json > $O/${TAG}_conv_gen.txt 2>&1; cut -c1-75 $O/${TAG}_conv_gen.txt | tail -18
FFWM_CONV_OCC2=1 timeout 300 python -m benchmarks.conv --gen --out $O/${TAG}_conv_gen_occ1.json > $O/${TAG}_conv_gen_occ1.txt 2>&1; cut -c1-75 $O/${TAG}_conv_gen_occ1.txt | tail -18
FFWM_CONV_OCC2=2 timeout 300 python -m benchmarks.conv --gen --out $O/${TAG}_conv_gen_occ2.json > $O/${TAG}_conv_gen_occ2.txt 2>&1; cut -c1-75 $O/${TAG}_conv_gen_occ2.txt | tail -18
bash scripts/gpu_step_ab.sh c70 FFWM_CONV_OCC2=1 FFWM_CONV_OCC2=2 | grep -v "^ " | tail -8
