#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out; TAG=c66
timeout 900 python -m pytest tests/test_conv_gen_gpu.py -q -k "wgrad" 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -30
FFWM_WGRAD_NO_ROWS=1 timeout 900 python -m pytest tests/test_conv_gen_gpu.py -q -k "wgrad" 2>&1 | grep -E "passed|failed" | head -3
timeout 300 python -m benchmarks.conv --wgrad --out $O/${TAG}_conv_wgrad.json > $O/${TAG}_conv_wgrad.txt 2>&1; tail -9 $O/${TAG}_conv_wgrad.txt | cut -c1-120
FFWM_WGRAD_NO_ROWS=1 timeout 300 python -m benchmarks.conv --wgrad --out $O/${TAG}_conv_wgrad_norows.json > $O/${TAG}_conv_wgrad_norows.txt 2>&1; tail -9 $O/${TAG}_conv_wgrad_norows.txt | cut -c1-120
bash scripts/gpu_step_ab.sh c66 FFWM_WGRAD_NO_ROWS=1 | grep -v "^ " | tail -8
