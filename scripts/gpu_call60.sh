#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out; mkdir -p $O; TAG=c60
timeout 300 python -m benchmarks.conv --wgrad --out $O/${TAG}_conv_wgrad64.json > $O/${TAG}_conv_wgrad64.txt 2>&1; tail -9 $O/${TAG}_conv_wgrad64.txt | cut -c1-120
FFWM_WGRAD_CHAIN=128 timeout 300 python -m benchmarks.conv --wgrad --out $O/${TAG}_conv_wgrad128.json > $O/${TAG}_conv_wgrad128.txt 2>&1; tail -9 $O/${TAG}_conv_wgrad128.txt | cut -c1-120
FFWM_WGRAD_CHAIN=256 timeout 300 python -m benchmarks.conv --wgrad --out $O/${TAG}_conv_wgrad256.json > $O/${TAG}_conv_wgrad256.txt 2>&1; tail -9 $O/${TAG}_conv_wgrad256.txt | cut -c1-120
bash scripts/gpu_step_ab.sh c60 FFWM_WGRAD_CHAIN=128 FFWM_WGRAD_CHAIN=256
FFWM_WGRAD_CHAIN=128 timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -12
