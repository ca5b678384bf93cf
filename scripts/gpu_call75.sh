#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | tail -6
bash scripts/gpu_step_ab.sh c75 FFWM_LOSSNET_MATH_FWD=0 | grep -v "^ " | tail -6
