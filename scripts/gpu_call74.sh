#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | tail -6
bash scripts/gpu_step_ab.sh c74 FFWM_BN_NO_SMALL=1 | grep -v "^ " | tail -6
head -3 $O/c74_launches_train_summary.txt | cut -c1-120
