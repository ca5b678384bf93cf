#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 600 python -m pytest tests/test_spectral_gpu.py -q -x 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -20
bash scripts/gpu_step_ab.sh c72 FFWM_FUSED_SN=0 | grep -v "^ " | tail -8
head -3 $O/c72_launches_train_summary.txt; grep "sn_" $O/c72_launches_train_summary.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED" | tail -5
