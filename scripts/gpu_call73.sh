#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 900 python -m pytest tests/test_pool_gpu.py tests/test_conv_gen_gpu.py tests/test_spectral_gpu.py -q -x 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -10
bash scripts/gpu_step_ab.sh c73 FFWM_CONV3X3_WIDTHS=128,64 FFWM_CONV3X3_WIDTHS=128 FFWM_FUSED_POOL=0 | grep -v "^ " | tail -9
grep "wgrad_reduce\|sn_wtu\|max_pool" $O/c73_launches_train_summary.txt | cut -c1-100
