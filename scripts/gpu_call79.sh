#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 900 python -m pytest tests/test_conv_few_gpu.py -q -x 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -12
timeout 900 python -m pytest tests/test_conv_gen_gpu.py tests/test_conv_tc_gpu.py tests/test_networks.py tests/test_orchestrators.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed|FAILED" | head -8
bash scripts/gpu_step_ab.sh c79 FFWM_CONV_FEW=0 | grep -v "^ " | tail -6
grep "conv_few" $O/c79_launches_train_summary.txt | cut -c1-110
