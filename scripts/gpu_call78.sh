#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gen_gpu.py -q -x 2>&1 | grep -E "^E  |passed|failed|FAILED|Error" | head -6
bash scripts/gpu_step_ab.sh c78 FFWM_CONV_SPLIT_FILL=1 | grep -v "^ " | tail -6
python - <<'PY'
import json
d=json.loads(open("gpurun_out/c78_bench_default.json").read().strip().splitlines()[-1])
for k,v in d.get("kernels",{}).items():
    if "batch_norm" in k or "channel_sum" in k: print(k, v)
PY
