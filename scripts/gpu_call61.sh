#!/bin/bash
cd "$(dirname "$0")/.."; O=gpurun_out; mkdir -p $O
bash scripts/gpu_step_ab.sh c61
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -15
