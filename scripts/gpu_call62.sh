#!/bin/bash
cd "$(dirname "$0")/.."
for e in "FFWM_NOP=1" "FFWM_FUSED_BN=0" "FFWM_FUSED_ADAM=0" "FFWM_FUSED_BN=0 FFWM_FUSED_ADAM=0"; do
  echo "=== $e"; env $e timeout 600 python -m pytest tests/test_train_step.py -m gpu -q -k matches_reference_on_gpu 2>&1 | grep -E "^E  |passed|failed" | head -8
done
