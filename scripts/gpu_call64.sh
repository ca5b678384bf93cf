#!/bin/bash
cd "$(dirname "$0")/.."
for e in "FFWM_NOP=1" "FFWM_FUSED_ADAM=0" "FFWM_FUSED_BN=0"; do echo "=== $e"; env $e timeout 600 python scripts/diag_graph.py 2>&1 | grep "^loss"; done
