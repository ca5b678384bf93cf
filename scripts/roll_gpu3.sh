#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_warp_gpu.py -q -k "scatter or roll or full_size" --tb=short > gpurun_out/pytest_scatter.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_scatter.log | cut -c1-300
timeout 600 python scripts/roll_ab.py --iters 10 > gpurun_out/roll_ab.txt 2>&1; echo "ab rc=$?"
grep -v "gflow\|_fwd" gpurun_out/roll_ab.txt | tail -40
