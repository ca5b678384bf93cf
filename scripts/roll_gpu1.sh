#!/bin/bash
# GPU pass for the rolling-strip kernels: parity tests, a memcheck of one small case, A/B timing.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests/test_warp_gpu.py -q -k "roll" -x --tb=short > gpurun_out/pytest_roll.log 2>&1; echo "pytest roll rc=$?"
tail -40 gpurun_out/pytest_roll.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_warp_gpu.py -q -x --tb=line \
  -k "roll and (33 or 21 or 37-3-rand or 50)" > gpurun_out/memcheck_roll.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/memcheck_roll.log | head -20
timeout 600 python scripts/roll_ab.py --iters 10 > gpurun_out/roll_ab.txt 2>&1; echo "ab rc=$?"
cat gpurun_out/roll_ab.txt | tail -60
