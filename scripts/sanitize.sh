#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (the reference has no race / memory checking; its
# backward kernels lean on atomicAdd).  Slow (10-50x): small-shape tests only.  Round 1 never got to run this
# (gpurun_out/memcheck_roll.log of that round deselected every test); queued for round 2:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/sanitize.sh'
mkdir -p gpurun_out
SEL='vs_oracle or forced_tiled or tiled_scatter or bit_exact or honour_strides or empty_and_ragged'
for tool in memcheck racecheck synccheck initcheck; do
    timeout 1100 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 30 \
        python -m pytest tests/test_warp_gpu.py -q -x -k "$SEL" > gpurun_out/sanitize_warp_$tool.log 2>&1
    echo "warp $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_warp_$tool.log | tail -3
done
# the tcgen05 convolution: memcheck only (shared-memory race tools do not model the async proxy / TMEM)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 \
    python -m pytest tests/test_conv_tc_gpu.py -q -x -k "dgrad_packing or rejects" > gpurun_out/sanitize_conv_memcheck.log 2>&1
echo "conv memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_conv_memcheck.log | tail -3
