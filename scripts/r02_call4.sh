#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gen_gpu.py -q > $O/r02d_pytest.log 2>&1; echo "conv_gen pytest rc=$?"; tail -40 $O/r02d_pytest.log | cut -c1-220
