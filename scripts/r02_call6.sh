#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_conv_gen_gpu.py -q -k "wgrad" > $O/r02f_pytest.log 2>&1; echo "wgrad-gen pytest rc=$?"; tail -25 $O/r02f_pytest.log | cut -c1-200
FFWM_WG_DESC_SWAP=1 timeout 600 python -m pytest tests/test_conv_gen_gpu.py -q -k "wgrad" > $O/r02f_pytest_swap.log 2>&1; echo "wgrad-gen (LBO/SBO swapped) pytest rc=$?"; tail -8 $O/r02f_pytest_swap.log | cut -c1-200
timeout 600 python -m pytest tests/test_conv_wgrad_gpu.py -q > $O/r02f_pytest3.log 2>&1; echo "3x3 wgrad pytest rc=$?"; tail -4 $O/r02f_pytest3.log | cut -c1-200
timeout 300 python -m benchmarks.conv --wgrad --out $O/r02f_conv_wgrad.json > $O/r02f_conv_wgrad.txt 2>&1; tail -10 $O/r02f_conv_wgrad.txt
