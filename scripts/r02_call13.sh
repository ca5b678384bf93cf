#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_fused_losses.py -q > $O/r02j_pytest.log 2>&1; echo "fused losses pytest rc=$?"; tail -30 $O/r02j_pytest.log | cut -c1-220
timeout 600 python -m pytest tests/test_flownet_step.py tests/test_orchestrators.py -m gpu -q > $O/r02j_pytest2.log 2>&1; echo "flownet step goldens rc=$?"; tail -5 $O/r02j_pytest2.log | cut -c1-220
timeout 400 python bench.py --workload flownet --no-cpu-baseline > $O/r02j_bench_flownet.json 2> $O/r02j_bench_flownet.err; echo "flownet rc=$?"; cut -c1-330 $O/r02j_bench_flownet.json; tail -3 $O/r02j_bench_flownet.err
