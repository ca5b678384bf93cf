#!/bin/bash
# Convolution kernels: parity tests, per-shape benches against cuDNN, train-step bench.   gpurun --timeout 1500 -- 'bash scripts/gpu_conv_check.sh TAG'
TAG=${1:-x}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_conv_gen_gpu.py tests/test_conv_tc_gpu.py tests/test_conv_nt128_gpu.py -q -x > $O/${TAG}_pytest.log 2>&1; echo "conv pytest rc=$?"; tail -3 $O/${TAG}_pytest.log | cut -c1-200
timeout 300 python -m benchmarks.conv --out $O/${TAG}_conv.json > $O/${TAG}_conv.txt 2>&1; tail -9 $O/${TAG}_conv.txt
timeout 300 python -m benchmarks.conv --gen --out $O/${TAG}_conv_gen.json > $O/${TAG}_conv_gen.txt 2>&1; tail -18 $O/${TAG}_conv_gen.txt
timeout 300 python -m benchmarks.conv --wgrad --out $O/${TAG}_conv_wgrad.json > $O/${TAG}_conv_wgrad.txt 2>&1; tail -9 $O/${TAG}_conv_wgrad.txt
timeout 500 python bench.py --no-cpu-baseline --no-warp --no-library-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
python - $TAG <<'PY'
import json, sys
d = json.loads(open("gpurun_out/%s_bench.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms/step", d.get("gpu_launches"))
PY
