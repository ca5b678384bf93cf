#!/bin/bash
# Round 2, GPU call 1: run the kernels written blind at the end of round 1 (tcgen05 weight gradient,
# 128-channel CTA tile, fused MFM, fused guided filter), A/B every opt-in switch on the train step,
# record the warp microbench (iid + smooth flow) and BASELINE config 2.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/r02_call1.sh'
mkdir -p gpurun_out
O=gpurun_out
run_test() {  # name, env, files...
    local name=$1 env=$2; shift 2
    env $env FFWM_EXPERIMENTAL=1 timeout 300 python -m pytest "$@" -x -q > $O/r02_${name}_pytest.log 2>&1
    local rc=$?; echo "$name pytest rc=$rc"; tail -6 $O/r02_${name}_pytest.log; return $rc
}
bench() {  # name, env
    env $2 timeout 400 python bench.py --no-cpu-baseline --no-warp > $O/r02_bench_$1.json 2> $O/r02_bench_$1.err
    echo "bench $1 rc=$?"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
run_test wgrad "X=1" tests/test_zz_wgrad_tc_gpu.py; WG=$?
run_test wgrad_direct "FFWM_WGRAD_DIRECT_EPILOGUE=1" tests/test_zz_wgrad_tc_gpu.py; WGD=$?
run_test nt128 "X=1" tests/test_zz_conv_nt128_gpu.py; NT=$?
run_test mfm "X=1" tests/test_zz_mfm_gpu.py; MFM=$?
run_test gf "X=1" tests/test_zz_guided_filter_gpu.py; GF=$?
if [ $WG -eq 0 ]; then timeout 300 python -m benchmarks.conv --wgrad --out $O/r02_conv_wgrad.json > $O/r02_conv_wgrad.txt 2>&1; tail -12 $O/r02_conv_wgrad.txt; fi
if [ $WGD -eq 0 ]; then FFWM_WGRAD_DIRECT_EPILOGUE=1 timeout 300 python -m benchmarks.conv --wgrad --out $O/r02_conv_wgrad_direct.json > $O/r02_conv_wgrad_direct.txt 2>&1; tail -12 $O/r02_conv_wgrad_direct.txt; fi
if [ $NT -eq 0 ]; then timeout 300 python -m benchmarks.conv --nt128 --out $O/r02_conv_nt128.json > $O/r02_conv_nt128.txt 2>&1; tail -8 $O/r02_conv_nt128.txt; fi
bench base "X=1"
K=""
[ $WG -eq 0 ] && { bench wgrad "FFWM_WGRAD_TC=1"; K="$K FFWM_WGRAD_TC=1"; }
[ $NT -eq 0 ] && { bench nt128 "FFWM_CONV_NT128=1"; K="$K FFWM_CONV_NT128=1"; }
[ $MFM -eq 0 ] && { bench mfm "FFWM_FUSED_MFM=1"; K="$K FFWM_FUSED_MFM=1"; }
[ $GF -eq 0 ] && { bench gf "FFWM_FUSED_GF=1"; K="$K FFWM_FUSED_GF=1"; }
run_test host "FFWM_BATCHED_SN=1 FFWM_BATCHED_VGG=1 FFWM_CACHE_PACKED=1 FFWM_FLOW_STREAMS=1" tests/test_networks.py tests/test_train_step.py -m gpu; HOST=$?
bench sn "FFWM_BATCHED_SN=1"
bench vgg "FFWM_BATCHED_VGG=1"
bench cache "FFWM_CACHE_PACKED=1"
bench streams "FFWM_FLOW_STREAMS=1"
bench host "FFWM_BATCHED_SN=1 FFWM_BATCHED_VGG=1 FFWM_CACHE_PACKED=1"
bench all "X=1 $K FFWM_BATCHED_SN=1 FFWM_BATCHED_VGG=1 FFWM_CACHE_PACKED=1"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], "ms/step", d.get("gpu_launches"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 400 python bench.py --workload warp --no-cpu-baseline > $O/r02_warp_iid.json 2> $O/r02_warp_iid.err; echo "warp iid rc=$?"
FFWM_BENCH_FLOW=smooth timeout 400 python bench.py --workload warp --no-cpu-baseline > $O/r02_warp_smooth.json 2> $O/r02_warp_smooth.err; echo "warp smooth rc=$?"
python - <<'PY'
import json
for f in ("r02_warp_iid", "r02_warp_smooth"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.0f GB/s" % d["value"], {k: (v["ms"], v["frac_hbm"]) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 400 python bench.py --workload flownet --no-cpu-baseline > $O/r02_bench_flownet.json 2> $O/r02_bench_flownet.err; echo "flownet rc=$?"; cat $O/r02_bench_flownet.json
ls -la $O | tail -30
