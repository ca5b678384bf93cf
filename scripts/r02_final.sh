#!/bin/bash
# Round-2 evidence run on one B200:   gpurun --timeout 3000 -- 'bash scripts/r02_final.sh TAG'
# full GPU suite, the driver's bench lines (ours + reference arm), the warp microbench (iid + smooth flow), cfg2, the per-kernel
# launch list of one eager step, per-shape convolution tables.
TAG=${1:-r02z}
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest -m gpu rc=$?"; tail -4 $O/${TAG}_pytest.log | cut -c1-200
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -2 $O/${TAG}_bench.err
timeout 900 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "bench reference rc=$?"
timeout 400 python bench.py --workload warp --no-cpu-baseline > $O/${TAG}_bench_warp.json 2> $O/${TAG}_bench_warp.err; echo "warp rc=$?"
FFWM_BENCH_FLOW=smooth timeout 400 python bench.py --workload warp --no-cpu-baseline > $O/${TAG}_bench_warp_smooth.json 2>> $O/${TAG}_bench_warp.err; echo "warp smooth rc=$?"
timeout 400 python bench.py --workload flownet > $O/${TAG}_bench_flownet.json 2> $O/${TAG}_bench_flownet.err; echo "flownet rc=$?"
FFWM_CONV_MATH_FWD=1 timeout 400 python bench.py --no-cpu-baseline --no-warp --no-library-baseline > $O/${TAG}_bench_fwdbf16.json 2> /dev/null; echo "bench (3xBF16 forwards) rc=$?"
timeout 300 python -m benchmarks.conv --out $O/${TAG}_conv.json > $O/${TAG}_conv.txt 2>&1
timeout 300 python -m benchmarks.conv --gen --out $O/${TAG}_conv_gen.json > $O/${TAG}_conv_gen.txt 2>&1
timeout 300 python -m benchmarks.conv --wgrad --out $O/${TAG}_conv_wgrad.json > $O/${TAG}_conv_wgrad.txt 2>&1
timeout 300 python benchmarks/batch_norm.py > $O/${TAG}_batch_norm.txt 2>&1
for sw in FFWM_FUSED_BN=0 FFWM_FUSED_ADAM=0 FFWM_FUSED_SN=0 FFWM_FUSED_POOL=0 FFWM_BATCHED_L1=0 FFWM_LOSSNET_MATH_FWD=0 FFWM_CONV_OCC2=1 FFWM_WGRAD_NO_ROWS=1; do
    env $sw timeout 400 python bench.py --no-cpu-baseline --no-warp --no-library-baseline --no-e2e > $O/${TAG}_bench_${sw%%=*}_off.json 2> /dev/null; echo "bench $sw rc=$?"
done
# compute-sanitizer over the kernels added since the last sanitizer pass (batch norm / channel sum, weight packing, wgrad reduction)
{
for tool in memcheck racecheck; do
    echo "== batch_norm_$tool"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_batch_norm_gpu.py -q -x \
        -k "2-5-7-9 or 3-4-2-2 or 8-195-32-32 or 2-6-16-16 or 7-3-20-28 or channel_sum or bit_identical" 2>&1 | grep -E "SUMMARY|passed|failed" | tail -2
done
for tool in memcheck racecheck; do
    echo "== spectral_pool_$tool"
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests/test_spectral_gpu.py tests/test_pool_gpu.py -q -x \
        -k "discriminator or declines or 3-5-7-9 or 2-3-8-8 or modules_use" 2>&1 | grep -E "SUMMARY|passed|failed" | tail -2
done
echo "== convgen_pack_wgrad_memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_conv_gen_gpu.py -q -x \
    -k "modules_autograd or deterministic or strided_views or (wgrad and (96-200-32 or 120-40-18 or 130-24-17))" 2>&1 | grep -E "SUMMARY|passed|failed" | tail -2
} > $O/${TAG}_sanitizer_summary.txt 2>&1; cat $O/${TAG}_sanitizer_summary.txt
FFWM_BENCH_GRAPH=0 FFWM_BENCH_NCU_RANGE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv --log-file $O/launches_train_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-warp --no-library-baseline > $O/ncu_launches_train_$TAG.log 2>&1; echo "ncu train launches rc=$?"
[[ -f $O/launches_train_$TAG.csv ]] && python scripts/launch_summary.py $O/launches_train_$TAG.csv $O/${TAG}_launches_train_summary.txt --rm
python - $TAG <<'PY'
import json, sys
t = sys.argv[1]
for f in ("bench", "bench_reference", "bench_warp", "bench_warp_smooth", "bench_flownet", "bench_fwdbf16", "bench_FFWM_FUSED_BN_off", "bench_FFWM_FUSED_ADAM_off", "bench_FFWM_FUSED_SN_off", "bench_FFWM_FUSED_POOL_off", "bench_FFWM_BATCHED_L1_off",
          "bench_FFWM_LOSSNET_MATH_FWD_off", "bench_FFWM_CONV_OCC2_off", "bench_FFWM_WGRAD_NO_ROWS_off"):
    try:
        d = json.loads(open("gpurun_out/%s_%s.json" % (t, f)).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 3), "ms/step", "e2e", d.get("e2e", {}).get("value"))
        if f == "bench":
            print("  roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "traffic", "whole_step_TFLOP/s")})
            print("  library", d.get("gpu_library_baseline"), d.get("measured_tensor_peaks"))
            print("  cpu", d.get("cpu_baseline"))
            print("  warp", {k: (v["ms"], v["frac_hbm"]) for k, v in d["warp_microbench"]["kernels"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
head -32 $O/${TAG}_launches_train_summary.txt | cut -c1-140
