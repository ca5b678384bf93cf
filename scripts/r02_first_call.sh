#!/bin/bash
# First GPU call of round 2: validate and measure what was written blind at the end of round 1
# (the tcgen05 weight-gradient kernel, the 128-channel CTA tile of the forward kernel, the fused MFM and guided-filter kernels), without touching the established suite's context.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/r02_first_call.sh'     (about 30 minutes of box time; comment out what is not needed)
# Each step runs in its own process under its own timeout: a trap in the unproven kernel ends that step only.
mkdir -p gpurun_out
FFWM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_wgrad_tc_gpu.py -x -q > gpurun_out/r02_wgrad_pytest.log 2>&1; echo "wgrad pytest rc=$?"
tail -25 gpurun_out/r02_wgrad_pytest.log
# fallback / A-B: REDs straight from the TMEM registers instead of the shared-memory transposition
FFWM_WGRAD_DIRECT_EPILOGUE=1 FFWM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_wgrad_tc_gpu.py -x -q > gpurun_out/r02_wgrad_pytest_direct.log 2>&1; echo "wgrad pytest (direct epilogue) rc=$?"
tail -5 gpurun_out/r02_wgrad_pytest_direct.log
FFWM_WGRAD_DIRECT_EPILOGUE=1 timeout 300 python -m benchmarks.conv --wgrad --out gpurun_out/r02_conv_wgrad_direct.json > gpurun_out/r02_conv_wgrad_direct.txt 2>&1; echo "wgrad bench (direct epilogue) rc=$?"
tail -10 gpurun_out/r02_conv_wgrad_direct.txt
timeout 300 python -m benchmarks.conv --wgrad --out gpurun_out/r02_conv_wgrad.json > gpurun_out/r02_conv_wgrad.txt 2>&1; echo "wgrad bench rc=$?"
cat gpurun_out/r02_conv_wgrad.txt | tail -12
# the 128-channel CTA tile of the forward / data-gradient kernel (W = 128)
FFWM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_conv_nt128_gpu.py -x -q > gpurun_out/r02_nt128_pytest.log 2>&1; echo "nt128 pytest rc=$?"
tail -8 gpurun_out/r02_nt128_pytest.log
timeout 300 python -m benchmarks.conv --nt128 --out gpurun_out/r02_conv_nt128.json > gpurun_out/r02_conv_nt128.txt 2>&1; echo "nt128 bench rc=$?"
cat gpurun_out/r02_conv_nt128.txt | tail -6
# fused LightCNN max-feature-map kernels
FFWM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_mfm_gpu.py -x -q > gpurun_out/r02_mfm_pytest.log 2>&1; echo "mfm pytest rc=$?"
tail -8 gpurun_out/r02_mfm_pytest.log
# fused guided filter
FFWM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_guided_filter_gpu.py -x -q > gpurun_out/r02_gf_pytest.log 2>&1; echo "guided filter pytest rc=$?"
tail -8 gpurun_out/r02_gf_pytest.log
# A/B on the headline: the train step with cuDNN weight gradients (default) and with the tcgen05 ones
timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_cudnn_wgrad.json 2> gpurun_out/r02_bench_a.err; echo "bench (cuDNN wgrad) rc=$?"
FFWM_WGRAD_TC=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_tc_wgrad.json 2> gpurun_out/r02_bench_b.err; echo "bench (tcgen05 wgrad) rc=$?"
FFWM_CONV_NT128=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_nt128.json 2> gpurun_out/r02_bench_c.err; echo "bench (nt128) rc=$?"
FFWM_FUSED_MFM=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_mfm.json 2> gpurun_out/r02_bench_d.err; echo "bench (fused mfm) rc=$?"
FFWM_FUSED_GF=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_gf.json 2> gpurun_out/r02_bench_g.err; echo "bench (fused guided filter) rc=$?"
# batched spectral norm is plain torch ops, CPU-verified against the reference goldens: parity on the GPU, then the A/B
FFWM_BATCHED_SN=1 timeout 600 python -m pytest tests/test_networks.py tests/test_train_step.py -m gpu -x -q > gpurun_out/r02_sn_pytest.log 2>&1; echo "batched SN gpu goldens rc=$?"; tail -3 gpurun_out/r02_sn_pytest.log
FFWM_BATCHED_SN=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_sn.json 2> gpurun_out/r02_bench_h.err; echo "bench (batched SN) rc=$?"
# batched VGG19 / LightCNN loss passes (6 + 2 passes instead of 14 + 4): also plain torch, CPU-verified
FFWM_BATCHED_VGG=1 timeout 600 python -m pytest tests/test_train_step.py -m gpu -x -q > gpurun_out/r02_vgg_pytest.log 2>&1; echo "batched VGG gpu goldens rc=$?"; tail -3 gpurun_out/r02_vgg_pytest.log
FFWM_BATCHED_VGG=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_vgg.json 2> gpurun_out/r02_bench_i.err; echo "bench (batched VGG/LightCNN) rc=$?"
# flowNetF / flowNetB on two streams (fork-join inside the captured graph): parity first, then the A/B
FFWM_FLOW_STREAMS=1 timeout 600 python -m pytest tests/test_train_step.py -m gpu -x -q > gpurun_out/r02_streams_pytest.log 2>&1; echo "flow streams gpu goldens rc=$?"; tail -3 gpurun_out/r02_streams_pytest.log
FFWM_FLOW_STREAMS=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_streams.json 2> gpurun_out/r02_bench_m.err; echo "bench (flow streams) rc=$?"
FFWM_CACHE_PACKED=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_cache.json 2> gpurun_out/r02_bench_k.err; echo "bench (packed-weight cache) rc=$?"
FFWM_BATCHED_SN=1 FFWM_BATCHED_VGG=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_host.json 2> gpurun_out/r02_bench_j.err; echo "bench (both host-side restructurings) rc=$?"
FFWM_WGRAD_TC=1 FFWM_CONV_NT128=1 FFWM_FUSED_MFM=1 FFWM_FUSED_GF=1 FFWM_CACHE_PACKED=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_cache.json 2> gpurun_out/r02_bench_k.err; echo "bench (packed-weight cache) rc=$?"
FFWM_BATCHED_SN=1 FFWM_BATCHED_VGG=1 FFWM_CACHE_PACKED=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/r02_bench_all.json 2> gpurun_out/r02_bench_e.err; echo "bench (all seven) rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_cudnn_wgrad", "r02_bench_tc_wgrad", "r02_bench_nt128", "r02_bench_mfm", "r02_bench_gf", "r02_bench_sn", "r02_bench_vgg", "r02_bench_streams", "r02_bench_cache", "r02_bench_host", "r02_bench_all"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], "ms/step")
    except Exception as e:
        print(f, "unreadable:", e)
PY
# the warp microbench with a smooth (network-like) displacement field next to cfg5's independent per-pixel noise
timeout 600 python bench.py --workload warp --no-cpu-baseline > gpurun_out/r02_bench_warp_iid.json 2> gpurun_out/r02_bench_l.err; echo "bench warp (iid flow) rc=$?"
FFWM_BENCH_FLOW=smooth timeout 600 python bench.py --workload warp --no-cpu-baseline > gpurun_out/r02_bench_warp_smooth.json 2>> gpurun_out/r02_bench_l.err; echo "bench warp (smooth flow) rc=$?"
python - <<'PY'
import json
for f in ("r02_bench_warp_iid", "r02_bench_warp_smooth"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.0f GB/s" % d["value"], {k: (v["ms"], v["frac_hbm"]) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
# BASELINE config 2 (FlowNetF forward+backward, batch 6) was never recorded in round 1
timeout 600 python bench.py --workload flownet --no-cpu-baseline > gpurun_out/r02_bench_flownet.json 2> gpurun_out/r02_bench_f.err; echo "bench (flownet, cfg2) rc=$?"; cat gpurun_out/r02_bench_flownet.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad -s 2 -c 3 -f -o gpurun_out/prof_r02_wgrad \
    python -m benchmarks.conv --wgrad --out gpurun_out/conv_ncu_tmp.json > gpurun_out/r02_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_r02_wgrad.ncu-rep r02_wgrad gpurun_out > /dev/null 2>&1
python scripts/ncu_hot.py gpurun_out/prof_r02_wgrad.ncu-rep conv3x3_wgrad 0x400 > gpurun_out/r02_wgrad_ncu_hot.txt 2>/dev/null
[[ -n "$KEEP_REP" ]] || rm -f gpurun_out/prof_r02_wgrad.ncu-rep; rm -f gpurun_out/conv_ncu_tmp.json
ls -la gpurun_out | tail -12
