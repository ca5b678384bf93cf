#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_input_pipeline.py -q > $O/r02m_pytest.log 2>&1; echo "input pipeline pytest rc=$?"; tail -3 $O/r02m_pytest.log | cut -c1-200
timeout 300 python -m benchmarks.fused_losses --out $O/r02m_fused_losses.json > $O/r02m_fused_losses.txt 2>&1; echo "fused losses bench rc=$?"; cat $O/r02m_fused_losses.txt | cut -c1-400
# ---- compute-sanitizer over small-shape parity tests (each pass bounded)
san() {  # name, tool, pytest args...
    local name=$1 tool=$2; shift 2
    timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest "$@" -q -x > $O/sanitize_${name}_$tool.log 2>&1
    echo "sanitize $name $tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $O/sanitize_${name}_$tool.log | tail -2
}
san warp memcheck tests/test_warp_gpu.py -k "vs_oracle"
san warp racecheck tests/test_warp_gpu.py -k "vs_oracle and not reference_cuda"
san convgen memcheck tests/test_conv_gen_gpu.py -k "1-5-7-9-11 or 3-17-33 or 2-12-20-7-9 or 1-16-16-1-1 or strided_views"
san conv3x3 memcheck tests/test_conv_tc_gpu.py -k "dgrad_packing or rejects"
san fused memcheck tests/test_fused_losses.py tests/test_input_pipeline.py tests/test_mfm_gpu.py -k "not full_size"
san fused racecheck tests/test_fused_losses.py -k "affine_reg_ragged or 64-19"
# ---- ncu --set full of the tensor-core kernels (one capture each, summaries reduced on the box)
prof() {  # tag, regex, skip, count, command...
    local tag=$1 rx=$2 skip=$3 cnt=$4; shift 4
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o $O/prof_$tag "$@" > $O/ncu_$tag.log 2>&1; echo "ncu $tag rc=$?"
    python scripts/ncu_summary.py $O/prof_$tag.ncu-rep r02m_$tag $O > /dev/null 2>&1
    rm -f $O/prof_$tag.ncu-rep
}
prof conv3x3 conv3x3_tc_kernel 3 4 python -m benchmarks.conv --out $O/tmp_conv.json
prof convgen conv_gen_tc_kernel 0 8 python -m benchmarks.conv --gen --out $O/tmp_gen.json
prof wgradgen conv_gen_wgrad_tc_kernel 0 3 python -m benchmarks.conv --wgrad --out $O/tmp_wg.json
prof corrmax corr_max_kernel 0 2 python -m benchmarks.fused_losses --out $O/tmp_fl.json
rm -f $O/tmp_*.json
ls $O | grep r02m
