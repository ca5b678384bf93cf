#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (train step + warp microbench), cfg5 sweep, ncu.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag] [skip-list]
TAG=${1:-r01}
SKIP=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi_$TAG.csv
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 1200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
[[ $SKIP == *tf32* ]] || { FFWM_BENCH_TF32=1 timeout 600 python bench.py --no-cpu-baseline --no-warp > gpurun_out/bench_tf32_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "bench tf32 rc=$?"; cat gpurun_out/bench_tf32_$TAG.json; }
[[ $SKIP == *warpbench* ]] || { timeout 600 python bench.py --workload warp > gpurun_out/bench_warp_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "bench warp rc=$?"; }
[[ $SKIP == *sweep* ]] || { timeout 1200 python -m benchmarks.sweep --out gpurun_out/sweep_$TAG.json > gpurun_out/sweep_$TAG.txt 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep_$TAG.txt; }
KR='regex:resample2d|block_extractor|lar_tiled|local_attn|grid_warp|scatter_tiled|scatter_rows|gather_quad'
[[ $SKIP == *ncu* ]] || {
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_warp_$TAG.csv \
    python bench.py --workload warp --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k "$KR" -s 10 -c 10 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload warp --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
# reduce the report on the box (gpurun_out is capped at 64 MiB) and drop the .ncu-rep unless KEEP_REP=1
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep $TAG gpurun_out > /dev/null 2>&1
for k in resample2d_fwd_roll gather_quad scatter_rows block_extractor_fwd block_extractor_bwd; do python scripts/ncu_hot.py gpurun_out/prof_$TAG.ncu-rep $k 0x400 >> gpurun_out/${TAG}_ncu_hot.txt 2>/dev/null; done
[[ -n "$KEEP_REP" ]] || rm -f gpurun_out/prof_$TAG.ncu-rep
[[ $SKIP == *trainlaunch* ]] || FFWM_BENCH_GRAPH=0 FFWM_BENCH_NCU_RANGE=1 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv --log-file gpurun_out/launches_train_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-warp > gpurun_out/ncu_launches_train_$TAG.log 2>&1; echo "ncu train launches rc=$?"
}
[[ $SKIP == *conv* ]] || { timeout 300 python -m benchmarks.conv --out gpurun_out/conv_$TAG.json > gpurun_out/conv_$TAG.txt 2>&1; echo "conv bench rc=$?"; cat gpurun_out/conv_$TAG.txt;
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 6 -c 6 -f -o gpurun_out/prof_conv_$TAG python -m benchmarks.conv --out gpurun_out/conv_ncu_tmp.json > gpurun_out/ncu_conv_$TAG.log 2>&1; echo "ncu conv rc=$?";
python scripts/ncu_summary.py gpurun_out/prof_conv_$TAG.ncu-rep ${TAG}_conv gpurun_out > /dev/null 2>&1; python scripts/ncu_hot.py gpurun_out/prof_conv_$TAG.ncu-rep conv3x3_tc 0x400 > gpurun_out/${TAG}_conv_ncu_hot.txt 2>/dev/null
[[ -n "$KEEP_REP" ]] || rm -f gpurun_out/prof_conv_$TAG.ncu-rep; rm -f gpurun_out/conv_ncu_tmp.json; }
# the train-step launch list is large: keep the per-kernel aggregate only
[[ -f gpurun_out/launches_train_$TAG.csv ]] && python scripts/launch_summary.py gpurun_out/launches_train_$TAG.csv gpurun_out/${TAG}_launches_train_summary.txt --rm
ls -la gpurun_out | tail -30
