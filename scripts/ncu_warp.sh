#!/bin/bash
# ncu evidence for the warp microbench: launch list (time per launch) + one --set full capture of every warp kernel,
# reduced on the box to text summaries (gpurun_out is capped at 64 MiB).  Usage: bash scripts/ncu_warp.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
KR='regex:resample2d|block_extractor|lar_tiled|local_attn|grid_warp|scatter_tiled|scatter_rows|gather_quad'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_warp_$TAG.csv \
    python bench.py --workload warp --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k "$KR" -s 10 -c 10 -f -o gpurun_out/prof_$TAG \
    python bench.py --workload warp --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep $TAG gpurun_out > /dev/null 2>&1
for k in resample2d_fwd_roll gather_quad scatter_rows block_extractor_fwd block_extractor_bwd grid_warp_fwd; do python scripts/ncu_hot.py gpurun_out/prof_$TAG.ncu-rep $k 0x400 >> gpurun_out/${TAG}_ncu_hot.txt 2>/dev/null; done
[[ -n "$KEEP_REP" ]] || rm -f gpurun_out/prof_$TAG.ncu-rep
ls -la gpurun_out | tail -12
