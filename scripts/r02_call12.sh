#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_mfm_gpu.py tests/test_networks.py -m gpu -q > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02i_pytest.log | cut -c1-200
FFWM_BENCH_GRAPH=0 FFWM_BENCH_NCU_RANGE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv --log-file $O/launches_train_r02i.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-warp --no-library-baseline > $O/ncu_launches_train_r02i.log 2>&1; echo "ncu train launches rc=$?"
cp $O/launches_train_r02i.csv $O/launches_train_r02i_full.csv
python scripts/launch_summary.py $O/launches_train_r02i.csv $O/r02i_launches_train_summary.txt --rm
head -50 $O/r02i_launches_train_summary.txt | cut -c1-150
python - <<'PY'
# per-launch detail of the general kernel: grid sizes and durations
import csv, collections
rows = []
with open("gpurun_out/launches_train_r02i_full.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    k = r.get("Kernel Name", "")
    if "conv_gen" in k or "conv3x3" in k:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        key = (k[:40], r.get("Grid Size"), r.get("Block Size"))
        agg[key][0] += 1
        agg[key][1] += v
unit = "ns"
for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-42s grid %-18s n=%3d total %.3f ms  avg %.1f us" % (key[0], key[1], n, t / 1e6, t / n / 1e3))
PY
rm -f $O/launches_train_r02i_full.csv
