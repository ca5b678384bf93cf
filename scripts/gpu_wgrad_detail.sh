#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
FFWM_BENCH_GRAPH=0 FFWM_BENCH_NCU_RANGE=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"conv_gen_wgrad|conv_gen_tc" -c 40000 --csv --log-file $O/launches_wg.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-warp --no-library-baseline > $O/ncu_launches_wg.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
lines = [l for l in open("gpurun_out/launches_wg.csv") if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    k = r.get("Kernel Name", "")
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    key = (k[:34], r.get("Grid Size"))
    agg[key][0] += 1
    agg[key][1] += v
tot = collections.defaultdict(float)
for (k, g), (n, t) in agg.items():
    tot[k] += t
print({k: round(v / 1e6, 3) for k, v in tot.items()})
for key, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%-36s grid %-16s n=%3d total %.3f ms  avg %.1f us" % (key[0], key[1], n, t / 1e6, t / n / 1e3))
PY
rm -f $O/launches_wg.csv
