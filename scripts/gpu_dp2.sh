#!/bin/bash
# 2-GPU data-parallel diagnostics: phases of the segmented step, N=1 one-graph vs N=1 segmented vs N=2
cd "$(dirname "$0")/.."; O=gpurun_out; mkdir -p $O; TAG=${1:-dp2}
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-warp --no-e2e"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("%-28s %.2f ms  %.1f img/s" % (sys.argv[2], d["ms_per_step"], d["value"]))
except Exception as e: print(sys.argv[2], "failed", e)
PY
}
timeout 600 python $B > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err; show $O/${TAG}_n1.json "N=1 one graph"
FFWM_BENCH_SEGMENTED=1 FFWM_BENCH_PHASES=1 timeout 600 python $B > $O/${TAG}_n1seg.json 2> $O/${TAG}_n1seg.err; show $O/${TAG}_n1seg.json "N=1 three graphs"; grep phases $O/${TAG}_n1seg.err
FFWM_BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 $B --gpus 2 > $O/${TAG}_n2.json 2> $O/${TAG}_n2.err; show $O/${TAG}_n2.json "N=2"; grep phases $O/${TAG}_n2.err
NCCL_DEBUG=INFO FFWM_BENCH_PHASES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 $B --gpus 2 --steps 5 > $O/${TAG}_n2dbg.json 2> $O/${TAG}_n2dbg.err; grep -i "channels\|NVLS\|via P2P\|Connected\|Algo\|nranks" $O/${TAG}_n2dbg.err | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dp_check.py 2>&1 | grep -v Warn | tail -5
