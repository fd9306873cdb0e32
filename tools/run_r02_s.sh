#!/bin/bash
# N = 8: the bench line as the driver launches it (defaults), with and without the all-reduce overlap (short)
set -u
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 > $O/r02s_bench8.json 2> $O/r02s_bench8.err; echo "bench8 rc=$? wall=$(( $(date +%s) - T0 ))s"; tail -2 $O/r02s_bench8.err | cut -c1-300
for ov in 0 1; do
B200F_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 300 --warmup 5 --no-cpu-baseline --no-gallery > $O/r02s_bench8_ov$ov.json 2> $O/r02s_bench8_ov$ov.err; echo "bench8 ov$ov rc=$?"
done
python - <<'PY'
import json
for f in ("r02s_bench8","r02s_bench8_ov0","r02s_bench8_ov1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "burst", (d.get("burst") or {}).get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "eff", d.get("efficiency_vs_cfg4_1gpu"), d.get("efficiency_vs_cfg4_1gpu_sustained"), "anchor", d.get("anchor_1gpu_ms"), "clocks", d.get("clocks"))
        print("   parity", json.dumps(d.get("parity"))[:300])
        print("   kernel_ms", d.get("kernel_ms"), "gallery", json.dumps(d.get("gallery"))[:400])
    except Exception as e:
        print(f, "no line", e)
PY
