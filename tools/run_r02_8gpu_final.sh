#!/bin/bash
# 8 GPUs, final tree: the default cfg4 bench line as the driver launches it
set -u
mkdir -p gpurun_out
O=gpurun_out
T0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 > $O/r02bh_bench8.json 2> $O/r02bh_bench8.err; echo "bench8 rc=$? wall=$(( $(date +%s) - T0 ))s"; tail -2 $O/r02bh_bench8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02bh_bench8.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","burst","anchor_1gpu_ms","efficiency_vs_cfg4_1gpu","parity","kernel_ms","clocks","roofline"):
    print(k, json.dumps(d.get(k))[:400])
print("gallery", json.dumps(d.get("gallery"))[:500])
PY
