#!/bin/bash
# 2 GPUs, final tree: NCCL multi-rank parity tests, then the default cfg4 bench line as the driver launches it
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $O/r02bl_pytest.log 2>&1; echo "multirank pytest rc=$?"; tail -4 $O/r02bl_pytest.log | cut -c1-400
T0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > $O/r02bl_bench2.json 2> $O/r02bl_bench2.err; echo "bench2 rc=$? wall=$(( $(date +%s) - T0 ))s"; tail -2 $O/r02bl_bench2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02bl_bench2.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","burst","anchor_1gpu_ms","efficiency_vs_cfg4_1gpu","parity","kernel_ms","clocks","roofline"):
    print(k, json.dumps(d.get(k))[:500])
print("gallery", json.dumps(d.get("gallery"))[:600])
PY
