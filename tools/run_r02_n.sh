#!/bin/bash
# 2 GPUs: NCCL multi-rank parity test, then the cfg4 bench line with and without the all-reduce overlap
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $O/r02n_pytest.log 2>&1; echo "multirank pytest rc=$?"; tail -6 $O/r02n_pytest.log | cut -c1-400
for ov in 1 0; do
  B200F_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 200 --warmup 5 --no-cpu-baseline --no-gallery > $O/r02n_bench_ov$ov.json 2> $O/r02n_bench_ov$ov.err; echo "bench ov=$ov rc=$?"; tail -2 $O/r02n_bench_ov$ov.err | cut -c1-300
done
python - <<'PY'
import json
for ov in (1,0):
    try:
        d=json.loads(open(f"gpurun_out/r02n_bench_ov{ov}.json").read().strip().splitlines()[-1])
        print("overlap", ov, "ms/step", d["ms_per_step"], "value", d["value"], "burst", d.get("burst"), "eff", d.get("efficiency_vs_cfg4_1gpu"), "parity", json.dumps(d.get("parity"))[:400], "kernel_ms", d.get("kernel_ms"))
    except Exception as e:
        print(ov, "no line", e)
PY
