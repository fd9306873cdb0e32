#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_adamw.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02t_pytest.log | cut -c1-400
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
timeout 300 python bench.py $B > $O/r02t_bench.json 2> $O/r02t_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02t_bench.json").read().strip().splitlines()[-1])
k=d["kernel_ms"]
print("ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"])
PY
