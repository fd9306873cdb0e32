#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02q_pytest.log | cut -c1-300
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for v in 1 0 1 0; do
  timeout 300 python bench.py $B --tune x_whole=$v > $O/r02q_x$v.json 2> $O/r02q_x$v.err || { echo "x$v failed"; tail -5 $O/r02q_x$v.err; }
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02q_x$v.json").read().strip().splitlines()[-1])
k=d["kernel_ms"]
print("x_whole", $v, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"])
PY
done
