#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for v in 0 1 2 3 0 3; do
  timeout 300 python bench.py $B --tune k1_hints=$v > $O/r02o_h$v.json 2> $O/r02o_h$v.err || { echo "h$v failed"; tail -5 $O/r02o_h$v.err; }
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02o_h$v.json").read().strip().splitlines()[-1])
k=d["kernel_ms"]
print("k1_hints", $v, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"])
PY
done
