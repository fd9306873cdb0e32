#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02f_pytest.log
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for v in 1 2; do
  timeout 300 python bench.py $B --tune k3b_groups=$v > $O/r02f_k3bg$v.json 2> $O/r02f_k3bg$v.err || { echo "k3bg$v failed"; tail -5 $O/r02f_k3bg$v.err; }
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_k3bg*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms"]
        print(f.split("r02f_")[1][:-5], "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"],
              "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"])
    except Exception as e:
        print(f, "no line", e)
PY
