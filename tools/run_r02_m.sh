#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for v in 0 1; do
  timeout 300 python bench.py $B --tune k2_prep=$v > $O/r02m_prep$v.json 2> $O/r02m_prep$v.err || { echo "prep$v failed"; tail -5 $O/r02m_prep$v.err; }
done
python - <<'PY'
import json
for v in (0,1):
    try:
        d=json.loads(open(f"gpurun_out/r02m_prep{v}.json").read().strip().splitlines()[-1])
        k=d["kernel_ms"]
        print("k2_prep", v, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"])
    except Exception as e:
        print(v, "no line", e)
PY
timeout 600 python -m pytest tests/test_gpu_head.py -m gpu -x -q -k "k1w_inside" > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest.log | cut -c1-300
