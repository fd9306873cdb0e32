#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_adamw.py tests/test_gpu_tail.py -m gpu -x -q > $O/r02ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02ab_pytest.log | cut -c1-400
timeout 300 python tools/careful_probe.py 2>&1 | tail -2
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
timeout 300 python bench.py $B > $O/r02ab_bench.json 2> $O/r02ab_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02ab_bench.json").read().strip().splitlines()[-1])
k=d["kernel_ms"]
print("ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"], "parity", d.get("parity",{}).get("dx_rel"), d.get("parity",{}).get("dw_rel"), d.get("parity",{}).get("loss_rel"))
PY
