#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02at_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02at_pytest.log | cut -c1-300
timeout 900 python bench.py > $O/r02at_bench.json 2> $O/r02at_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02at_bench20.json 2> $O/r02at_bench20.err; echo "bench20 rc=$?"
python - <<'PY'
import json
for f in ("r02at_bench","r02at_bench20"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clk", d["clocks"], "parity", d["parity"]["dx_rel"], d["parity"]["dw_rel"], d["parity"]["loss_rel"])
    print("   roofline", json.dumps(d["roofline"])[:300])
PY
