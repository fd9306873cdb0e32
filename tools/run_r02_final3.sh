#!/bin/bash
# last tree of the round: full GPU suite, smoke, the bench line in the driver's short form
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02bm_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02bm_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02bm_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02bm_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02bm_bench20.json 2> $O/r02bm_bench20.err; echo "bench20 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02bm_bench20.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","parity","kernel_ms","cfg4_single_gpu","clocks"):
    print(" ", k, json.dumps(d.get(k))[:300])
PY
