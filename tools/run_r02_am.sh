#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02am_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02am_pytest.log | cut -c1-300
B="--steps 200 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for i in 1 2 3; do
for t in "early=1" "early=2"; do
timeout 300 python bench.py $B --tune $t > $O/r02am_bench_$t.json 2> $O/r02am_bench.err; python - "$t" <<'PY'
import json,sys
t=sys.argv[1]
d=json.loads(open(f"gpurun_out/r02am_bench_{t}.json").read().strip().splitlines()[-1])
k=d["kernel_ms"]
print(t, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("k2","k3a","k3b","k3c")), "loss", d["loss"], "parity", d.get("parity",{}).get("dx_rel"), d.get("parity",{}).get("dw_rel"), "clk", d["clocks"]["sm_mhz"])
PY
done; done
