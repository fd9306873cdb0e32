#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_tail.py -m gpu -x -q > $O/r02an_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02an_pytest.log | cut -c1-300
B="--steps 200 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for i in 1 2; do
timeout 300 python bench.py $B > $O/r02an_bench.json 2> $O/r02an_bench.err; python - <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r02an_bench.json").read().strip().splitlines()[-1])
k=d["kernel_ms"]
print("ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("k2","k3a","k3b","k3c")), "loss", d["loss"], "clk", d["clocks"]["sm_mhz"])
PY
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02an_launches_raw.csv $CMD > $O/r02an_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
grep -c . $O/r02an_launches_raw.csv; grep "reduce_splits\|reduce_row" $O/r02an_launches_raw.csv | tail -4 | cut -c1-200
