#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for s in 0 28 24 32 36 0 28; do
B200F_BWD_SPLIT=$s timeout 300 python bench.py $B > $O/r02ar_bench_$s.json 2> $O/r02ar_bench.err; python - "$s" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/r02ar_bench_{t}.json").read().strip().splitlines()[-1])
    print("split", t, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss"], "parity", d.get("parity",{}).get("dx_rel"), d.get("parity",{}).get("dw_rel"), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("split", t, "failed", e); print(open("gpurun_out/r02ar_bench.err").read()[-1500:])
PY
done
B200F_BWD_SPLIT=28 timeout 1500 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_tail.py -m gpu -x -q > $O/r02ar_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ar_pytest.log | cut -c1-300
