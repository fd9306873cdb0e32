#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 200 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
run() { # split extra
B200F_BWD_SPLIT=$1 timeout 300 python bench.py $B $2 > $O/r02as_bench.json 2> $O/r02as_bench.err; python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r02as_bench.json").read().strip().splitlines()[-1])
    print("split", sys.argv[1], sys.argv[2], "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "loss", d["loss"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r02as_bench.err").read()[-1500:])
PY
}
run 0 ""
run 24 ""
run 20 ""
run 16 ""
run 12 ""
run 24 "--tune k3b_reverse=0"
run 20 "--tune k3b_reverse=0"
run 0 ""
run 20 ""
