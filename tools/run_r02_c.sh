#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
HTUNE=xw_prefetch=0,epi_groups=1 timeout 600 python tools/k2_probe.py > $O/r02c_probe_eg1.log 2>&1; echo "probe eg1 rc=$?"; cat $O/r02c_probe_eg1.log
HTUNE=xw_prefetch=0,epi_groups=2 timeout 600 python tools/k2_probe.py > $O/r02c_probe_eg2.log 2>&1; echo "probe eg2 rc=$?"; cat $O/r02c_probe_eg2.log
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
timeout 300 python bench.py $B --tune epi_groups=1 --tune xw_prefetch=0 > $O/r02c_bench.json 2> $O/r02c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c_bench.json").read().strip().splitlines()[-1])
k=d["kernel_ms"]
print("ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")))
PY
