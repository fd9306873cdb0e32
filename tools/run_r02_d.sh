#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_umma.py tests/test_gpu_adamw.py -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02d_pytest.log
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4 --tune xw_prefetch=0"
for v in "eg=1 early=0" "eg=1 early=1" "eg=2 early=0" "eg=2 early=1"; do
  set -- $v; eg=${1#eg=}; ea=${2#early=}
  n="eg${eg}_early${ea}"
  timeout 300 python bench.py $B --tune epi_groups=$eg --tune early=$ea > $O/r02d_$n.json 2> $O/r02d_$n.err || { echo "$n failed"; tail -5 $O/r02d_$n.err; }
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d_eg*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms"]
        print(f.split("r02d_")[1][:-5], "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches"]//d["steps"],
              "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"], d.get("e2e_loss"))
    except Exception as e:
        print(f, "no line", e)
PY
