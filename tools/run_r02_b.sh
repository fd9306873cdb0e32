#!/bin/bash
# Round-2 call B: tests of the fused small kernels + sweep of the L2 tile prefetch distance and epilogue groups.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r02b_pytest.log
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -E "^cfg3|^.cfg4|epi_groups|passed|failed" | tail -8
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
for v in "eg=1 pf=0" "eg=2 pf=0" "eg=1 pf=2" "eg=2 pf=1" "eg=2 pf=2" "eg=2 pf=3" "eg=2 pf=5"; do
  set -- $v; eg=${1#eg=}; pf=${2#pf=}
  n="eg${eg}_pf${pf}"
  timeout 300 python bench.py $B --tune epi_groups=$eg --tune xw_prefetch=$pf > $O/r02b_$n.json 2> $O/r02b_$n.err || { echo "$n failed"; tail -5 $O/r02b_$n.err; }
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02b_eg*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms"]
        print(f.split("r02b_")[1][:-5], "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches"]//d["steps"],
              "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"])
    except Exception as e:
        print(f, "no line", e)
PY
