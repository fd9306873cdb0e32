#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4 --tune xw_prefetch=0 --tune epi_groups=2"
for h in 0 4 2 6 7 14 22 38 47 63; do
  timeout 300 python bench.py $B --tune l2_hints=$h > $O/r02e_h$h.json 2> $O/r02e_h$h.err || { echo "h$h failed"; tail -5 $O/r02e_h$h.err; }
done
python - <<'PY'
import json,glob
for h in (0,4,2,6,7,14,22,38,47,63):
    f=f"gpurun_out/r02e_h{h}.json"
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernel_ms"]
        print("hints", h, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "e2e", d["e2e"]["value"],
              "k1w %.1f k2 %.1f k3a %.1f k3b %.1f k3c %.1f" % tuple(1e3*k[x] for x in ("l2norm_rows_w","k2","k3a","k3b","k3c")), "loss", d["loss"])
    except Exception as e:
        print(f, "no line", e)
PY
