#!/bin/bash
# full default bench line (N = 1) + reference arm + gallery launch list
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 300 --warmup 5 > $O/r02g_bench.json 2> $O/r02g_bench.err; echo "bench rc=$?"; tail -3 $O/r02g_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02g_ref.json 2> $O/r02g_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","burst","parity","cpu_baseline","cfg1_cpu","roofline","loss","e2e_loss","gpu_launches"):
    print(k, json.dumps(d.get(k))[:600])
print("gallery", json.dumps(d.get("gallery"))[:3000])
print("cfg4_single", d.get("cfg4_single_gpu"))
print(open("gpurun_out/r02g_ref.json").read()[:800])
PY
GTIME=1 timeout 120 python tools/gallery_prof.py > $O/r02g_gal_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02g_gal_launches.csv python tools/gallery_prof.py > $O/r02g_gal_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r02g_gal_plain.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02g_gal_launches.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:90], r[-1])
PY
