#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for cfg in "8192 125000" "256 1000000"; do
set -- $cfg
GQ=$1 GN=$2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02x_gal_$1.csv python tools/gallery_prof.py > $O/r02x_gal_ncu_$1.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02x_gal_$1.csv")) if len(r)>10 and r[0].isdigit()]
print("Q=$1 N=$2")
for r in rows[-8:]:
    print("  ", r[4][:100], r[-1], r[7] if len(r)>7 else "")
PY
done
