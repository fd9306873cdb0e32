#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
GTIME=1 timeout 120 python tools/gallery_prof.py > $O/r02u_gal_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02u_gal_launches.csv python tools/gallery_prof.py > $O/r02u_gal_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r02u_gal_plain.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02u_gal_launches.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-9:]:
    print(r[4][:100], r[-1])
PY
