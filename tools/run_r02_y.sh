#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_gallery_tc.py tests/test_gpu_gallery.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02y_gal.log 2>&1; echo "gallery pytest rc=$?"; tail -4 $O/r02y_gal.log | cut -c1-300
GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=256 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=1024 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=8192 GN=125000 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=8192 GN=125000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02y_gal_8192.csv python tools/gallery_prof.py > $O/r02y_gal_ncu_8192.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02y_gal_8192.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]:
    print("  ", r[4][:100], r[-1])
PY
