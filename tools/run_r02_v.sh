#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
GTIME=1 GTUNE="gallery_compact=1" timeout 120 python tools/gallery_prof.py > $O/r02v_gal_plain.log 2>&1; tail -1 $O/r02v_gal_plain.log
cat > /tmp/galc.py <<'PY'
import os, sys
sys.path.insert(0, "/root/repo")
import torch, b200face
from b200face import _lib
b200face.load_library().b200f_set_tunable(b"gallery_compact", 1)
exec(open("/root/repo/tools/gallery_prof.py").read().split("import b200face\nfrom b200face import _lib\n")[1])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02v_gal_launches.csv python /tmp/galc.py > $O/r02v_gal_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02v_gal_launches.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows[-9:]:
    print(r[4][:100], r[-1])
PY
