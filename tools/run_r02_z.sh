#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_gallery_tc.py tests/test_gpu_gallery.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02z_gal.log 2>&1; echo "gallery pytest rc=$?"; tail -3 $O/r02z_gal.log | cut -c1-300
GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=8192 GN=125000 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-train-step --no-cfg4 > $O/r02z_bench.json 2> $O/r02z_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z_bench.json").read().strip().splitlines()[-1])
g=d.get("gallery",{})
for k,v in g.items():
    if isinstance(v,dict): print("gallery",k,{kk:v[kk] for kk in v if kk in("ms","frac_of_hbm_peak","queries_per_sec","e2e_queries_per_sec","redo_per_call")}, (v.get("pipelined") or {}).get("frac_of_hbm_peak"))
PY
