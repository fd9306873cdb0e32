#!/bin/bash
# gallery: parity tests + timing of the streaming and the large-Q regime
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_gallery_tc.py tests/test_gpu_gallery.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02p_gal.log 2>&1; echo "gallery pytest rc=$?"; tail -4 $O/r02p_gal.log | cut -c1-300
GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -2
GQ=256 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=1024 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=8192 GN=125000 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
