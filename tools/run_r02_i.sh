#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_tail.py -m gpu -x -q > $O/r02i_tail.log 2>&1; echo "tail pytest rc=$?"; tail -12 $O/r02i_tail.log
timeout 900 python -m pytest tests/test_gpu_gallery_tc.py tests/test_gpu_gallery.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02i_gal.log 2>&1; echo "gallery pytest rc=$?"; tail -12 $O/r02i_gal.log
GTIME=1 GTUNE="gallery_compact=0;gallery_compact=1" timeout 300 python tools/gallery_prof.py 2>&1 | tail -4
GQ=64 GTIME=1 GTUNE="gallery_compact=0;gallery_compact=1" timeout 300 python tools/gallery_prof.py 2>&1 | tail -2
