#!/bin/bash
# gallery scan: biases staged through shared memory one tile ahead -- gallery tests, per-call times at Q = 128 / 256 / 1024 / 8192
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_gallery.py tests/test_gpu_gallery_tc.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/r02bo_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02bo_pytest.log | cut -c1-200
GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=256 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=1024 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
GQ=8192 GN=125000 GTIME=1 timeout 300 python tools/gallery_prof.py 2>&1 | tail -1
