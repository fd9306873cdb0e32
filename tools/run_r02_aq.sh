#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
export B200FACE_LIB=$PWD/tools/build_tl/libb200face_tl.so
for n in 74 37 44 30; do
  echo "== clusters $n" >> $O/r02aq.log
  B200F_TL_CLUSTERS=$n timeout 200 python tools/ab_probe.py early=2 2>&1 | tail -3 | head -2 | cut -c1-200 >> $O/r02aq.log
done
cat $O/r02aq.log
unset B200FACE_LIB
timeout 600 python -m pytest tests/test_gpu_gallery_tc.py tests/test_gpu_gallery.py -m gpu -x -q 2>&1 | tail -3
