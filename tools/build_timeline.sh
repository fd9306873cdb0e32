#!/bin/bash
# Instrumented build of the same sources (-DB200F_TIMELINE: clock stamps per tile in the X-stationary kernel, XW_TL) for
# tools/timeline_probe.py.  Output: tools/build_tl/libb200face_tl.so (not the product library; loaded via B200FACE_LIB).
set -e
cd "$(dirname "$0")/.."
S=facerecognition-multiarchitecture-pipeline_b200/csrc
# usage: build_timeline.sh [suffix extra-nvcc-flags...]   e.g.  build_timeline.sh _ring3 -DB200F_TL_RING=3
SUF=${1:-}; shift || true
O=tools/build_tl$SUF
mkdir -p $O
for f in $S/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DB200F_TIMELINE "$@" -c $f -o $O/$(basename $f).o &
done
wait
nvcc -shared -o $O/libb200face_tl$SUF.so $O/*.o -gencode arch=compute_100a,code=sm_100a -ldl
ls -la $O/libb200face_tl$SUF.so
