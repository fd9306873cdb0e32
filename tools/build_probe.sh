#!/bin/bash
# Probe build of the same sources (extra -D flags; results may be WRONG by design) -> tools/build_probe<suffix>/libb200face_probe<suffix>.so,
# loaded via B200FACE_LIB.  usage: build_probe.sh _gtblocked -DB200F_GT_BLOCKED_PROBE
set -e
cd "$(dirname "$0")/.."
S=facerecognition-multiarchitecture-pipeline_b200/csrc
SUF=${1:-}; shift || true
O=tools/build_probe$SUF
mkdir -p $O
for f in $S/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o $O/$(basename $f).o &
done
wait
nvcc -shared -o $O/libb200face_probe$SUF.so $O/*.o -gencode arch=compute_100a,code=sm_100a -ldl
ls -la $O/libb200face_probe$SUF.so
