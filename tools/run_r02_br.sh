#!/bin/bash
# per-tile timelines of the X-stationary kernels on the final tree (instrumented build of the same sources)
set -u
mkdir -p gpurun_out
O=gpurun_out
export B200FACE_LIB=$PWD/tools/build_tl/libb200face_tl.so
for v in k2 k3a k3b; do timeout 200 python tools/timeline_probe.py $v > $O/r02br_timeline_$v.txt 2>&1; echo "timeline $v rc=$?"; done
