#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
export B200FACE_LIB=$PWD/tools/build_tl/libb200face_tl.so
for v in "k3a epi_groups=1" "k3a epi_groups=2" "k3a epi_groups=4" "k2" "k3b"; do
  n=$(echo $v | tr ' =' '__')
  timeout 200 python tools/timeline_probe.py $v > $O/r02af_tl_$n.log 2>&1; echo "$v rc=$?"
done
tail -30 $O/r02af_tl_k3a_epi_groups_1.log
