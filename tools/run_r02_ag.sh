#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
export B200FACE_LIB=$PWD/tools/build_tl/libb200face_tl.so
i=0
for v in "k2" "k2 xw_prefetch=2" "k2 xw_prefetch=4" "k2 C=40000 noflush=1" "k2 C=40000" "k2 pair=1" "k3a epi_groups=1 xw_prefetch=2" "k2 B=256" ; do
  i=$((i+1))
  timeout 200 python tools/timeline_probe.py $v > $O/r02ag_tl_$i.log 2>&1; echo "$i: $v rc=$?"
  echo "== $v" >> $O/r02ag_all.log; head -22 $O/r02ag_tl_$i.log >> $O/r02ag_all.log
done
