#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/r02ah_all.log
run() { # name lib env... -- args
  echo "== $1" >> $O/r02ah_all.log
  timeout 200 env B200FACE_LIB=$PWD/$2 $3 python tools/timeline_probe.py $4 2>&1 | grep -v "^tile |" | head -34 >> $O/r02ah_all.log
}
run "k2 5 stages" tools/build_tl/libb200face_tl.so X=1 "k2"
run "k2 3 stages" tools/build_tl_ring3/libb200face_tl_ring3.so X=1 "k2"
run "k2 5 stages, 37 clusters" tools/build_tl/libb200face_tl.so B200F_TL_CLUSTERS=37 "k2"
run "k2 5 stages, 18 clusters" tools/build_tl/libb200face_tl.so B200F_TL_CLUSTERS=18 "k2"
run "k2 5 stages, 8 clusters" tools/build_tl/libb200face_tl.so B200F_TL_CLUSTERS=8 "k2"
run "k2 3 stages, 18 clusters" tools/build_tl_ring3/libb200face_tl_ring3.so B200F_TL_CLUSTERS=18 "k2"
