#!/bin/bash
# ncu --set full of the gallery main scan at one rank's cfg5 shard (Q = 8192 x 125 k, k = 5): stall samples of its epilogue
set -u
mkdir -p gpurun_out
O=gpurun_out
GQ=8192 GN=125000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel" --launch-skip 3 --launch-count 2 -f -o $O/r02_gallery_q8192 python tools/gallery_prof.py > $O/r02bn_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r02bn_ncu.log; ls -la $O/r02_gallery_q8192.ncu-rep
