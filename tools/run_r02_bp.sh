#!/bin/bash
# final tree: ncu --set full of the four GEMM kernels at one rank's share of cfg4 (B = 4096 x 125 k), for the before / after table
set -u
mkdir -p gpurun_out
O=gpurun_out
HTIME=1 timeout 300 python tools/head_prof.py > $O/r02bp_plain.log 2>&1; tail -1 $O/r02bp_plain.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel|gemm_kernel" --launch-skip 8 --launch-count 4 -f -o $O/r02_cfg4rank_after python tools/head_prof.py > $O/r02bp_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $O/r02bp_ncu.log
