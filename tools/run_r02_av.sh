#!/bin/bash
# One rank's share of cfg4 (B = 4096 x 125 k classes) on one GPU: per-kernel event pairs, then ncu --set full of its four
# GEMM kernels, then the launch list of the same command.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 pair=2 > $O/r02av_ab.log 2>&1; echo "ab rc=$?"; tail -2 $O/r02av_ab.log | cut -c1-400
HTIME=1 timeout 300 python tools/head_prof.py > $O/r02av_plain.log 2>&1; echo "plain rc=$?"; tail -2 $O/r02av_plain.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel|gemm_kernel" --launch-skip 8 --launch-count 4 -f -o $O/r02_cfg4rank python tools/head_prof.py > $O/r02av_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02av_ncu.log; ls -la $O/r02_cfg4rank.ncu-rep
