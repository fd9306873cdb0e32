#!/bin/bash
# ncu --set full of K3a alone at the cfg4 rank shape (after the fast-path restructure), source view for the stall samples
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel" --launch-skip 5 --launch-count 2 -f -o $O/r02_cfg4rank_b python tools/head_prof.py > $O/r02az_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02az_ncu.log; ls -la $O/r02_cfg4rank_b.ncu-rep
