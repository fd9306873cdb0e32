#!/bin/bash
# ncu --set full of K2 with K1(W) inside (one launch), source-level counters
set -u
mkdir -p gpurun_out
O=gpurun_out
CMD2="python bench.py --steps 2 --warmup 3 --no-gallery --no-cpu-baseline --no-train-step --no-cfg4 --eager --tune k2_prep=1"
timeout 300 $CMD2 > $O/r02l_plain.json 2> $O/r02l_plain.err && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel" --launch-skip 6 --launch-count 3 -f -o $O/r02l_k2prep $CMD2 > $O/r02l_ncu.log 2>&1
echo "ncu full rc=$?"; tail -3 $O/r02l_ncu.log; ls -la $O/r02l_k2prep.ncu-rep
