#!/bin/bash
# ncu --set full capture of the four GEMM kernels of one eager cfg3 step (after the same command exited 0 without ncu)
set -u
mkdir -p gpurun_out
O=gpurun_out
TUNE="${TUNE:---tune epi_groups=2 --tune xw_prefetch=0}"
CMD="python bench.py --steps 2 --warmup 3 --no-gallery --no-cpu-baseline --no-train-step --no-cfg4 --eager $TUNE"
timeout 300 $CMD > $O/r02_ncu_plain.json 2> $O/r02_ncu_plain.err && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel|gemm_kernel" --launch-skip 16 --launch-count 4 -f -o $O/r02_step $CMD > $O/r02_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $O/r02_ncu.log; ls -la $O/r02_step.ncu-rep
