#!/bin/bash
# Round-1 closing GPU call: row-kernel timings + their ncu --set full capture, the gallery pipelining measurement,
# the bench line and the GPU test suite of the final tree.  Every step writes under gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
python tools/rowops_prof.py $O/rowops_live.json > $O/rowops_live.log 2>&1; echo "rowops_prof rc=$?"
REPS=1 timeout 300 ncu --set full --import-source on --clock-control none \
  -k regex:"l2norm_rows|l2norm_bwd|adamw_rows|gallery_prepare" --launch-count 16 -f -o $O/r01_rowops \
  python tools/rowops_prof.py > $O/rowops_ncu.log 2>&1; echo "ncu rowops rc=$?"
python tools/ncu_summary.py $O/r01_rowops.ncu-rep $O/r01_rowops_full > $O/rowops_summary.log 2>&1; echo "summary rc=$?"
timeout 120 python tools/gallery_pipe.py $O/gallery_pipe.json > $O/gallery_pipe.log 2>&1; echo "gallery_pipe rc=$?"
GQ=256 timeout 120 python tools/gallery_pipe.py $O/gallery_pipe_q256.json > $O/gallery_pipe_q256.log 2>&1; echo "gallery_pipe q256 rc=$?"
timeout 400 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
cat $O/rowops_live.log | tail -8
cat $O/gallery_pipe.log | tail -4
tail -5 $O/gallery_pipe_q256.log
