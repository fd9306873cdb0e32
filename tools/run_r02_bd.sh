#!/bin/bash
# stream-K of the dx GEMM: head + fullsize tests, per-kernel numbers at the cfg4 rank shape (on / off), cfg3 unchanged
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_umma.py -m gpu -q -x > $O/r02bd_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02bd_pytest.log | cut -c1-300
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 stream_k=0 stream_k=1 > $O/r02bd_ab_cfg4.log 2>&1; grep 4096x $O/r02bd_ab_cfg4.log | cut -c1-220
timeout 300 python tools/ab_probe.py stream_k=0 stream_k=1 > $O/r02bd_ab_cfg3.log 2>&1; grep 512x $O/r02bd_ab_cfg3.log | cut -c1-220
HTIME=1 timeout 300 python tools/head_prof.py > $O/r02bd_plain.log 2>&1; tail -1 $O/r02bd_plain.log
HTIME=1 HTUNE=stream_k=0 timeout 300 python tools/head_prof.py > $O/r02bd_plain0.log 2>&1; tail -1 $O/r02bd_plain0.log
