#!/bin/bash
# G^T in 2 KB blocks at batch > 512 (tunable gt_blocked): head + fullsize tests, per-kernel A/B at the cfg4 rank shape
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/r02bk_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02bk_pytest.log | cut -c1-300
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 gt_blocked=0 gt_blocked=1 > $O/r02bk_ab_cfg4.log 2>&1; grep 4096x $O/r02bk_ab_cfg4.log | cut -c1-220
HTIME=1 timeout 300 python tools/head_prof.py > $O/r02bk_plain.log 2>&1; tail -1 $O/r02bk_plain.log
HTIME=1 HTUNE=gt_blocked=0 timeout 300 python tools/head_prof.py > $O/r02bk_plain0.log 2>&1; tail -1 $O/r02bk_plain0.log
