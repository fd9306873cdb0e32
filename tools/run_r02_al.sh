#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_probe.py target_patch=1 target_patch=2 > $O/r02al_ab.log 2>&1; echo "ab rc=$?"; tail -5 $O/r02al_ab.log | cut -c1-300
timeout 300 python tools/ab_probe.py --shape 512,2000,512 target_patch=1 target_patch=2 > $O/r02al_ab2.log 2>&1; echo "ab2 rc=$?"; tail -5 $O/r02al_ab2.log | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02al_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02al_pytest.log | cut -c1-300
