#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ac_pytest.log | cut -c1-400
timeout 300 python tools/careful_probe.py 2>&1 | tail -5
