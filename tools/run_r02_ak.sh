#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/e2e_probe.py > $O/r02ak_e2e.log 2>&1; echo "e2e rc=$?"; tail -12 $O/r02ak_e2e.log
timeout 1200 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02ak_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ak_pytest.log | cut -c1-300
