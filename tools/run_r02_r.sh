#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > $O/r02r_pytest.log 2>&1; echo "multirank pytest rc=$?"; tail -5 $O/r02r_pytest.log | cut -c1-600
