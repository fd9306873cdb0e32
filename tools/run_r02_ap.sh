#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "tma_stores or variants" > $O/r02ap_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ap_pytest.log | cut -c1-300
timeout 300 python tools/ab_probe.py k3b_groups=1 k3b_groups=2 > $O/r02ap_ab.log 2>&1; tail -5 $O/r02ap_ab.log | cut -c1-200
