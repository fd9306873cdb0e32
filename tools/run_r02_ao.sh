#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/k3b_store_probe.py > $O/r02ao_k3b.log 2>&1; echo "probe rc=$?"; tail -9 $O/r02ao_k3b.log | cut -c1-400
timeout 1500 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02ao_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ao_pytest.log | cut -c1-300
