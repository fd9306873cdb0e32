#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/ab_probe.py epi_groups=2 epi_groups=4 epi_groups=1 > $O/r02ae_ab.log 2>&1; echo "ab rc=$?"; tail -8 $O/r02ae_ab.log | cut -c1-400
timeout 300 python tools/ab_probe.py --shape 4096,24000,512 epi_groups=2 epi_groups=4 > $O/r02ae_ab2.log 2>&1; echo "ab2 rc=$?"; tail -5 $O/r02ae_ab2.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "variants" > $O/r02ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ae_pytest.log | cut -c1-400
