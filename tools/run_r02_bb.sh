#!/bin/bash
# lazy logits (forward() -> LazyArcLogits) + FMNMX3 in K3a: head / tail / fullsize tests, per-kernel numbers
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py tests/test_gpu_tail.py -m gpu -q > $O/r02bb_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02bb_pytest.log | cut -c1-300
timeout 300 python tools/ab_probe.py pair=2 > $O/r02bb_ab_cfg3.log 2>&1; grep 512x $O/r02bb_ab_cfg3.log | cut -c1-200
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 pair=2 > $O/r02bb_ab_cfg4.log 2>&1; grep 4096x $O/r02bb_ab_cfg4.log | cut -c1-200
