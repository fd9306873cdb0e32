#!/bin/bash
# K3a with deferred G^T stores (+ FMNMX3): head tests, per-kernel numbers at cfg3 and the cfg4 rank shape, short bench
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02ba_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ba_pytest.log | cut -c1-300
timeout 300 python tools/ab_probe.py pair=2 > $O/r02ba_ab_cfg3.log 2>&1; grep 512x $O/r02ba_ab_cfg3.log | cut -c1-200
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 pair=2 > $O/r02ba_ab_cfg4.log 2>&1; grep 4096x $O/r02ba_ab_cfg4.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg4 --no-train-step --no-gallery --no-cpu-baseline > $O/r02ba_bench.json 2> $O/r02ba_bench.err
python -c "import json; d=json.load(open('$O/r02ba_bench.json')); print(d['ms_per_step'], d['burst']['ms_per_step'], d['e2e']['value'], d['kernel_ms'])"
