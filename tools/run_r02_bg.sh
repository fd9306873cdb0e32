#!/bin/bash
# K3a with G^T through staging + TMA tensor stores (tunable k3a_tma_store): bit-identity test, per-kernel A/B at both shapes, step A/B
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_head.py -m gpu -q -x -k "work_order or few_classes or bf16_vs_oracle" > $O/r02bg_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02bg_pytest.log | cut -c1-300
timeout 300 python tools/ab_probe.py k3a_tma_store=0 k3a_tma_store=1 > $O/r02bg_ab_cfg3.log 2>&1; grep 512x $O/r02bg_ab_cfg3.log | cut -c1-220
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 k3a_tma_store=0 k3a_tma_store=1 > $O/r02bg_ab_cfg4.log 2>&1; grep 4096x $O/r02bg_ab_cfg4.log | cut -c1-220
for rnd in 1 2; do for f in 0 1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg4 --no-train-step --no-gallery --no-cpu-baseline --tune k3a_tma_store=$f > $O/r02bg_bench_f${f}_$rnd.json 2> $O/r02bg_bench.err
  python -c "import json; d=json.load(open('$O/r02bg_bench_f${f}_$rnd.json')); print('k3a_tma_store=$f', d['ms_per_step'], d['burst']['ms_per_step'], d['e2e']['value'], d['parity']['dw_rel'])"
done; done
