#!/bin/bash
# k3c_follow (dx GEMM walks the class rows in the dW kernel's order) and dw_n_fastest (streamed dW GEMM): head tests, then
# A/B of the cfg3 step (bench --steps 20, interleaved twice) and of the cfg4 rank shape per kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/r02ax_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02ax_pytest.log | cut -c1-300
for rnd in 1 2; do for f in 0 1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg4 --no-train-step --no-gallery --no-cpu-baseline --tune k3c_follow=$f > $O/r02ax_bench_f${f}_$rnd.json 2> $O/r02ax_bench.err
  python -c "import json; d=json.load(open('$O/r02ax_bench_f${f}_$rnd.json')); print('follow=$f', d['ms_per_step'], d['burst']['ms_per_step'], d['e2e']['value'])"
done; done
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 dw_n_fastest=0 dw_n_fastest=1 > $O/r02ax_ab_cfg4.log 2>&1; grep 4096x $O/r02ax_ab_cfg4.log | cut -c1-200
