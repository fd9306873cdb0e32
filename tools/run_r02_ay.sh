#!/bin/bash
# k3b_wh_prefetch: w_hat boxes of the dW epilogue pulled into L2 ahead of the two in flight.  Per-kernel A/B (K3b alone on the
# whole chip) and the cfg3 step (K3b beside K3c), interleaved twice.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_probe.py k3b_wh_prefetch=0 k3b_wh_prefetch=4 k3b_wh_prefetch=8 k3b_wh_prefetch=16 > $O/r02ay_ab.log 2>&1; grep 512x $O/r02ay_ab.log | cut -c1-220
for rnd in 1 2; do for f in 0 4 8 16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg4 --no-train-step --no-gallery --no-cpu-baseline --tune k3b_wh_prefetch=$f > $O/r02ay_bench_f${f}_$rnd.json 2> $O/r02ay_bench.err
  python -c "import json; d=json.load(open('$O/r02ay_bench_f${f}_$rnd.json')); print('wh_prefetch=$f', d['ms_per_step'], d['burst']['ms_per_step'], d['e2e']['value'])"
done; done
