#!/bin/bash
# K3a / K3b walk directions (what the previous kernel touched last is read first): cfg3 step A/B, interleaved twice
set -u
mkdir -p gpurun_out
O=gpurun_out
for rnd in 1 2; do for v in "0 1" "1 0" "1 1" "0 0"; do
  set -- $v
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg4 --no-train-step --no-gallery --no-cpu-baseline --tune k3a_reverse=$1 --tune k3b_reverse=$2 > $O/r02bc_bench_$1$2_$rnd.json 2> $O/r02bc_bench.err
  python -c "import json; d=json.load(open('$O/r02bc_bench_$1$2_$rnd.json')); print('k3a_rev=$1 k3b_rev=$2', d['ms_per_step'], d['burst']['ms_per_step'], d['e2e']['value'], d['parity']['dx_rel'], d['parity']['dw_rel'])"
done; done
