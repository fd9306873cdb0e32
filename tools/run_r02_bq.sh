#!/bin/bash
# K3b beside K3c: share of the CTA pairs for the dx side, re-swept on the final tree (B200F_BWD_SPLIT), interleaved twice
set -u
mkdir -p gpurun_out
O=gpurun_out
for rnd in 1 2; do for s in 18 22 24 26 30; do
  B200F_BWD_SPLIT=$s timeout 300 python bench.py --steps 20 --warmup 5 --no-cfg4 --no-train-step --no-gallery --no-cpu-baseline > $O/r02bq_bench_${s}_$rnd.json 2> $O/r02bq_bench.err
  python -c "import json; d=json.load(open('$O/r02bq_bench_${s}_$rnd.json')); print('split=$s', d['ms_per_step'], d['burst']['ms_per_step'], d['e2e']['value'])"
done; done
