#!/bin/bash
# GPU call: full GPU suite, row-kernel timings, bench line (with the pipelined gallery figure).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/pytest_gpu.log
python tools/rowops_prof.py $O/rowops_live.json > $O/rowops_live.log 2>&1; echo "rowops_prof rc=$?"
tail -2 $O/rowops_live.log
timeout 400 python bench.py > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('$O/bench_final.json'));print(d['value'],json.dumps(d['gallery']['stream_q128_n1m']))"
