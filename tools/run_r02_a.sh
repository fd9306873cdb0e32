#!/bin/bash
# Round-2 call A: GPU test suite (incl. the new full-size parity tests) + A/B of the K2 / K3a epilogue geometry.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/r02a_pytest.log
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | grep -E "cfg3|cfg4|epi_groups|passed|failed" | tail -12
B="--steps 300 --warmup 5 --no-cpu-baseline --no-gallery --no-train-step --no-cfg4"
timeout 300 python bench.py $B --tune epi_groups=1 > $O/r02a_bench_eg1.json 2> $O/r02a_bench_eg1.err; echo "bench eg1 rc=$?"
timeout 300 python bench.py $B --tune epi_groups=2 > $O/r02a_bench_eg2.json 2> $O/r02a_bench_eg2.err; echo "bench eg2 rc=$?"
python - <<'PY'
import json
for n in ("eg1","eg2"):
    try:
        d=json.loads(open(f"gpurun_out/r02a_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", d["ms_per_step"], "burst", d["burst"]["ms_per_step"], "kernel_ms", d["kernel_ms"], "loss", d["loss"])
    except Exception as e:
        print(n, "no line", e); print(open(f"gpurun_out/r02a_bench_{n}.err").read()[-1500:])
PY
