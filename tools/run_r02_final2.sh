#!/bin/bash
# round-2 final tree (after the K3a epilogue work, the n-fastest dW GEMM and LazyArcLogits): full GPU test suite + smoke, default
# bench line + the driver's --steps 20 form + reference arm, launch list, ncu --set full of one eager cfg3 step, gallery lines
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02bf_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02bf_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02bf_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02bf_smoke.log
timeout 900 python bench.py > $O/r02bf_bench.json 2> $O/r02bf_bench.err; echo "bench rc=$?"; tail -2 $O/r02bf_bench.err | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02bf_bench20.json 2> $O/r02bf_bench20.err; echo "bench20 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02bf_ref.json 2> $O/r02bf_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02bf_bench.json", "gpurun_out/r02bf_bench20.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f)
    for k in ("value","ms_per_step","e2e","burst","parity","roofline","kernel_ms","loss","gpu_launches","clocks"):
        print(" ", k, json.dumps(d.get(k))[:400])
    g=d.get("gallery",{})
    for k,v in g.items():
        if isinstance(v,dict): print("  gallery",k,{kk:v[kk] for kk in v if kk in("ms","frac_of_hbm_peak","queries_per_sec","frac","algorithmic_tflops")}, (v.get("pipelined") or {}).get("frac_of_hbm_peak"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > $O/r02bf_short.json 2> $O/r02bf_short.err && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02bf_launches_raw.csv $CMD > $O/r02bf_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
python tools/condense_launches.py $O/r02bf_launches_raw.csv $O/r02_launches_bench_cfg3.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv $CMD"; echo "condense rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-gallery --no-cpu-baseline --no-train-step --no-cfg4 --eager"
timeout 300 $CMD2 > $O/r02bf_ncu_plain.json 2> $O/r02bf_ncu_plain.err && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel|gemm_kernel|l2norm_rows" --launch-skip 21 --launch-count 8 -f -o $O/r02bf_step $CMD2 > $O/r02bf_ncu.log 2>&1
echo "ncu full rc=$?"; tail -2 $O/r02bf_ncu.log; ls -la $O/r02bf_step.ncu-rep
