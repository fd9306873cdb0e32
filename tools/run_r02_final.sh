#!/bin/bash
# round-2 final tree: full GPU test suite, default bench line + reference arm, launch list, ncu --set full of one eager step
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02q_pytest.log
timeout 900 python bench.py > $O/r02q_bench.json 2> $O/r02q_bench.err; echo "bench rc=$?"; tail -3 $O/r02q_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02q_ref.json 2> $O/r02q_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02q_bench.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","burst","parity","roofline","kernel_ms","loss","gpu_launches","clocks"):
    print(k, json.dumps(d.get(k))[:500])
g=d.get("gallery",{})
for k,v in g.items():
    if isinstance(v,dict): print("gallery",k,{kk:v[kk] for kk in v if kk in("ms","frac_of_hbm_peak","queries_per_sec","frac","algorithmic_tflops")}, (v.get("pipelined") or {}).get("frac_of_hbm_peak"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > $O/r02q_short.json 2> $O/r02q_short.err && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02q_launches_raw.csv $CMD > $O/r02q_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
python tools/condense_launches.py $O/r02q_launches_raw.csv $O/r02_launches_bench_cfg3.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv $CMD"; echo "condense rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-gallery --no-cpu-baseline --no-train-step --no-cfg4 --eager"
timeout 300 $CMD2 > $O/r02q_ncu_plain.json 2> $O/r02q_ncu_plain.err && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel|gemm_kernel|l2norm_rows" --launch-skip 21 --launch-count 8 -f -o $O/r02q_step $CMD2 > $O/r02q_ncu.log 2>&1
echo "ncu full rc=$?"; tail -3 $O/r02q_ncu.log; ls -la $O/r02q_step.ncu-rep
# gallery: launch list + full capture of the streaming call
GTIME=1 timeout 120 python tools/gallery_prof.py > $O/r02q_gal_plain.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_gallery_q128.csv python tools/gallery_prof.py > $O/r02q_gal_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"xw_kernel|gallery_select|gallery_tau" --launch-skip 6 --launch-count 4 -f -o $O/r02q_gallery python tools/gallery_prof.py > $O/r02q_gal_ncu_full.log 2>&1
echo "gallery ncu rc=$?"; tail -2 $O/r02q_gal_plain.log
# per-tile timelines of the X-stationary kernels (instrumented build, tools/build_timeline.sh) and the TMA ingest microbenchmark
export B200FACE_LIB=$PWD/tools/build_tl/libb200face_tl.so
for v in k2 k3a k3b; do timeout 200 python tools/timeline_probe.py $v > $O/r02q_timeline_$v.txt 2>&1; echo "timeline $v rc=$?"; done
unset B200FACE_LIB
timeout 200 tools/microbench/tma_bw > $O/r02q_tma_bw.txt 2>&1; echo "tma_bw rc=$?"
