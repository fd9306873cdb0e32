#!/bin/bash
# ncu launch list of the bench command (durations only), condensed into gpurun_out/r01_launches_bench_cfg3.csv
set -u
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > $O/bench_short.json 2> $O/bench_short.err; echo "bench short rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_raw.csv $CMD > $O/ncu_launch.log 2>&1; echo "ncu rc=$?"
python tools/condense_launches.py $O/launches_raw.csv $O/r01_launches_bench_cfg3.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv $CMD"; echo "condense rc=$?"
wc -l $O/r01_launches_bench_cfg3.csv; grep -c "xw_kernel" $O/r01_launches_bench_cfg3.csv
