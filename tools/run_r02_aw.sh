#!/bin/bash
# After the K3a fast-path restructure + pinned tcgen05.ld prefetches: full GPU suite, per-kernel A/B numbers at cfg3 and the
# cfg4 rank shape, gallery per-call times (Q = 128 / 256 / 8192 shard), short bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02aw_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02aw_pytest.log | cut -c1-300
timeout 300 python tools/ab_probe.py pair=2 > $O/r02aw_ab_cfg3.log 2>&1; tail -1 $O/r02aw_ab_cfg3.log | cut -c1-300
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 pair=2 > $O/r02aw_ab_cfg4.log 2>&1; tail -1 $O/r02aw_ab_cfg4.log | cut -c1-300
GTIME=1 timeout 300 python tools/gallery_prof.py > $O/r02aw_gal128.log 2>&1; tail -2 $O/r02aw_gal128.log | cut -c1-300
GQ=256 GTIME=1 timeout 300 python tools/gallery_prof.py > $O/r02aw_gal256.log 2>&1; tail -2 $O/r02aw_gal256.log | cut -c1-300
GQ=8192 GN=125000 GTIME=1 timeout 300 python tools/gallery_prof.py > $O/r02aw_gal8192.log 2>&1; tail -2 $O/r02aw_gal8192.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cfg4 --no-train-step > $O/r02aw_bench.json 2> $O/r02aw_bench.err; echo "bench rc=$?"; cut -c1-400 $O/r02aw_bench.json
