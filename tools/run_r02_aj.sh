#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_probe.py --shape 4096,32000,512 epi_groups=1 epi_groups=2 epi_groups=4 > $O/r02aj_ab.log 2>&1; echo "ab rc=$?"; tail -7 $O/r02aj_ab.log | cut -c1-300
timeout 300 python tools/ab_probe.py --shape 1024,60000,512 epi_groups=1 epi_groups=2 epi_groups=4 > $O/r02aj_ab2.log 2>&1; echo "ab2 rc=$?"; tail -7 $O/r02aj_ab2.log | cut -c1-300
timeout 300 python tools/e2e_probe.py > $O/r02aj_e2e.log 2>&1; echo "e2e rc=$?"; tail -12 $O/r02aj_e2e.log
