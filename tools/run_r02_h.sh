#!/bin/bash
set -u
mkdir -p gpurun_out
GTIME=1 GTUNE="xw_prefetch=0;xw_prefetch=1;xw_prefetch=2;xw_prefetch=4" timeout 300 python tools/gallery_prof.py 2>&1 | tail -6
