#!/bin/bash
# Hypothesis test (probe build, WRONG gradients by design): does K3a get faster when a warp's G^T slice is 2 KB contiguous
# (tile-blocked layout) instead of 32 row pieces of 64 bytes?  K3a's own time is all that is read from this.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_probe.py pair=2 2>&1 | grep 512x | cut -c1-120
B200FACE_LIB=$PWD/tools/build_probe_gtblocked/libb200face_probe_gtblocked.so timeout 300 python tools/ab_probe.py pair=2 2>&1 | grep 512x | cut -c1-120
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 pair=2 2>&1 | grep 4096x | cut -c1-120
B200FACE_LIB=$PWD/tools/build_probe_gtblocked/libb200face_probe_gtblocked.so timeout 300 python tools/ab_probe.py --shape 4096,125000,512 pair=2 2>&1 | grep 4096x | cut -c1-120
