#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/k3b_store_probe.py > $O/r02ad_k3b_store.log 2>&1; echo "probe rc=$?"; tail -12 $O/r02ad_k3b_store.log | cut -c1-600
