#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 tools/microbench/tma_bw > gpurun_out/r02ai_tma_bw.log 2>&1; echo "rc=$?"
cat gpurun_out/r02ai_tma_bw.log
