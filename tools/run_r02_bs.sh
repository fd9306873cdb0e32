#!/bin/bash
# L2 prefetch distance of the streamed operand (tiles ahead of the TMA ring), re-measured now that K3a's tile is bound by the
# latency of its ring loads (timeline: 1.6-2.6 us from issue to full, 5 stages)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/ab_probe.py xw_prefetch=0 xw_prefetch=1 xw_prefetch=2 xw_prefetch=3 > $O/r02bs_ab_cfg3.log 2>&1; grep 512x $O/r02bs_ab_cfg3.log | cut -c1-130
timeout 300 python tools/ab_probe.py --shape 4096,125000,512 xw_prefetch=0 xw_prefetch=1 xw_prefetch=2 > $O/r02bs_ab_cfg4.log 2>&1; grep 4096x $O/r02bs_ab_cfg4.log | cut -c1-130
