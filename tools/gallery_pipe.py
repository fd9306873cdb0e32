#!/usr/bin/env python
"""Throughput of the streaming-regime gallery match (Q = 128 vs 1 M x 512, top-5) with 1, 2 and 3 query batches in
flight on separate CUDA streams (scratch is cached per stream, b200face/_lib.py:workspace).  One call is a chain of
dependent kernels (query prepare -> sample scan -> tau -> main scan -> select); with several batches in flight the
latency-bound head and tail of one call can overlap the main scan of another.  Measurement aid: prints queries/s and
the streamed bytes / time against the measured HBM peak for each depth."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200face
from b200face import _lib

dev = torch.device("cuda:0")
Q, N, D, k = int(os.environ.get("GQ", 128)), int(os.environ.get("GN", 1_000_000)), 512, 5
CALLS = int(os.environ.get("CALLS", 60))
peak = 6548.2
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
g = torch.Generator(device=dev).manual_seed(1)
G = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=dev), dim=1)
prep = b200face.PreparedGallery(G, "l2eps")
batches = [torch.nn.functional.normalize(torch.randn(Q, D, generator=g, device=dev), dim=1) for _ in range(6)]
ref = [b200face.gallery_topk(b, G, k, 1.0, "l2eps", engine=_lib.ENGINE_TCGEN05, prepared=prep) for b in batches]
torch.cuda.synchronize()
byt = N * D * 2 + N * 4 + Q * D * 4 + Q * k * 12
res = {}
for depth in (1, 2, 3, 4):
    streams = [torch.cuda.Stream(dev) for _ in range(depth)]
    outs = [None] * len(batches)

    def run(n):
        for i in range(n):
            s = streams[i % depth]
            with torch.cuda.stream(s):
                outs[i % len(batches)] = b200face.gallery_topk(batches[i % len(batches)], G, k, 1.0, "l2eps",
                                                               engine=_lib.ENGINE_TCGEN05, prepared=prep)
    run(2 * depth + 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s in streams:
        s.wait_stream(torch.cuda.current_stream(dev))
    e0.record()
    for s in streams:
        s.wait_event(e0)
    run(CALLS)
    for s in streams:
        torch.cuda.current_stream(dev).wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / CALLS
    same = all(torch.equal(outs[j][0], ref[j][0]) and torch.equal(outs[j][1], ref[j][1]) for j in range(len(batches))
               if outs[j] is not None)
    res[f"depth{depth}"] = {"us_per_call": round(ms * 1e3, 1), "queries_per_sec": round(Q / (ms * 1e-3), 1),
                            "GBps": round(byt / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(byt / (ms * 1e-3) / 1e9 / peak, 4),
                            "results_identical_to_serial": bool(same)}
    print(f"depth {depth}:", res[f"depth{depth}"], flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
