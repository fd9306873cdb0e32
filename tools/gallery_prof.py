#!/usr/bin/env python
"""One streaming-regime gallery call (Q=128 vs 1M x 512) for ncu launch lists.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face import _lib
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
Q, N, D, k = int(os.environ.get("GQ", 128)), int(os.environ.get("GN", 1000000)), 512, 5
G = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=dev), dim=1)
Qm = torch.nn.functional.normalize(torch.randn(Q, D, generator=g, device=dev), dim=1)
prep = b200face.PreparedGallery(G, "l2eps")
for _ in range(3):
    out = b200face.gallery_topk(Qm, G, k, 1.0, "l2eps", engine=_lib.ENGINE_TCGEN05, prepared=prep)
torch.cuda.synchronize()
print("ok", out[0][0].tolist())
if os.environ.get("GTIME"):                                  # per-call time, CUDA events, for each listed tunable setting
    lib = b200face.load_library()
    for setting in os.environ.get("GTUNE", "").split(";"):
        for kv in setting.split(","):
            if "=" in kv:
                lib.b200f_set_tunable(kv.split("=")[0].encode(), int(kv.split("=")[1]))
        ts = []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b200face.gallery_topk(Qm, G, k, 1.0, "l2eps", engine=_lib.ENGINE_TCGEN05, prepared=prep)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"Q={Q} N={N} [{setting}]: median {ts[len(ts)//2]*1e3:.1f} us, min {ts[0]*1e3:.1f} us")
