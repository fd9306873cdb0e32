#!/usr/bin/env python
"""Time the X-stationary kernel's TMA / tcgen05 / TMEM pipeline with a do-nothing epilogue (b200f_umma_xw_probe)
at the cfg3 shape, single CTAs vs CTA pairs.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face import _lib
lib = b200face.load_library()
dev = torch.device("cuda:0")
for (B, C, D) in ((512, 100000, 512), (4096, 125000, 512), (128, 1000000, 512)):
    g = torch.Generator(device=dev).manual_seed(1)
    x = (torch.randn(B, D, generator=g, device=dev) / 16).half()
    w = (torch.randn(C, D, generator=g, device=dev) / 16).half()
    ref = (x.float() @ w.float().t()).sum(1) if B * C <= 512 * 100000 else None
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for pair in (1, 2):
        out = torch.zeros(B, device=dev)
        ts = []
        for i in range(6):
            out.zero_(); flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.b200f_umma_xw_probe(_lib.ptr(x), _lib.ptr(w), _lib.ptr(out), B, C, D, pair, _lib.stream_ptr(dev)), "probe")
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        err = float((out - ref).abs().max() / ref.abs().max()) if ref is not None else None
        t = min(ts[1:])
        print(f"B={B} C={C} D={D} pair={pair}: {t*1e3:.1f} us  {2.0*B*C*D/t/1e9:.0f} TFLOP/s  err={err}  timeout={lib.b200f_umma_timeout_flag(1)}", flush=True)
