#!/usr/bin/env python
"""K3b with dW through shared-memory staging + TMA tensor stores (tunable k3b_tma_store) against the row-per-lane 32-byte
stores: dW compared bit for bit (same arithmetic, another way out) on cfg3 and on ragged class counts, per-kernel event
pairs (stage_events) with the L2 flushed before each stage.  Development aid."""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face import _lib
from b200face import head as H
lib = b200face.load_library()
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ms_c = ctypes.c_float()


def run(B, C, D, mode, iters, seed=1):
    lib.b200f_set_tunable(b"k3b_tma_store", mode)
    g = torch.Generator(device=dev).manual_seed(seed)
    w = (torch.randn(C, D, generator=g, device=dev) * 0.006).bfloat16()
    x = torch.randn(B, D, generator=g, device=dev).bfloat16()
    y = torch.randint(0, C, (B,), generator=g, device=dev)
    cfg = H._head_cfg(0.45, 6.72, 0.05, False, C + 1, _lib.ENGINE_AUTO)
    acc = {k: [] for k in ("k2", "k3a", "k3b", "k3c")}
    dw = None
    for it in range(iters):
        flush.zero_()
        lib.b200f_set_tunable(b"stage_events", 1)
        out = H._fwd_kernels(x, w, y, cfg, 0, False)
        lse = torch.empty(B, dtype=torch.float32, device=dev); out2 = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(lib.b200f_arcface_loss(_lib.ptr(out[4]), B, cfg, _lib.ptr(lse), _lib.ptr(out2), _lib.ptr(out2[1:]), _lib.stream_ptr(dev)), "loss")
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        _lib.check(lib.b200f_arcface_hook_scale(_lib.ptr(out2[1:]), None, B, 6.72, 0, 1.0, 1, 0, _lib.ptr(out4), _lib.stream_ptr(dev)), "hook")
        flush.zero_()
        dxhat, dw = H._bwd_kernels(out[0], out[1], y, out[2], out[3], lse, out4, cfg, 0)
        torch.cuda.synchronize()
        lib.b200f_set_tunable(b"stage_events", 0)
        for k in acc:
            _lib.check(lib.b200f_stage_ms(k.encode(), ctypes.byref(ms_c)), "stage_ms")
            if it >= 2: acc[k].append(float(ms_c.value) * 1e3)
    return dw, {k: (round(statistics.mean(v), 1) if v else None) for k, v in acc.items()}


for (B, C, D) in ((512, 100000, 512), (512, 99999, 512), (300, 4097, 512), (64, 1000, 256), (512, 33, 512), (640, 24000, 512)):
    it = 8 if C >= 99999 else 3
    a, ta = run(B, C, D, 0, it)
    b, tb = run(B, C, D, 1, it)
    same = torch.equal(a, b)
    nbad = int((a != b).sum()) if not same else 0
    print(f"B={B} C={C} D={D}: identical={same} differing={nbad} finite={bool(torch.isfinite(b).all())} | ST.G {ta} | TMA {tb}", flush=True)
    if not same:
        idx = (a != b).nonzero()[:5].tolist()
        print("   first differences at", idx, [(float(a[i, j]), float(b[i, j])) for i, j in idx])
a, ta = run(512, 100000, 512, 0, 8)
b, tb = run(512, 100000, 512, 1, 8)
print("repeat cfg3: ST.G", ta, "| TMA", tb)
print("timeout flag", lib.b200f_umma_timeout_flag(0))
