#!/usr/bin/env python
"""Where does the time of K2 / K3a / K3b / K3c go at batch 512?  Times the do-nothing-epilogue pipeline and the real kernels
over a range of class counts, cold (L2 flushed) and hot (same launch repeated; the streamed operand fits L2 for small
C): the slope over C is the per-tile cost, the intercept the fixed cost, hot vs cold the share of DRAM latency.
Development aid (prints a table)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face import _lib
from b200face import head as H
lib = b200face.load_library()
dev = torch.device("cuda:0")
B, D = int(os.environ.get("PB", 512)), 512
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ms_c = ctypes.c_float()
for kv in os.environ.get("HTUNE", "").split(","):
    if "=" in kv:
        k, v = kv.split("="); print("tunable", k, v, "was", lib.b200f_set_tunable(k.encode(), int(v)))

def timed(fn, cold, reps=6):
    ts = []
    for _ in range(reps):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts[1:])

print(f"B={B} D={D}; times in us (min of 5)")
print(f"{'C':>8} {'tiles/cl':>8} | {'null cold':>9} {'null hot':>9} | " + " | ".join(f"{k+' cold':>9} {k+' hot':>9}" for k in ("k2", "k3a", "k3b", "k3c")))
for C in [int(c) for c in os.environ.get("PCS", "9472,18944,37888,75776,100000,151552").split(",")]:
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, D, generator=g, device=dev).bfloat16()
    w = (torch.randn(C, D, generator=g, device=dev) * 0.0063).bfloat16()
    y = torch.randint(0, C, (B,), generator=g, device=dev)
    xo, inv_nx = H._k1(x, True); wo, inv_nw = H._k1(w, True)
    out = torch.zeros(B, device=dev)
    null = lambda: _lib.check(lib.b200f_umma_xw_probe(_lib.ptr(xo), _lib.ptr(wo), _lib.ptr(out), B, C, D, 2, _lib.stream_ptr(dev)), "probe")
    res = {"null": (timed(null, True), timed(null, False))}
    cfg = H._head_cfg(0.45, 6.72, 0.05, False, C, _lib.ENGINE_AUTO)
    cache = {"static": (wo, inv_nw)}
    lib.b200f_set_tunable(b"stage_events", 1)
    for cold in (True, False):
        acc = {k: [] for k in ("k2", "k3a", "k3b", "k3c")}
        for _ in range(6):
            if cold: flush.zero_()
            o = H._fwd_kernels(x, w, y, cfg, 0, False, cache, None, fused_hook=_lib.HookCfg(0, 1.0, 1, 0))
            lse, out4 = o[10], o[13]
            if cold: flush.zero_()
            H._bwd_kernels(o[0], o[1], y, o[2], o[3], lse, out4, cfg, 0)
            torch.cuda.synchronize()
            for k in acc:
                _lib.check(lib.b200f_stage_ms(k.encode(), ctypes.byref(ms_c)), "stage_ms"); acc[k].append(ms_c.value * 1e3)
        for k in acc:
            res.setdefault(k, [None, None])[0 if cold else 1] = min(acc[k][1:])
    lib.b200f_set_tunable(b"stage_events", 0)
    tiles = (C + 255) // 256 * ((B + 255) // 256) / 74
    print(f"{C:>8} {tiles:>8.1f} | {res['null'][0]:>9.1f} {res['null'][1]:>9.1f} | " +
          " | ".join(f"{res[k][0]:>9.1f} {res[k][1]:>9.1f}" for k in ("k2", "k3a", "k3b", "k3c")), flush=True)
    del x, w, xo, wo
