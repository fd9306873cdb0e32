#!/usr/bin/env python
"""A/B of tunables on the cfg3 head step: per-kernel event pairs (stage_events), L2 flushed before each stage, variants
interleaved twice; dx_hat / dW of every variant against the first one (max relative difference, norm-wise).
usage: ab_probe.py [--shape B,C,D] name=value[,name=value...] name=value[,...] ...      Development aid."""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face import _lib
from b200face import head as H
lib = b200face.load_library()
dev = torch.device("cuda:0")
args = sys.argv[1:]
B, C, D = 512, 100000, 512
if args and args[0] == "--shape":
    B, C, D = (int(v) for v in args[1].split(","))
    args = args[2:]
variants = [[(kv.split("=")[0], int(kv.split("=")[1])) for kv in a.split(",")] for a in args]
g = torch.Generator(device=dev).manual_seed(1)
w = (torch.randn(C, D, generator=g, device=dev) * 0.006).bfloat16()
x = torch.randn(B, D, generator=g, device=dev).bfloat16()
y = torch.randint(0, C, (B,), generator=g, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ms_c = ctypes.c_float()
cfg = H._head_cfg(0.45, 6.72, 0.05, False, C + 1, _lib.ENGINE_AUTO)


def run(var, iters=8):
    old = [(n, lib.b200f_set_tunable(n.encode(), v)) for n, v in var]
    acc = {k: [] for k in ("k2", "k3a", "k3b", "k3c")}
    try:
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
    finally:
        for n, v in old:
            lib.b200f_set_tunable(n.encode(), v)
    return (float(out2[0]), dxhat.clone(), dw.clone()), {k: round(statistics.mean(v), 1) for k, v in acc.items()}


base = None
for rnd in range(2):
    for var in variants:
        res, t = run(var)
        if base is None: base = res
        rel = lambda a, b: float((a - b).norm() / b.norm())
        print(f"{B}x{C}x{D} {var}: {t} sum {round(sum(t.values()), 1)} | loss {res[0]:.6f} dx rel {rel(res[1], base[1]):.2e} dw rel {rel(res[2], base[2]):.2e} "
              f"finite {bool(torch.isfinite(res[2]).all() and torch.isfinite(res[1]).all())}", flush=True)
print("timeout flag", lib.b200f_umma_timeout_flag(0))
