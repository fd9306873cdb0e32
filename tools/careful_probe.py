#!/usr/bin/env python
"""How much do the 'careful' epilogue slices (those that hold a target column) cost K2 / K3a?  The same cfg3 step with the
class range moved away from every label (class_offset: no slice ever holds a target) against the normal one; per-kernel
event pairs (stage_events), L2 flushed before each stage.  Development aid."""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face import _lib
from b200face import head as H
lib = b200face.load_library()
dev = torch.device("cuda:0")
B, C, D = 512, 100000, 512
g = torch.Generator(device=dev).manual_seed(1)
w = (torch.randn(C, D, generator=g, device=dev) * 0.006).bfloat16()
x = torch.randn(B, D, generator=g, device=dev).bfloat16()
y = torch.randint(0, C, (B,), generator=g, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ms_c = ctypes.c_float()
for off, name, patch in ((0, "labels in range, whole slices element-wise (round 1)", 0), (0, "labels in range, target elements patched", 1),
                         (10_000_000, "labels out of range (no target in any slice)", 1), (0, "labels in range, whole slices element-wise (round 1)", 0),
                         (0, "labels in range, target elements patched", 1)):
    lib.b200f_set_tunable(b"target_patch", patch)
    cfg = H._head_cfg(0.45, 6.72, 0.05, False, C + off + 1, _lib.ENGINE_AUTO)
    acc = {k: [] for k in ("k2", "k3a", "k3b", "k3c")}
    for it in range(8):
        flush.zero_()
        lib.b200f_set_tunable(b"stage_events", 1)
        out = H._fwd_kernels(x, w, y, cfg, off, False)
        lse = torch.empty(B, dtype=torch.float32, device=dev); out2 = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(lib.b200f_arcface_loss(_lib.ptr(out[4]), B, cfg, _lib.ptr(lse), _lib.ptr(out2), _lib.ptr(out2[1:]), _lib.stream_ptr(dev)), "loss")
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        _lib.check(lib.b200f_arcface_hook_scale(_lib.ptr(out2[1:]), None, B, 6.72, 0, 1.0, 1, 0, _lib.ptr(out4), _lib.stream_ptr(dev)), "hook")
        flush.zero_()
        H._bwd_kernels(out[0], out[1], y, out[2], out[3], lse, out4, cfg, off)
        torch.cuda.synchronize()
        lib.b200f_set_tunable(b"stage_events", 0)
        for k in acc:
            _lib.check(lib.b200f_stage_ms(k.encode(), ctypes.byref(ms_c)), "stage_ms")
            if it >= 2: acc[k].append(float(ms_c.value) * 1e3)
    print(name, {k: round(statistics.mean(v), 1) for k, v in acc.items()})
