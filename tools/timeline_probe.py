#!/usr/bin/env python
"""Per-tile timeline of the X-stationary kernel at the cfg3 shape (instrumented build, tools/build_timeline.sh):
where do the MMA issuer, the epilogue warps and the TMA producer of ONE CTA spend a tile?  SM clock stamps (XW_TL) of the
first cluster, printed per tile in microseconds at the nominal 1.965 GHz.  Development aid.
usage: B200FACE_LIB=tools/build_tl/libb200face_tl.so timeline_probe.py k2|k3a|k3b [name=value ...]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face import _lib
from b200face import head as H
lib = b200face.load_library()
dev = torch.device("cuda:0")
which = sys.argv[1]
B, C, D = 512, 100000, 512
NOFLUSH = False
for kv in sys.argv[2:]:
    n, v = kv.split("=")
    if n == "C": C = int(v)
    elif n == "B": B = int(v)
    elif n == "noflush": NOFLUSH = bool(int(v))
    else: lib.b200f_set_tunable(n.encode(), int(v))
g = torch.Generator(device=dev).manual_seed(1)
w = (torch.randn(C, D, generator=g, device=dev) * 0.006).bfloat16()
x = torch.randn(B, D, generator=g, device=dev).bfloat16()
y = torch.randint(0, C, (B,), generator=g, device=dev)
cfg = H._head_cfg(0.45, 6.72, 0.05, False, C + 1, _lib.ENGINE_AUTO)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
buf = torch.zeros(4 * 4 * 32 * 4, dtype=torch.int64, device=dev)
lib.b200f_xw_timeline.restype = ctypes.c_int
lib.b200f_xw_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]


def step(record):
    if not NOFLUSH: flush.zero_()
    if record and which == "k2": lib.b200f_xw_timeline(buf.data_ptr(), 0)
    out = H._fwd_kernels(x, w, y, cfg, 0, False)
    lse = torch.empty(B, dtype=torch.float32, device=dev); out2 = torch.empty(2, dtype=torch.float32, device=dev)
    _lib.check(lib.b200f_arcface_loss(_lib.ptr(out[4]), B, cfg, _lib.ptr(lse), _lib.ptr(out2), _lib.ptr(out2[1:]), _lib.stream_ptr(dev)), "loss")
    out4 = torch.empty(4, dtype=torch.float32, device=dev)
    _lib.check(lib.b200f_arcface_hook_scale(_lib.ptr(out2[1:]), None, B, 6.72, 0, 1.0, 1, 0, _lib.ptr(out4), _lib.stream_ptr(dev)), "hook")
    if not NOFLUSH: flush.zero_()
    if record and which in ("k3a", "k3b"): lib.b200f_xw_timeline(buf.data_ptr(), 0 if which == "k3a" else 1)
    H._bwd_kernels(out[0], out[1], y, out[2], out[3], lse, out4, cfg, 0)
    torch.cuda.synchronize()
    lib.b200f_xw_timeline(None, -1)


for _ in range(3): step(False)
step(True)
t = buf.cpu().view(4, 4, 32, 4).numpy().astype("int64")
GHZ = 1.965e3   # cycles per microsecond
for cta in (0, 1):
    t0 = min(int(v) for v in t[cta].reshape(-1) if v > 0) if (t[cta] > 0).any() else 0
    us = lambda v: (v - t0) / GHZ if v > 0 else float("nan")
    print(f"--- {which} CTA {cta} (leader = CTA 0 of the pair issues the MMAs) ---")
    print("tile | MMA: wait_acc granted first_full last_mma (dur grant->last) | epi0: wait full done (E) | epiL: wait full done (E) | TMA: first last")
    for i in range(16):                                        # slots 16.. hold tile 3's k-blocks (below)
        m, e0, e1, pr = t[cta, 0, i], t[cta, 1, i], t[cta, 2, i], t[cta, 3, i]
        if not (m.any() or e0.any() or pr.any()): continue
        f = lambda a: " ".join(f"{us(v):7.2f}" for v in a)
        md = (m[3] - m[1]) / GHZ if m[3] else float("nan")
        print(f"{i:3d} | {f(m)} ({md:5.2f}) | {f(e0[:3])} ({(e0[2]-e0[1])/GHZ:5.2f}) | {f(e1[:3])} ({(e1[2]-e1[1])/GHZ:5.2f}) | {f(pr[:2])}")
    print("tile 3, per k-block | TMA: slot wait, slot granted (load issued) | MMA: stage full seen, MMAs + commit issued | issue -> full")
    for kb in range(8):
        pr, m = t[cta, 3, 16 + kb], t[cta, 0, 16 + kb]
        if not (pr.any() or m.any()): continue
        print(f"  kb {kb} | {us(pr[0]):7.2f} {us(pr[1]):7.2f} | {us(m[0]):7.2f} {us(m[1]):7.2f} | {(m[0] - pr[1]) / GHZ if m[0] and pr[1] else float('nan'):5.2f}")
