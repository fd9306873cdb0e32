#!/usr/bin/env python
"""A few eager fused head steps at an arbitrary shape (env HB, HC) for ncu launch lists -- e.g. one rank's share of
cfg4 (HB=4096 HC=125000).  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
from b200face.head import arcface_loss
B, C, D = int(os.environ.get("HB", 4096)), int(os.environ.get("HC", 125000)), 512
for kv in os.environ.get("HTUNE", "").split(","):            # e.g. HTUNE=k3a_ablate=2,pair=1
    if "=" in kv:
        b200face.load_library().b200f_set_tunable(kv.split("=")[0].encode(), int(kv.split("=")[1]))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
w = (torch.randn(C, D, generator=g, device=dev) * 0.006).bfloat16()
wm = w.float().requires_grad_(True)
x = torch.randn(B, D, generator=g, device=dev).bfloat16()
y = torch.randint(0, C, (B,), generator=g, device=dev)
for i in range(3):
    wm.grad = None
    xi = x.clone().requires_grad_(True)
    loss = arcface_loss(xi, wm, y, compute_weight=w, m_eff=0.45, s_eff=6.72, label_smoothing=0.05)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
if os.environ.get("HTIME"):                                  # eager step time, CUDA events (not for use under ncu)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        wm.grad = None
        xi = x.clone().requires_grad_(True)
        loss = arcface_loss(xi, wm, y, compute_weight=w, m_eff=0.45, s_eff=6.72, label_smoothing=0.05)
        loss.backward()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B={B} C={C}: {ms * 1e3:.1f} us per eager step, {3 * 2.0 * B * C * D / ms / 1e9:.0f} TFLOP/s")
