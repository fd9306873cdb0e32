#!/usr/bin/env python
"""Where do the ~25 us per step between `value` (graph replay, inputs resident) and `e2e` (ArcMarginProduct.graphed_step
called with pinned host tensors + the loss read back every step) go?  Same captured step, 300 steps per variant, variants
interleaved twice.  Development aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
dev = torch.device("cuda:0")
B, C, D = 512, 100000, 512
g = torch.Generator().manual_seed(1)
head = b200face.ArcMarginProduct(D, C).to(dev)
head.update_epoch(12); head.train()
head.compute_dtype = torch.bfloat16
head.cache_weight_prep = False
step = head.graphed_step(B, 0.05, torch.bfloat16)
xh = torch.randn(B, D, generator=g).bfloat16().pin_memory()
yh = torch.randint(0, C, (B,), generator=g).pin_memory()
xd, yd = xh.to(dev), yh.to(dev)
step(xd, yd)
gs = head.__dict__["_graphed"]["step"]          # the GraphedHeadStep behind the public closure
loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
loss_ev = [torch.cuda.Event() for _ in range(2)]
N = 300


def loop(call, readback):
    for i in range(N):
        l = call()
        if readback:
            loss_host[i & 1].copy_(l.detach(), non_blocking=True)
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                float(loss_host[(i - 1) & 1])
    if readback:
        loss_ev[(N - 1) & 1].synchronize()


variants = {
    "replay only (value)": (lambda: gs.replay(), False),
    "replay + loss readback": (lambda: gs.replay(), True),
    "device tensors through step(x, y)": (lambda: step(xd, yd), False),
    "host tensors through step(xh, yh)": (lambda: step(xh, yh), False),
    "host tensors + loss readback (e2e)": (lambda: step(xh, yh), True),
}
for rnd in range(2):
    for name, (call, rb) in variants.items():
        loop(call, rb); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(); loop(call, rb); e1.record(); torch.cuda.synchronize()
        print(f"{name:42s} {e0.elapsed_time(e1) / N * 1e3:7.1f} us/step (host wall {1e6 * (time.time() - t0) / N:7.1f} us/step)", flush=True)
step.close()
