#!/usr/bin/env python
"""Development aid: compare the class-major K3b (XwDwT) against the feature-major one (XwDw) on the same step and
print WHERE they differ (class rows mod tile, feature columns)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import b200face
lib = b200face.load_library()
dev = torch.device("cuda:0")
for (B, C, D) in ((256, 4096, 512), (512, 10240, 512), (512, 100000, 512)):
    g = torch.Generator().manual_seed(B + C)
    w = (torch.randn(C, D, generator=g) * (2.0 / (C + D)) ** 0.5 * 2 ** 0.5).bfloat16().float()
    x = torch.randn(B, D, generator=g).bfloat16()
    y = torch.randint(0, C, (B,), generator=g)
    def run():
        head = b200face.ArcMarginProduct(D, C).to(dev)
        with torch.no_grad():
            head.weight.copy_(w)
        head.train(); head.update_epoch(12)
        xg = x.to(dev).requires_grad_(True)
        loss = head.forward_loss(xg, y.to(dev), 0.05)
        loss.backward(); torch.cuda.synchronize()
        return head.weight.grad.clone()
    lib.b200f_set_tunable(b"k3b_class_major", 0)
    ref = run()
    lib.b200f_set_tunable(b"k3b_class_major", 1)
    for trial in range(6):
        got = run()
        bad = (got - ref).abs() > 1e-3 * ref.abs().max()
        nbad = int(bad.sum())
        print(f"B={B} C={C} trial {trial}: mismatching elements {nbad}  timeout={lib.b200f_umma_timeout_flag(1)}", flush=True)
        if nbad:
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            print("  rows:", rows[:24].tolist(), "... n =", len(rows), " rows mod 32:", sorted(set((rows % 32).tolist()))[:32])
            print("  row blocks (//32):", sorted(set((rows // 32).tolist()))[:40])
            print("  col slices (//32):", sorted(set((cols // 32).tolist())), " n cols =", len(cols))
            r0 = int(rows[0]); c0 = int(bad[r0].nonzero()[0])
            print("  sample row", r0, "col", c0, "got", got[r0, c0:c0 + 4].tolist(), "ref", ref[r0, c0:c0 + 4].tolist())
