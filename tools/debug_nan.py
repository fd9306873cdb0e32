"""Scratch diagnostic for tests/test_gpu_head.py::test_nan_inf_scrub (prints per-row forward statistics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import b200face, oracle
from b200face import _lib
from b200face import head as H

g = torch.Generator().manual_seed(5)
B, C, D = 16, 50, 64
w = torch.randn(C, D, generator=g) * (2.0 / (C + D)) ** 0.5 * 2 ** 0.5
x = torch.randn(B, D, generator=g)
y = torch.randint(0, C, (B,), generator=g)
x[3, 7] = float("inf")
dev = torch.device("cuda:0")
cfg = oracle.HeadConfig(current_epoch=12, training=True, label_smoothing=0.05)
m_eff, s_eff = oracle.effective_margin_scale(cfg)
hc = H._head_cfg(m_eff, s_eff, 0.05, False, C, _lib.ENGINE_AUTO)
out = H._fwd_kernels(x.to(dev), w.to(dev), y.to(dev), hc, 0, True)
xo, wo, inv_nx, inv_nw, row_stats, row_best, row_argmax, cos_minmax, nan_flag, logits = out
torch.cuda.synchronize()
print("inv_nx", inv_nx.cpu().numpy())
print("row_stats\n", row_stats.cpu().numpy())
print("nan_flag", nan_flag.item(), "cos_minmax", cos_minmax.cpu().numpy())
z, _, _, nan_seen = oracle.arc_logits(x.numpy(), w.numpy(), y.numpy(), cfg)
print("oracle row3 logits (first 8)", z[3, :8], "ours", logits[3, :8].cpu().numpy())
se = np.exp(z - s_eff).sum(1)
print("oracle sumexp", se)
print("ours   sumexp", row_stats[:, 0].cpu().numpy())
print("oracle ztarget", z[np.arange(B), y.numpy()])
print("ours   ztarget", row_stats[:, 2].cpu().numpy())
print("oracle sumz", z.sum(1))
print("ours   sumz", row_stats[:, 3].cpu().numpy())
