"""The AdamW oracle (oracle/adamw_oracle.py) against torch.optim.AdamW -- the optimizer the reference constructs for
the head (src/training.py:343-348) -- on CPU."""
import numpy as np
import pytest
import torch

from oracle import adamw_oracle


@pytest.mark.parametrize("amsgrad", [True, False])
@pytest.mark.parametrize("wd", [0.0, 1e-4, 1e-2])
def test_oracle_matches_torch_adamw(amsgrad, wd):
    g = torch.Generator().manual_seed(3)
    w0 = torch.randn(37, 64, generator=g) * 0.05
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.AdamW([p], lr=3e-3, weight_decay=wd, amsgrad=amsgrad)
    w = w0.numpy().copy(); m = np.zeros_like(w); v = np.zeros_like(w); vmax = np.zeros_like(w) if amsgrad else None
    for step in range(1, 8):
        grad = torch.randn(37, 64, generator=g) * (0.02 if step % 3 else 2.0)      # AMSGrad's max matters
        p.grad = grad.clone()
        opt.step()
        w, m, v, vmax = adamw_oracle.adamw_step(w, grad.numpy(), m, v, vmax, step, lr=3e-3, weight_decay=wd)
        st = opt.state[p]
        # max-norm relative error (elements that cancel to ~0 carry the absolute rounding of their terms)
        rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        assert rel(w, p.detach().numpy()) < 1e-6
        assert rel(m, st["exp_avg"].numpy()) < 1e-6
        assert rel(v, st["exp_avg_sq"].numpy()) < 1e-6
        if amsgrad:
            assert rel(vmax, st["max_exp_avg_sq"].numpy()) < 1e-6
