"""CPU restatement of the optimizer step the reference applies to the head's class weights
(TEST INFRASTRUCTURE ONLY -- tests/ may import this; the product never does).

Reference call sites: ``optim.AdamW(model.parameters(), lr=lr, weight_decay=weight_decay, amsgrad=True)``
/root/reference/src/training.py:343-348 and ``torch.optim.AdamW(..., amsgrad=use_amsgrad)``
src/hyperparameter_tuning.py:114-120; the step runs after the optional ``clip_grad_norm_`` of
src/training.py:528-533.  The arithmetic lives in the third-party dependency torch (requirements.txt:1
``torch>=1.13.0``, unpinned; 2.11.0 here): torch/optim/adamw.py ``_single_tensor_adamw``, restated below in numpy
float32 in the same operation order.  Pinned by tests/test_adamw_oracle.py against torch.optim.AdamW itself (the
very optimizer the reference constructs) on CPU -- that is the golden source; no fixture is needed because torch
travels to the GPU box."""
from __future__ import annotations

import math

import numpy as np


def adamw_step(w, g, m, v, vmax, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
    """One AdamW step (amsgrad when vmax is not None), float32, returns new (w, m, v, vmax).  step counts from 1."""
    f = np.float32
    b1, b2 = f(betas[0]), f(betas[1])
    g = (g.astype(f) * f(grad_scale)).astype(f)
    w = (w.astype(f) * f(1.0 - lr * weight_decay)).astype(f)                    # param.mul_(1 - lr * wd)
    m = (m + (g - m) * f(1.0 - betas[0])).astype(f)                             # exp_avg.lerp_(grad, 1 - beta1)
    v = (v * b2 + f(1.0 - betas[1]) * g * g).astype(f)                          # mul_(beta2).addcmul_(g, g, 1 - beta2)
    bc1 = 1.0 - betas[0] ** step
    bc2_sqrt = math.sqrt(1.0 - betas[1] ** step)
    if vmax is not None:
        vmax = np.maximum(vmax, v).astype(f)
        den = (np.sqrt(vmax) / f(bc2_sqrt) + f(eps)).astype(f)
    else:
        den = (np.sqrt(v) / f(bc2_sqrt) + f(eps)).astype(f)
    w = (w - f(lr / bc1) * (m / den)).astype(f)                                 # addcdiv_(exp_avg, denom, -step_size)
    return w, m, v, vmax
