"""K5 -- the optimizer step of the class weights, fused with next step's K1.

The reference trains the ArcFace head with ``optim.AdamW(model.parameters(), lr, weight_decay, amsgrad=True)``
(src/training.py:343-348; ``create_optimizer`` src/hyperparameter_tuning.py:114-120) after an optional
``clip_grad_norm_`` (src/training.py:528-533).  ``HeadAdamW`` applies exactly that update to ONE parameter -- the
[C, D] fp32 class-weight matrix -- with one kernel (``b200f_head_adamw``) that also emits the normalised fp16 rows and
inverse norms the next forward needs, so the step that follows runs no K1 over W.  Everything else of the model
keeps its ordinary torch optimizer."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

NORM_EPS = 1e-12            # F.normalize's eps (src/face_models.py:351-352)
OPERAND_SCALE = 256.0       # K1's power-of-two operand scale (head.py)


class HeadAdamW:
    """AdamW (+ AMSGrad) for the class-weight matrix of a head.

    ``head``: an ``ArcMarginProduct`` (its ``weight`` is updated and its K1 cache refreshed) or the weight tensor.
    ``step(grad=None, grad_scale=None)``: grad defaults to ``weight.grad``; ``grad_scale`` is an optional 0-dim CUDA
    tensor multiplied into the gradient (the coefficient ``clip_grad_norm_`` would apply).
    ``weight_cache``: pass it to ``GraphedHeadStep`` / ``ArcMarginProduct.graphed_step(optimizer=...)`` -- the captured
    graph then reads the operands this optimizer refreshes in place instead of running K1 over W.
    ``state_dict()`` has torch.optim.AdamW's layout for one parameter."""

    def __init__(self, head, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = True):
        self.head = head if isinstance(head, nn.Module) else None
        weight = head.weight if self.head is not None else head
        require_cuda(weight)
        if weight.dtype != torch.float32 or not weight.is_contiguous() or weight.dim() != 2:
            raise TypeError("HeadAdamW updates a contiguous fp32 [C, D] CUDA matrix")
        if weight.shape[1] % 4 != 0 or weight.shape[1] > 1024:
            raise ValueError("HeadAdamW: D % 4 == 0 and D <= 1024")
        self.weight = weight
        self.lr, self.betas, self.eps, self.weight_decay, self.amsgrad = float(lr), tuple(betas), float(eps), float(weight_decay), bool(amsgrad)
        self.step_count = 0
        w = weight.detach()
        self.exp_avg = torch.zeros_like(w)
        self.exp_avg_sq = torch.zeros_like(w)
        self.max_exp_avg_sq = torch.zeros_like(w) if amsgrad else None
        self.w_hat = torch.empty(w.shape, dtype=torch.float16, device=w.device)
        self.inv_norm = torch.empty(w.shape[0], dtype=torch.float32, device=w.device)
        self.refresh_operands()

    # -- the operands the next forward reads -------------------------------------------------------------------------
    @property
    def weight_cache(self) -> dict:
        return {"static": (self.w_hat, self.inv_norm)}

    def refresh_operands(self):
        """K1 over the current weights (construction, load_state_dict, any out-of-band change of the parameter)."""
        lib = _lib.load_library()
        w = self.weight.detach()
        check(lib.b200f_l2norm_rows(ptr(w), _lib.dtype_code(w), w.shape[0], w.shape[1], NORM_EPS, ptr(self.inv_norm),
                                    ptr(self.w_hat), _lib.F16N, OPERAND_SCALE, stream_ptr(w.device)), "b200f_l2norm_rows")
        self._publish()

    def _publish(self):
        head = self.head
        if head is not None and getattr(head, "cache_weight_prep", False) and hasattr(head, "_w_prep"):
            w = self.weight.detach()
            head._w_prep["key"] = (w.data_ptr(), w._version, tuple(w.shape), w.dtype, True)
            head._w_prep["val"] = (self.w_hat, self.inv_norm)
            # the head may use the cached operands in training mode too: this optimizer keeps them current, and
            # re-derives them when the head's weights are replaced (load_state_dict -> _drop_weight_caches)
            head._w_prep["optimizer_current"] = True
            head._w_prep["static_refresh"] = self.refresh_operands

    # -- the step ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, grad: Optional[torch.Tensor] = None, grad_scale: Optional[torch.Tensor] = None):
        g = grad if grad is not None else self.weight.grad
        if g is None:
            return
        require_cuda(g, grad_scale)
        if g.dtype != torch.float32 or g.shape != self.weight.shape:
            raise TypeError("HeadAdamW.step: the gradient must be fp32 with the weight's shape")
        g = g.contiguous()
        if grad_scale is not None and (grad_scale.dtype != torch.float32 or grad_scale.numel() != 1):
            raise TypeError("HeadAdamW.step: grad_scale must be a 1-element fp32 CUDA tensor")
        self.step_count += 1
        lib = _lib.load_library()
        w = self.weight.detach()
        check(lib.b200f_head_adamw(ptr(w), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.max_exp_avg_sq),
                                   w.shape[0], w.shape[1], self.lr, self.betas[0], self.betas[1], self.eps,
                                   self.weight_decay, self.step_count, ptr(grad_scale), ptr(self.w_hat), OPERAND_SCALE,
                                   NORM_EPS, ptr(self.inv_norm), stream_ptr(w.device)), "b200f_head_adamw")
        torch.autograd.graph.increment_version(self.weight)     # the kernel wrote the parameter behind autograd's back
        self._publish()

    @staticmethod
    def clip_coef(max_norm: float, head_dw_sqnorm: torch.Tensor, other_sqnorm: Optional[torch.Tensor] = None,
                  group=None) -> torch.Tensor:
        """The coefficient ``torch.nn.utils.clip_grad_norm_(params, max_norm)`` multiplies every gradient with
        (src/training.py:528-533: max_norm / (total_norm + 1e-6), clamped to 1), from the head's ||dW||^2 -- a side output
        of the dW epilogues (``head.track_dw_norm = True`` -> ``head.last_stats.dw_sqnorm``; no pass over dW) -- plus the
        squared norm of everything else the caller clips together with it.  group: class shards SUM their parts first.
        Returns a [1] fp32 device tensor: pass it to ``step(grad_scale=...)`` and scale the other gradients with it."""
        total = head_dw_sqnorm.detach().to(torch.float32).reshape(1).clone()
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        if other_sqnorm is not None:
            total = total + other_sqnorm.detach().to(torch.float32).reshape(1)
        return torch.clamp(float(max_norm) / (torch.sqrt(total) + 1e-6), max=1.0)

    def zero_grad(self, set_to_none: bool = True):
        if self.weight.grad is not None:
            if set_to_none:
                self.weight.grad = None
            else:
                self.weight.grad.zero_()

    # -- torch.optim.AdamW-compatible state ------------------------------------------------------------------------------
    def state_dict(self) -> dict:
        st = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}
        if self.amsgrad:
            st["max_exp_avg_sq"] = self.max_exp_avg_sq
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": self.amsgrad, "params": [0]}
        return {"state": {0: st}, "param_groups": [group]}

    def load_state_dict(self, sd: dict):
        group = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(group["lr"]), tuple(group["betas"]), float(group["eps"])
        self.weight_decay, amsgrad = float(group["weight_decay"]), bool(group.get("amsgrad", False))
        if amsgrad != self.amsgrad:
            raise ValueError("HeadAdamW.load_state_dict: amsgrad differs from the constructed optimizer")
        st = sd["state"].get(0)
        if st:
            self.step_count = int(st["step"])
            self.exp_avg.copy_(st["exp_avg"]); self.exp_avg_sq.copy_(st["exp_avg_sq"])
            if self.amsgrad:
                self.max_exp_avg_sq.copy_(st["max_exp_avg_sq"])
        self.refresh_operands()
