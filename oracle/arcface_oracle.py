"""numpy restatement of the reference ArcFace head (TEST INFRASTRUCTURE ONLY).

Follows, line by line, /root/reference/src/face_models.py:
  * warm-up schedule ............ ArcMarginProduct.forward  :336-348
  * row normalisation ........... :351-352  (F.normalize, eps=1e-12)
  * cosine logits ............... :355
  * min/max cosine side stats ... :358-360
  * clamp / acos / margin ....... :363-397
  * scale cap and damping ....... :401-412
  * NaN/Inf scrub ............... :423-427
  * ArcFaceNet backward hook .... :538-570   (Frobenius-norm renormalisation)
and the caller's criterion ``nn.CrossEntropyLoss(label_smoothing=eps)``
(/root/reference/src/training.py:341,515).

Nothing here is imported by the product.  The functions are pinned against the
reference itself by tests/golden/make_golden.py -> tests/golden/head_*.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Optional, Tuple

import numpy as np

PI_CLAMP = math.pi - 1e-4          # face_models.py:388
COS_LO = -1.0 + 1e-7               # face_models.py:363
COS_HI = 1.0 - 1e-7
MAX_SCALE = 24.0                   # face_models.py:403
NORM_EPS = 1e-12                   # face_models.py:351


@dataclass
class HeadConfig:
    """State of an ``ArcMarginProduct`` that matters for one forward call
    (ctor defaults: face_models.py:307-322)."""
    s: float = 32.0
    m: float = 0.5
    easy_margin: bool = False
    use_warm_up: bool = True
    warm_up_epochs: int = 10
    margin_factor: float = 0.0
    scale_factor: float = 0.3
    current_epoch: int = 0
    training: bool = True
    label_smoothing: float = 0.05   # training.py:341


def warmup_schedule(cfg: HeadConfig) -> Tuple[float, float]:
    """(margin_factor, scale_factor) after the schedule block, face_models.py:336-348.
    Only touched when training and use_warm_up; otherwise the stored values stand."""
    mf, sf = cfg.margin_factor, cfg.scale_factor
    if cfg.training and cfg.use_warm_up:
        if cfg.current_epoch < cfg.warm_up_epochs:
            progress = cfg.current_epoch / cfg.warm_up_epochs
            mf = min(0.9, progress * progress)
            sf = min(0.8, 0.3 + 0.5 * progress)
        else:
            mf, sf = 0.9, 0.8
    return mf, sf


def effective_margin_scale(cfg: HeadConfig) -> Tuple[float, float]:
    """(m_eff, s_eff) as applied by forward: face_models.py:369 and :401-409."""
    mf, sf = warmup_schedule(cfg)
    m_eff = cfg.m * mf if cfg.training else cfg.m
    s0 = min(cfg.s, MAX_SCALE)
    s_eff = s0 * min(0.8, sf) if cfg.training else s0
    if cfg.m > 0.4 and cfg.training:
        s_eff = s_eff * (0.8 - 0.5 * mf)
    return m_eff, s_eff


def l2_normalize_rows(v: np.ndarray, eps: float = NORM_EPS) -> Tuple[np.ndarray, np.ndarray]:
    """F.normalize(v, p=2, dim=1, eps): v / max(||v||, eps).  Returns (v_hat, norm)."""
    n = np.sqrt(np.sum(v * v, axis=1, keepdims=True))
    d = np.maximum(n, np.asarray(eps, dtype=v.dtype))
    return v / d, d[:, 0]


def _clamp_bounds(dtype) -> Tuple[float, float]:
    # torch.clamp takes python doubles and casts them to the tensor dtype
    lo = np.asarray(COS_LO, dtype=dtype)
    hi = np.asarray(COS_HI, dtype=dtype)
    return lo, hi


def arc_logits(x: np.ndarray, w: np.ndarray, label: np.ndarray, cfg: HeadConfig,
               dtype=np.float64):
    """Scaled logits [B,C] exactly in the reference's op order (face_models.py:351-427).
    Returns (logits, cos_max, cos_min, nan_seen)."""
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    m_eff, s_eff = effective_margin_scale(cfg)
    xh, _ = l2_normalize_rows(x)
    wh, _ = l2_normalize_rows(w)
    cos = xh @ wh.T                                            # :355
    cos_max, cos_min = float(cos.max()), float(cos.min())      # :358-360
    lo, hi = _clamp_bounds(dtype)
    c = np.clip(cos, lo, hi)                                   # :363
    theta = np.arccos(c)                                       # :366
    rows = np.arange(x.shape[0])
    out = c.copy()
    ct = c[rows, label]
    tt = theta[rows, label]
    m_eff_t = np.asarray(m_eff, dtype=dtype)
    if cfg.easy_margin:                                        # :372-384
        phi = np.where(ct > 0, np.cos(tt + m_eff_t), ct)
    else:                                                      # :385-397
        phi = np.cos(np.minimum(np.asarray(PI_CLAMP, dtype=dtype), tt + m_eff_t))
    out[rows, label] = phi
    out = out * np.asarray(s_eff, dtype=dtype)                 # :412
    bad = ~np.isfinite(out)                                    # :423-427
    nan_seen = bool(bad.any())
    if nan_seen:
        out = np.where(bad, np.zeros_like(out), out)
    return out, cos_max, cos_min, nan_seen


def smoothed_cross_entropy(z: np.ndarray, label: np.ndarray, eps: float):
    """nn.CrossEntropyLoss(label_smoothing=eps), mean reduction (training.py:341).
    L_i = lse(z_i) - (1-eps) z_{i,y} - (eps/C) sum_j z_ij.  Returns (loss, lse[B])."""
    B, C = z.shape
    zmax = z.max(axis=1, keepdims=True)
    lse = (zmax + np.log(np.exp(z - zmax).sum(axis=1, keepdims=True)))[:, 0]
    zt = z[np.arange(B), label]
    li = lse - (1.0 - eps) * zt - (eps / C) * z.sum(axis=1)
    return li.mean(dtype=z.dtype), lse


def hook_kappa(n: float, max_grad_norm: float = 1.0, phase: int = 1,
               current_epoch: int = 0) -> float:
    """Scalar the ArcFaceNet backward hook multiplies dL/dt by (face_models.py:538-567).
    n = ||dL/dt||_F.  Returns 1.0 when the hook leaves the gradient alone."""
    thr = max_grad_norm
    if phase == 1:
        thr = min(0.5, max_grad_norm)
    if current_epoch < 10:
        thr = min(thr, 0.5 + 0.05 * current_epoch)
    if n > 3.0:
        thr = min(thr, 0.5)
    if n > thr:
        return thr / (n + 1e-8)
    return 1.0


def _dphi_dc(ct: np.ndarray, m_eff: float, easy: bool, dtype) -> np.ndarray:
    """d(target logit, pre-scale)/d(clamped cosine) on the target column.
    autograd of acos -> (+m, minimum) -> cos: sin(theta+m)/sin(theta); zero where the
    pi-1e-4 clamp is active (torch.minimum routes the gradient to the constant)."""
    theta = np.arccos(ct)
    sin_t = np.sqrt((1.0 - ct) * (1.0 + ct))
    tm = theta + np.asarray(m_eff, dtype=dtype)
    if easy:
        return np.where(ct > 0, np.sin(tm) / sin_t, np.ones_like(ct))
    return np.where(tm < PI_CLAMP, np.sin(tm) / sin_t, np.zeros_like(ct))


def head_forward_backward(x, w, label, cfg: HeadConfig, dtype=np.float64,
                          upstream: float = 1.0, hook: Optional[dict] = None):
    """Closed form of  loss = CE_ls(ArcMarginProduct(x, y), y)  and its gradients.

    hook: None (bare ArcMarginProduct) or dict(max_grad_norm, phase, current_epoch)
    to apply the ArcFaceNet backward hook (active from the 2nd training forward on).
    Returns dict(loss, lse, dx, dw, cos_max, cos_min, gnorm, kappa, argmax)."""
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    B, C = x.shape[0], w.shape[0]
    eps = cfg.label_smoothing
    m_eff, s_eff = effective_margin_scale(cfg)
    z, cos_max, cos_min, _ = arc_logits(x, w, label, cfg, dtype=dtype)
    loss, lse = smoothed_cross_entropy(z, label, eps)
    p = np.exp(z - lse[:, None])
    rows = np.arange(B)
    q = np.full((B, C), eps / C, dtype=dtype)
    q[rows, label] += 1.0 - eps
    g_t = (p - q) * (np.asarray(s_eff * upstream, dtype=dtype) / B)   # dL/dt, t = pre-scale output
    gnorm = float(np.sqrt((g_t.astype(np.float64) ** 2).sum()))
    kappa = 1.0
    if hook is not None:
        kappa = hook_kappa(gnorm, hook.get("max_grad_norm", 1.0), hook.get("phase", 1),
                           hook.get("current_epoch", 0))
        g_t = g_t * np.asarray(kappa, dtype=dtype)
    xh, nx = l2_normalize_rows(x)
    wh, nw = l2_normalize_rows(w)
    cos = xh @ wh.T
    lo, hi = _clamp_bounds(dtype)
    c = np.clip(cos, lo, hi)
    g_c = g_t.copy()                                                   # dL/dc (clamped cosine)
    g_c[rows, label] *= _dphi_dc(c[rows, label], m_eff, cfg.easy_margin, dtype)
    g_cos = np.where((cos >= lo) & (cos <= hi), g_c, np.zeros_like(g_c))  # clamp backward
    dxh = g_cos @ wh
    dwh = g_cos.T @ xh
    dx = (dxh - xh * (xh * dxh).sum(axis=1, keepdims=True)) / nx[:, None]
    dw = (dwh - wh * (wh * dwh).sum(axis=1, keepdims=True)) / nw[:, None]
    return dict(loss=loss, lse=lse, dx=dx, dw=dw, cos_max=cos_max, cos_min=cos_min,
                gnorm=gnorm, kappa=kappa, argmax=z.argmax(axis=1), logits=z, g_cos=g_cos)


def sharded_head_forward_backward(x, w, label, cfg: HeadConfig, n_shards: int,
                                  dtype=np.float64):
    """Partial-FC algebra, serial emulation: classes split contiguously over
    n_shards; each shard yields per-row [sum exp(z - s_eff), target logit, sum z] with the
    constant shift s_eff (logits are bounded by s_eff), the shards' stats are SUMMED
    (the all-reduce), then each shard forms its slice of the gradient."""
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    B, C = x.shape[0], w.shape[0]
    eps = cfg.label_smoothing
    m_eff, s_eff = effective_margin_scale(cfg)
    bounds = [(C * r) // n_shards for r in range(n_shards + 1)]
    rows = np.arange(B)
    stats = np.zeros((B, 3), dtype=dtype)
    shard_z = []
    for r in range(n_shards):
        lo_c, hi_c = bounds[r], bounds[r + 1]
        owned = (label >= lo_c) & (label < hi_c)
        # a shard sees only its classes; rows whose label lives elsewhere get no margin
        loc = np.where(owned, label - lo_c, 0)
        zc = _shard_logits(x, w[lo_c:hi_c], loc, owned, cfg, dtype)
        shard_z.append(zc)
        stats[:, 0] += np.exp(zc - s_eff).sum(axis=1)
        stats[:, 1] += np.where(owned, zc[rows, loc], 0.0)
        stats[:, 2] += zc.sum(axis=1)
    lse = s_eff + np.log(stats[:, 0])
    loss = (lse - (1.0 - eps) * stats[:, 1] - (eps / C) * stats[:, 2]).mean()
    return dict(loss=loss, lse=lse, stats=stats, shard_logits=shard_z, bounds=bounds)


def _shard_logits(x, w_shard, loc, owned, cfg, dtype):
    m_eff, s_eff = effective_margin_scale(cfg)
    xh, _ = l2_normalize_rows(x)
    wh, _ = l2_normalize_rows(w_shard)
    cos = xh @ wh.T
    lo, hi = _clamp_bounds(dtype)
    c = np.clip(cos, lo, hi)
    rows = np.arange(x.shape[0])
    ct = c[rows, loc]
    tt = np.arccos(ct)
    if cfg.easy_margin:
        phi = np.where(ct > 0, np.cos(tt + m_eff), ct)
    else:
        phi = np.cos(np.minimum(PI_CLAMP, tt + m_eff))
    out = c.copy()
    out[rows, loc] = np.where(owned, phi, ct)
    out = out * s_eff
    return np.where(np.isfinite(out), out, 0.0)


def with_(cfg: HeadConfig, **kw) -> HeadConfig:
    return replace(cfg, **kw)
