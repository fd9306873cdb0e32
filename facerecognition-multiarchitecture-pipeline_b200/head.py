"""Host side of the ArcFace head: drop-in mirrors of the reference modules over libb200face.so.

Reference surface mirrored here (same names, arguments, attributes, state_dict keys, errors):
  ArcMarginProduct  /root/reference/src/face_models.py:297-445
  ArcFaceNet        /root/reference/src/face_models.py:447-613   (trunk = torchvision, unchanged)
  criterion         nn.CrossEntropyLoss(label_smoothing=eps)     src/training.py:341,515

New, fused entry: ``ArcMarginProduct.forward_loss(input, label, label_smoothing)`` /
``ArcFaceNet.forward_loss(x, labels, label_smoothing)`` -- loss directly, the B x C logits are never
written (kernels K1-K3, include/b200face.h).  ``forward(input, label) -> logits`` is kept for
callers that need the logits (small C); it runs the same kernels with the logits store enabled.

No CPU path: every call needs CUDA tensors and the built library, otherwise it raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._lib import HeadCfg, HookCfg, check, dtype_code, ptr, require_cuda, stream_ptr

MAX_SCALE = 24.0        # face_models.py:403
NORM_EPS = 1e-12        # face_models.py:351


def head_schedule(current_epoch, warm_up_epochs, use_warm_up, training, margin_factor, scale_factor):
    """Warm-up schedule, face_models.py:336-348.  Returns the (margin_factor, scale_factor) the
    module stores after the call."""
    if training and use_warm_up:
        if current_epoch < warm_up_epochs:
            progress = current_epoch / warm_up_epochs
            margin_factor = min(0.9, progress * progress)
            scale_factor = min(0.8, 0.3 + 0.5 * progress)
        else:
            margin_factor = 0.9
            scale_factor = 0.8
    return margin_factor, scale_factor


def effective_margin_scale(s, m, margin_factor, scale_factor, training):
    """(m_eff, s_eff) exactly as forward applies them, face_models.py:369,401-409."""
    m_eff = m * margin_factor if training else m
    effective_s = min(s, MAX_SCALE)
    s_eff = effective_s * min(0.8, scale_factor) if training else effective_s
    if m > 0.4 and training:
        s_eff = s_eff * (0.8 - 0.5 * margin_factor)
    return m_eff, s_eff


@dataclass
class HeadStats:
    """Device-resident side outputs of one fused forward (read them lazily: .item() syncs)."""
    row_best: Optional[torch.Tensor] = None      # [B] max logit per row (this shard)
    row_argmax: Optional[torch.Tensor] = None    # [B] its global class index
    cos_minmax: Optional[torch.Tensor] = None    # [2] {min, max} raw cosine (face_models.py:358-360)
    nan_flag: Optional[torch.Tensor] = None      # [1] int32, 1 if a logit was scrubbed (:423-427)
    # A caller-owned flag the forward sets instead of a freshly zeroed one: a captured step then holds no fill kernel
    # for it.  It stays set across replays until ArcMarginProduct.nan_seen reads (and clears) it.
    sticky_nan_flag: Optional[torch.Tensor] = None
    hook_out: Optional[torch.Tensor] = None      # [3] grad_scale, ||dL/dt||_F, kappa (after backward)
    lse: Optional[torch.Tensor] = None           # [B]
    dx_f32: Optional[torch.Tensor] = None        # [B,D] fp32 dL/dx before the cast to x.dtype (after backward)
    # ||dW||^2 of this shard's rows as a side output of the dW epilogues (b200f_head_request_dw_sqnorm): set
    # want_dw_sqnorm before the backward; dw_sqnorm is a [1] fp32 device tensor after it (class shards: SUM it over ranks)
    want_dw_sqnorm: bool = False
    dw_sqnorm: Optional[torch.Tensor] = None


@dataclass
class _Hook:
    enabled: bool = False
    max_grad_norm: float = 1.0
    phase: int = 1
    epoch: int = 0


def _head_cfg(m_eff, s_eff, label_smoothing, easy, c_total, engine) -> HeadCfg:
    return HeadCfg(float(m_eff), float(s_eff), float(label_smoothing), int(bool(easy)), int(c_total),
                   int(engine), OPERAND_SCALE)


TCGEN05_MAX_D = 512     # x_hat rows stay resident in shared memory (8 k-blocks of 64): csrc/umma_xw.cuh
OPERAND_SCALE = 256.0   # K1 emits x_hat * 2^8 / w_hat * 2^8 in fp16 for the tcgen05 engine (|x_hat| <= 1)


def _k1(t: torch.Tensor, want_f16n: bool):
    """K1, fused row L2-normalise.  Returns (operand, inv_norm): operand is t itself (CUDA-core engine: the
    1/||.|| factors are applied in the epilogue) or the fp16 normalised rows * OPERAND_SCALE (tcgen05 engine)."""
    lib = _lib.load_library()
    inv = torch.empty(t.shape[0], dtype=torch.float32, device=t.device)
    out = torch.empty(t.shape, dtype=torch.float16, device=t.device) if want_f16n else None
    check(lib.b200f_l2norm_rows(ptr(t), dtype_code(t), t.shape[0], t.shape[1], NORM_EPS, ptr(inv), ptr(out),
                                _lib.F16N, OPERAND_SCALE, stream_ptr(t.device)), "b200f_l2norm_rows")
    return (out if want_f16n else t), inv


def l2_normalize(t: torch.Tensor, out_dtype: Optional[torch.dtype] = None):
    """K1 with the normalised rows materialised: returns (t_hat, inv_norm)."""
    require_cuda(t)
    t = t.contiguous()
    lib = _lib.load_library()
    out = torch.empty(t.shape, dtype=out_dtype or t.dtype, device=t.device)
    inv = torch.empty(t.shape[0], dtype=torch.float32, device=t.device)
    check(lib.b200f_l2norm_rows(ptr(t), dtype_code(t), t.shape[0], t.shape[1], NORM_EPS, ptr(inv), ptr(out),
                                dtype_code(out), 1.0, stream_ptr(t.device)), "b200f_l2norm_rows")
    return out, inv


def use_tcgen05(x: torch.Tensor, engine: int, wants_logits: bool = False) -> bool:
    """Engine choice (include/b200face.h): bf16 inputs go to the tcgen05/TMEM/TMA engine, fp32 inputs to the
    fp32 CUDA-core engine (the 1e-5 bar needs fp32 products) unless the caller forces an engine."""
    if engine == _lib.ENGINE_SIMT or wants_logits or x.shape[1] % 8 != 0 or x.shape[1] > TCGEN05_MAX_D:
        if engine == _lib.ENGINE_TCGEN05:
            raise RuntimeError("the tcgen05 engine needs D % 8 == 0, D <= 512 and does not store logits")
        return False
    if not _lib.load_library().b200f_has_tcgen05():
        if engine == _lib.ENGINE_TCGEN05:
            raise RuntimeError("the tcgen05 engine needs an sm_100 device")
        return False
    return engine == _lib.ENGINE_TCGEN05 or x.dtype == torch.bfloat16


def _weight_key(w: torch.Tensor, f16n: bool):
    return (w.data_ptr(), w._version, tuple(w.shape), w.dtype, f16n)


def _cached_weight(w: torch.Tensor, f16n: bool, cache: Optional[dict]):
    """K1(weight) from the cache, or None.  Reused until the tensor's version changes (optimizer step /
    load_state_dict); "static" = operands an optimizer keeps current in place (optim.HeadAdamW)."""
    if cache is None:
        return None
    if f16n and "static" in cache:
        return cache["static"]
    if cache.get("key") == _weight_key(w, f16n):
        return cache["val"]
    return None


def _prepare_weight(w: torch.Tensor, f16n: bool, cache: Optional[dict]):
    """K1 on the class weights (cached when a cache dict is given: see _cached_weight)."""
    val = _cached_weight(w, f16n, cache)
    if val is not None:
        return val
    with _lib.timed("l2norm_rows_w", w.device):
        val = _k1(w, f16n)
    if cache is not None:
        cache["key"], cache["val"] = _weight_key(w, f16n), val
    return val


def _k1_pair(x: torch.Tensor, w: torch.Tensor, cache: Optional[dict]):
    """K1 over the batch rows AND the class weights in one launch (tcgen05 operands; b200f_l2norm_rows_pair)."""
    lib = _lib.load_library()
    dev = x.device
    xo = torch.empty(x.shape, dtype=torch.float16, device=dev)
    wo = torch.empty(w.shape, dtype=torch.float16, device=dev)
    inv_nx = torch.empty(x.shape[0], dtype=torch.float32, device=dev)
    inv_nw = torch.empty(w.shape[0], dtype=torch.float32, device=dev)
    with _lib.timed("l2norm_rows_w", dev):
        check(lib.b200f_l2norm_rows_pair(ptr(x), x.shape[0], ptr(inv_nx), ptr(xo), ptr(w), w.shape[0], ptr(inv_nw), ptr(wo),
                                         dtype_code(x), x.shape[1], NORM_EPS, _lib.F16N, OPERAND_SCALE, stream_ptr(dev)),
              "b200f_l2norm_rows_pair")
    if cache is not None:
        cache["key"], cache["val"] = _weight_key(w, True), (wo, inv_nw)
    return xo, inv_nx, wo, inv_nw


def _check_head_inputs(x, w, label, dlogits=None):
    """Shape / dtype / device checks the kernels cannot make (they see raw pointers): a mismatch here would be an
    out-of-bounds device read or a host pointer handed to a kernel.  The reference raises a Python error in each of
    these cases too (F.linear shape error, scatter_ index error, device mismatch)."""
    if x.dim() != 2 or w.dim() != 2:
        raise ValueError(f"head: input must be [B, D] and weight [C, D]; got {tuple(x.shape)} and {tuple(w.shape)}")
    if x.shape[1] != w.shape[1]:
        raise ValueError(f"head: input has {x.shape[1]} features, weight has {w.shape[1]}")
    if label.dim() != 1 or label.shape[0] != x.shape[0]:
        raise ValueError(f"head: label must be [B] = [{x.shape[0]}]; got {tuple(label.shape)}")
    if label.dtype != torch.int64:
        raise TypeError(f"head: label must be int64, got {label.dtype}")
    require_cuda(x, w, label, dlogits)
    if not (x.device == w.device == label.device):
        raise RuntimeError(f"head: input ({x.device}), weight ({w.device}) and label ({label.device}) must share a device")
    if not (x.is_contiguous() and w.is_contiguous() and label.is_contiguous()):
        raise ValueError("head: kernels take contiguous tensors")
    if dlogits is not None and (dlogits.dim() != 2 or dlogits.shape[0] != x.shape[0] or dlogits.shape[1] != w.shape[0]
                                or dlogits.device != x.device or not dlogits.is_contiguous()):
        raise ValueError(f"head: dlogits must be a contiguous [B, C] = [{x.shape[0]}, {w.shape[0]}] tensor on {x.device}")


def _fwd_kernels(x, w, label, cfg: HeadCfg, class_offset, want_logits, w_cache=None, nan_flag=None,
                 fused_hook: Optional[HookCfg] = None, x_ops=None):
    """K1 (x and W) + K2.  fused_hook (an unsharded head that wants its loss): K2b and the hook scalars for an upstream
    gradient of 1 come out of the same launch chain (b200f_arcface_fwd_loss); the tuple then ends with
    (lse, loss, pq_norm2, out4)."""
    lib = _lib.load_library()
    _check_head_inputs(x, w, label)
    B, D = x.shape
    C = w.shape[0]
    dev = x.device
    f16n = use_tcgen05(x, cfg.engine, want_logits) if x_ops is None else True
    if not f16n and x.dtype != w.dtype:
        raise TypeError(f"CUDA-core engine: x ({x.dtype}) and weight ({w.dtype}) must share a dtype")
    raw = None                                    # (x_raw or None, w_raw): K1 runs inside the forward call (b200f_arcface_fwd_raw)
    if x_ops is not None:                         # the embedding tail already emitted K1's outputs for x (fused_tail)
        xo, inv_nx = x_ops
        if want_logits or xo.dtype != torch.float16 or tuple(xo.shape) != tuple(x.shape):
            raise ValueError("x_operands: fp16 operand rows of x's shape for the fused (tcgen05) path only")
        if _cached_weight(w, True, w_cache) is None and w.dtype in (torch.float32, torch.bfloat16):
            raw = (None, w)
        else:
            wo, inv_nw = _prepare_weight(w, True, w_cache)
    elif f16n and _cached_weight(w, f16n, w_cache) is None and w.dtype in (torch.float32, torch.bfloat16) \
            and x.dtype in (torch.float32, torch.bfloat16):
        raw = (x, w)                              # K1(x) + K2 with K1(W) inside it: one call, W's fp16 rows made under the MMAs
        xo = torch.empty(x.shape, dtype=torch.float16, device=dev)
        inv_nx = torch.empty(B, dtype=torch.float32, device=dev)
    else:
        xo, inv_nx = _k1(x, f16n)
        wo, inv_nw = _prepare_weight(w, f16n, w_cache)
    if raw is not None:
        wo = torch.empty(w.shape, dtype=torch.float16, device=dev)
        inv_nw = torch.empty(C, dtype=torch.float32, device=dev)
        if w_cache is not None:
            w_cache["key"], w_cache["val"] = _weight_key(w, True), (wo, inv_nw)
    row_stats = torch.empty(B, _lib.STAT_COLS, dtype=torch.float32, device=dev)
    row_best = torch.empty(B, dtype=torch.float32, device=dev)
    row_argmax = torch.empty(B, dtype=torch.int64, device=dev)
    cos_minmax = torch.empty(2, dtype=torch.float32, device=dev)
    if nan_flag is None:
        nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    logits = torch.empty(B, C, dtype=torch.float32, device=dev) if want_logits else None
    nbytes = lib.b200f_head_workspace_bytes(B, C, D, dtype_code(xo), cfg.engine)
    ws = _lib.workspace(nbytes, dev, "head")
    if fused_hook is not None:
        lse = torch.empty(B, dtype=torch.float32, device=dev)
        # separate outputs: the loss is returned as the kernel wrote it (a clone of one slot of a shared buffer was a
        # copy node between the loss and the backward in every captured step)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        pq_norm2 = torch.empty(1, dtype=torch.float32, device=dev)
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
    if raw is not None:
        xr, wr = raw
        hook_p = fused_hook if fused_hook is not None else None
        with _lib.timed("arcface_fwd", dev):
            check(lib.b200f_arcface_fwd_raw(ptr(xr), dtype_code(xr) if xr is not None else 0, ptr(xo), ptr(inv_nx),
                                            ptr(wr), dtype_code(wr), ptr(wo), ptr(inv_nw), NORM_EPS, ptr(label), B, C,
                                            int(class_offset), D, cfg, hook_p, ptr(row_stats), ptr(row_best),
                                            ptr(row_argmax), ptr(cos_minmax), ptr(nan_flag),
                                            ptr(lse) if fused_hook is not None else None,
                                            ptr(loss) if fused_hook is not None else None,
                                            ptr(pq_norm2) if fused_hook is not None else None,
                                            ptr(out4) if fused_hook is not None else None,
                                            ptr(ws), ws.numel(), stream_ptr(dev)), "b200f_arcface_fwd_raw")
        if fused_hook is not None:
            return xo, wo, inv_nx, inv_nw, row_stats, row_best, row_argmax, cos_minmax, nan_flag, logits, lse, loss, pq_norm2, out4
        return xo, wo, inv_nx, inv_nw, row_stats, row_best, row_argmax, cos_minmax, nan_flag, logits
    if fused_hook is not None:
        with _lib.timed("arcface_fwd", dev):
            check(lib.b200f_arcface_fwd_loss(ptr(xo), ptr(wo), dtype_code(xo), ptr(inv_nx), ptr(inv_nw), ptr(label), B, C,
                                             int(class_offset), D, cfg, fused_hook, ptr(row_stats), ptr(row_best),
                                             ptr(row_argmax), ptr(cos_minmax), ptr(nan_flag), ptr(lse), ptr(loss),
                                             ptr(pq_norm2), ptr(out4), ptr(ws), ws.numel(), stream_ptr(dev)),
                  "b200f_arcface_fwd_loss")
        return xo, wo, inv_nx, inv_nw, row_stats, row_best, row_argmax, cos_minmax, nan_flag, logits, lse, loss, pq_norm2, out4
    with _lib.timed("arcface_fwd", dev):
        check(lib.b200f_arcface_fwd(ptr(xo), ptr(wo), dtype_code(xo), ptr(inv_nx), ptr(inv_nw), ptr(label), B, C,
                                    int(class_offset), D, cfg, ptr(row_stats), ptr(row_best), ptr(row_argmax),
                                    ptr(cos_minmax), ptr(nan_flag), ptr(logits), C, ptr(ws), ws.numel(),
                                    stream_ptr(dev)), "b200f_arcface_fwd")
    return xo, wo, inv_nx, inv_nw, row_stats, row_best, row_argmax, cos_minmax, nan_flag, logits


def _raw_rows(x_raw, xo):
    """(pointer, dtype code) of the rows the normalise-backward projects with: the raw input when the operand is K1's
    fp16 copy of it, else nothing (the operand IS the raw input)."""
    if x_raw is not None and x_raw is not xo and x_raw.dtype in (torch.float32, torch.bfloat16):
        return ptr(x_raw), dtype_code(x_raw)
    return None, 0


def _bwd_kernels(xo, wo, label, inv_nx, inv_nw, lse, grad4, cfg: HeadCfg, class_offset, dlogits=None, finish_dx=False,
                 x_raw=None, dx_bf16=False, phase=0, out=None):
    """K3.  finish_dx (unsharded head): dL/dx comes out of the same call (b200f_arcface_bwd_dx) -- returns
    (dxhat, dw, dx, dx_lowp or None).  phase 1 / 2 (class shards, b200f_arcface_bwd_phase): 1 = everything up to a complete
    dx_hat (the caller all-reduces it on another stream), 2 = the remaining dW GEMM into the (dxhat, dw) of phase 1 (`out`)."""
    lib = _lib.load_library()
    _check_head_inputs(xo, wo, label, dlogits)
    B, D = xo.shape
    C = wo.shape[0]
    dev = xo.device
    dxhat, dw = out if out is not None else (torch.empty(B, D, dtype=torch.float32, device=dev),
                                             torch.empty(C, D, dtype=torch.float32, device=dev))
    nbytes = lib.b200f_head_workspace_bytes(B, C, D, dtype_code(xo), cfg.engine)
    ws = _lib.workspace(nbytes, dev, "head")
    if phase:
        with _lib.timed("arcface_bwd", dev):
            check(lib.b200f_arcface_bwd_phase(ptr(xo), ptr(wo), dtype_code(xo), ptr(inv_nx), ptr(inv_nw), ptr(label), ptr(lse),
                                              ptr(grad4), B, C, int(class_offset), D, cfg, ptr(dxhat), ptr(dw), int(phase),
                                              ptr(ws), ws.numel(), stream_ptr(dev)), "b200f_arcface_bwd_phase")
        return dxhat, dw
    if finish_dx:
        dx = torch.empty(B, D, dtype=torch.float32, device=dev)
        lowp = torch.empty(B, D, dtype=torch.bfloat16, device=dev) if dx_bf16 else None
        rp, rd = _raw_rows(x_raw, xo)
        pairs_c = BWD_SIDE_BY_SIDE_PAIRS
        if (pairs_c > 0 and dlogits is None and xo.dtype == torch.float16
                and lib.b200f_arcface_bwd_parts_ok(B, C, D, dtype_code(xo))):
            total = max(2, torch.cuda.get_device_properties(dev).multi_processor_count // 2)
            pairs_c = min(pairs_c, total - 1)

            def part(n, limit):
                check(lib.b200f_arcface_bwd_part(ptr(xo), ptr(wo), dtype_code(xo), ptr(inv_nx), ptr(inv_nw), ptr(label),
                                                 ptr(lse), ptr(grad4), B, C, int(class_offset), D, cfg, ptr(dxhat), ptr(dw),
                                                 rp, rd, ptr(dx), ptr(lowp), n, limit, ptr(ws), ws.numel(), stream_ptr(dev)),
                      "b200f_arcface_bwd_part")
            with _lib.timed("arcface_bwd", dev):
                cur = torch.cuda.current_stream(dev)
                side = _overlap_stream(dev)
                part(1, 0)                                        # K3a: G^T and the r partials
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    part(3, pairs_c)                              # K3c + split reduction + dL/dx
                part(2, total - pairs_c)                          # K3b beside it
                cur.wait_stream(side)
            return dxhat, dw, dx, lowp
        with _lib.timed("arcface_bwd", dev):
            check(lib.b200f_arcface_bwd_dx(ptr(xo), ptr(wo), dtype_code(xo), ptr(inv_nx), ptr(inv_nw), ptr(label), ptr(lse),
                                           ptr(grad4), ptr(dlogits), (dlogits.shape[1] if dlogits is not None else 0),
                                           B, C, int(class_offset), D, cfg, ptr(dxhat), ptr(dw), rp, rd, ptr(dx), ptr(lowp),
                                           ptr(ws), ws.numel(), stream_ptr(dev)), "b200f_arcface_bwd_dx")
        return dxhat, dw, dx, lowp
    with _lib.timed("arcface_bwd", dev):
        check(lib.b200f_arcface_bwd(ptr(xo), ptr(wo), dtype_code(xo), ptr(inv_nx), ptr(inv_nw), ptr(label), ptr(lse),
                                    ptr(grad4), ptr(dlogits), (dlogits.shape[1] if dlogits is not None else 0),
                                    B, C, int(class_offset), D, cfg, ptr(dxhat), ptr(dw), ptr(ws), ws.numel(),
                                    stream_ptr(dev)), "b200f_arcface_bwd")
    return dxhat, dw


def _normalize_bwd(xo, inv_nx, dxhat, x_raw=None, dx_bf16=False):
    """dL/dx = normalise-backward of dxhat, projected with the raw input rows when given (exact x_hat = x * inv_nx)
    else with the operand rows.  dx_bf16: also return the rows rounded to bf16 -> (dx, dx_lowp)."""
    lib = _lib.load_library()
    dx = torch.empty_like(dxhat)
    lowp = torch.empty(dxhat.shape, dtype=torch.bfloat16, device=dxhat.device) if dx_bf16 else None
    v = x_raw if (x_raw is not None and x_raw.dtype in (torch.float32, torch.bfloat16)) else xo
    check(lib.b200f_l2norm_bwd(ptr(v), dtype_code(v), OPERAND_SCALE, ptr(inv_nx), ptr(dxhat), v.shape[0],
                               v.shape[1], ptr(dx), ptr(lowp), stream_ptr(v.device)), "b200f_l2norm_bwd")
    return (dx, lowp) if dx_bf16 else dx


_OVERLAP_STREAMS: dict = {}
# class-sharded backward: run the dx_hat all-reduce on a side stream under the last dW GEMM (B200F_OVERLAP=0: in stream order)
OVERLAP_DXHAT_ALLREDUCE = os.environ.get("B200F_OVERLAP", "1") != "0"
# Unsharded backward at batch <= 512 (one class chunk): the dW GEMM (K3b: bound by its epilogue's latency on every SM, the
# tensor pipe 30 % busy) and the dx GEMM (K3c: bound by the tensor pipe) run SIDE BY SIDE behind K3a, K3c + the split
# reduction + dL/dx on a side stream with this many CTA pairs, K3b on the rest of the chip (b200f_arcface_bwd_part).
# 0 = one after the other (b200f_arcface_bwd_dx).  B200F_BWD_SPLIT overrides.
BWD_SIDE_BY_SIDE_PAIRS = int(os.environ.get("B200F_BWD_SPLIT", "24"))


def _overlap_stream(dev) -> torch.cuda.Stream:
    """The side stream the class-sharded backward runs its dx_hat all-reduce on (one per device, made on first use)."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    s = _OVERLAP_STREAMS.get(key)
    if s is None:
        s = _OVERLAP_STREAMS[key] = torch.cuda.Stream(device=dev)
    return s


class _ArcFaceLossFn(torch.autograd.Function):
    """loss = CE_labelsmooth(ArcMargin(x, w, y), y) fused.
    Unsharded: K1(x, W) -> K2 -> statistics + loss + hook scalars | K3a -> K3b -> K3c -> split reduction + dL/dx.
    Class-sharded: ... K2 -> statistics -> all-reduce [B,4] -> loss + hook scalars | ... K3c -> split reduction ->
    all-reduce [B,D] -> normalise-backward."""

    @staticmethod
    def forward(ctx, x, weight, w, label, cfg: HeadCfg, class_offset, group, hook: _Hook, stats: HeadStats,
                w_cache, unit_upstream, x_ops):
        # weight: the tensor autograd differentiates (fp32 master or already x.dtype);
        # w: what K1 reads (== weight, or a bf16 compute copy of it)
        ctx.w_dtype, ctx.x_dtype = weight.dtype, x.dtype
        x_raw = x
        hk = HookCfg(int(hook.enabled), float(hook.max_grad_norm), int(hook.phase), int(hook.epoch))
        lib = _lib.load_library()
        if group is None:
            (x, w, inv_nx, inv_nw, row_stats, row_best, row_argmax, cos_minmax, nan_flag, _, lse, loss, pq_norm2,
             out4) = _fwd_kernels(x, w, label, cfg, class_offset, False, w_cache, stats.sticky_nan_flag, fused_hook=hk,
                                  x_ops=x_ops)
        else:
            x, w, inv_nx, inv_nw, row_stats, row_best, row_argmax, cos_minmax, nan_flag, _ = _fwd_kernels(
                x, w, label, cfg, class_offset, False, w_cache, stats.sticky_nan_flag, x_ops=x_ops)
            from . import parallel
            parallel.reduce_row_stats(row_stats, group)          # one SUM all-reduce of [B,4]
            B = x.shape[0]
            lse = torch.empty(B, dtype=torch.float32, device=x.device)
            loss = torch.empty((), dtype=torch.float32, device=x.device)
            pq_norm2 = torch.empty(1, dtype=torch.float32, device=x.device)
            out4 = torch.empty(4, dtype=torch.float32, device=x.device)
            check(lib.b200f_arcface_loss_hook(ptr(row_stats), B, cfg, hk, ptr(lse), ptr(loss), ptr(pq_norm2), ptr(out4),
                                              stream_ptr(x.device)), "b200f_arcface_loss_hook")
        stats.row_best, stats.row_argmax = row_best, row_argmax
        stats.cos_minmax, stats.nan_flag, stats.lse = cos_minmax, nan_flag, lse
        ctx.save_for_backward(x, w, label, inv_nx, inv_nw, lse, pq_norm2, out4, x_raw)
        ctx.cfg, ctx.class_offset, ctx.group, ctx.hook, ctx.stats = cfg, class_offset, group, hook, stats
        ctx.unit_upstream = bool(unit_upstream)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        x, w, label, inv_nx, inv_nw, lse, pq_norm2, out4, x_raw = ctx.saved_tensors
        cfg, hook = ctx.cfg, ctx.hook
        lib = _lib.load_library()
        if ctx.unit_upstream:
            grad4 = out4                                         # the forward already formed the scalars for dL/dloss = 1
        else:
            up = grad_out.to(torch.float32).contiguous()
            grad4 = torch.empty(4, dtype=torch.float32, device=x.device)
            check(lib.b200f_arcface_hook_scale(ptr(pq_norm2), ptr(up), x.shape[0], cfg.s_eff, int(hook.enabled),
                                               float(hook.max_grad_norm), int(hook.phase), int(hook.epoch),
                                               ptr(grad4), stream_ptr(x.device)), "b200f_arcface_hook_scale")
        ctx.stats.hook_out = grad4
        if ctx.stats.want_dw_sqnorm:                             # the calling thread's next backward leaves sum(dW^2) here
            ctx.stats.dw_sqnorm = torch.empty(1, dtype=torch.float32, device=x.device)
            check(lib.b200f_head_request_dw_sqnorm(ptr(ctx.stats.dw_sqnorm)), "b200f_head_request_dw_sqnorm")
        want_bf16 = ctx.x_dtype == torch.bfloat16
        if ctx.group is None:
            _, dw, dx, lowp = _bwd_kernels(x, w, label, inv_nx, inv_nw, lse, grad4, cfg, ctx.class_offset, finish_dx=True,
                                           x_raw=x_raw, dx_bf16=want_bf16)
        else:
            from . import parallel
            if OVERLAP_DXHAT_ALLREDUCE and x.is_cuda:
                # dx_hat first; its SUM all-reduce ([B,D] fp32) runs on a side stream under the last dW GEMM of this rank
                dxhat, dw = _bwd_kernels(x, w, label, inv_nx, inv_nw, lse, grad4, cfg, ctx.class_offset, phase=1)
                cur = torch.cuda.current_stream(x.device)
                side = _overlap_stream(x.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    parallel.reduce_dxhat(dxhat, ctx.group)
                _bwd_kernels(x, w, label, inv_nx, inv_nw, lse, grad4, cfg, ctx.class_offset, phase=2, out=(dxhat, dw))
                cur.wait_stream(side)
            else:
                dxhat, dw = _bwd_kernels(x, w, label, inv_nx, inv_nw, lse, grad4, cfg, ctx.class_offset)
                parallel.reduce_dxhat(dxhat, ctx.group)          # one SUM all-reduce of [B,D], in stream order
            if want_bf16:
                dx, lowp = _normalize_bwd(x, inv_nx, dxhat, x_raw, dx_bf16=True)
            else:
                dx, lowp = _normalize_bwd(x, inv_nx, dxhat, x_raw), None
        ctx.stats.dx_f32 = dx
        gx = lowp if lowp is not None else dx.to(ctx.x_dtype)
        return gx, dw.to(ctx.w_dtype), None, None, None, None, None, None, None, None, None, None


class _ArcLogitsFn(torch.autograd.Function):
    """Compatibility path: the scaled logits themselves (ArcMarginProduct.forward).  Backward takes the
    upstream dL/dlogits and runs the same K3 kernels with G = dlogits * s_eff * dphi * clamp-mask."""

    @staticmethod
    def forward(ctx, x, weight, w, label, cfg: HeadCfg, hook: _Hook, stats: HeadStats):
        ctx.w_dtype = weight.dtype
        x, w, inv_nx, inv_nw, _rs, row_best, row_argmax, cos_minmax, nan_flag, logits = _fwd_kernels(
            x, w, label, cfg, 0, True)
        stats.row_best, stats.row_argmax = row_best, row_argmax
        stats.cos_minmax, stats.nan_flag = cos_minmax, nan_flag
        ctx.save_for_backward(x, w, label, inv_nx, inv_nw)
        ctx.cfg, ctx.hook, ctx.stats = cfg, hook, stats
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, w, label, inv_nx, inv_nw = ctx.saved_tensors
        cfg, hook = ctx.cfg, ctx.hook
        dlogits = dlogits.to(torch.float32).contiguous()
        # the legacy hook sees grad_input = dL/d(pre-scale output) = dlogits * s_eff (face_models.py:541)
        n = torch.linalg.vector_norm(dlogits) * cfg.s_eff
        kappa = torch.ones((), dtype=torch.float32, device=x.device)
        if hook.enabled:
            thr = hook.max_grad_norm
            if hook.phase == 1:
                thr = min(0.5, hook.max_grad_norm)
            if hook.epoch < 10:
                thr = min(thr, 0.5 + 0.05 * hook.epoch)
            thr_t = torch.where(n > 3.0, torch.clamp(torch.full_like(n, thr), max=0.5), torch.full_like(n, thr))
            kappa = torch.where(n > thr_t, thr_t / (n + 1e-8), torch.ones_like(n))
        out3 = torch.stack([kappa * cfg.s_eff, n, kappa, torch.ones_like(kappa)]).to(torch.float32)
        ctx.stats.hook_out = out3
        lse_dummy = torch.zeros(1, dtype=torch.float32, device=x.device)
        dxhat, dw = _bwd_kernels(x, w, label, inv_nx, inv_nw, lse_dummy, out3, cfg, 0, dlogits=dlogits)
        dx = _normalize_bwd(x, inv_nx, dxhat)
        ctx.stats.dx_f32 = dx
        return dx.to(x.dtype), dw.to(ctx.w_dtype), None, None, None, None, None


def arcface_loss(x, weight, label, *, m_eff, s_eff, label_smoothing=0.05, easy_margin=False,
                 class_offset=0, num_classes_total=None, group=None, hook: Optional[_Hook] = None,
                 stats: Optional[HeadStats] = None, engine=_lib.ENGINE_AUTO,
                 compute_weight: Optional[torch.Tensor] = None, weight_cache: Optional[dict] = None,
                 unit_upstream: bool = False, x_operands=None):
    """Functional fused head: mean label-smoothed CE of the ArcFace logits of (x, weight).
    x [B,D] fp32 / bf16 CUDA, weight [C_local,D] (the tensor that receives the gradient: an fp32 master
    keeps an fp32 dW even when the kernels compute in bf16), label [B] int64 global ids.
    compute_weight: the copy of weight in x.dtype the kernels read (default: weight itself, or a cast).
    unit_upstream: the caller promises that backward's upstream gradient is exactly 1 (``loss.backward()`` of the bare
    loss; GraphedHeadStep's static root gradient): the hook scalars the forward formed are used as they are and the
    backward launches no scalar kernel.
    x_operands: (x_hat16, inv_norm) already made for x by the fused embedding tail (fused_tail): K1 over x is skipped and
    the tcgen05 engine runs whatever x's dtype (x itself still supplies the rows the backward projects with)."""
    require_cuda(x, weight, label, compute_weight)
    if compute_weight is None:
        if weight.dtype == x.dtype or use_tcgen05(x, engine) or x_operands is not None:
            compute_weight = weight.detach()          # K1 reads the master directly (fp32 or bf16)
        else:
            compute_weight = weight.detach().to(x.dtype)
    cfg = _head_cfg(m_eff, s_eff, label_smoothing, easy_margin,
                    num_classes_total if num_classes_total is not None else weight.shape[0], engine)
    return _ArcFaceLossFn.apply(x.contiguous(), weight, compute_weight.contiguous(),
                                label.contiguous().to(torch.int64), cfg, class_offset, group, hook or _Hook(),
                                stats if stats is not None else HeadStats(), weight_cache, unit_upstream, x_operands)


class LazyArcLogits(torch.Tensor):
    """What ``ArcMarginProduct.forward`` hands an UNMODIFIED reference trainer (``output = model(data, target); loss =
    criterion(output, target); loss.backward()``, src/training.py:508-521; ``_, predicted = outputs.max(1)``,
    src/hyperparameter_tuning.py:1001): a tensor that has the logits' shape, dtype and device but has not been computed.

    * ``F.cross_entropy`` / ``nn.CrossEntropyLoss`` on it with the labels of the forward (mean reduction, no class weights)
      runs the FUSED loss -- the head's own kernels, the B x C logits never stored, the same autograd node as
      ``forward_loss`` (so the ArcFaceNet hook scalars, ``last_stats`` and the tcgen05 engine for bf16 inputs all apply);
    * ``.max(1)`` / ``torch.max(.., 1)`` / ``.argmax(1)`` after that come from the forward's row statistics;
    * ``.detach()`` / ``.data`` stay lazy; any OTHER use materialises the logits once through the compatibility path
      (``_ArcLogitsFn``: the logits are stored, its backward takes dL/dlogits) and goes on with the real tensor.
    ``ArcMarginProduct.lazy_logits = False`` restores the eager compatibility path."""

    @staticmethod
    def __new__(cls, head, x, weight, w, label, m_eff, s_eff):
        r = torch.Tensor._make_wrapper_subclass(cls, (x.shape[0], weight.shape[0]), dtype=torch.float32, device=x.device,
                                                requires_grad=False)
        r._arc = {"head": head, "x": x, "weight": weight, "w": w, "label": label, "m_eff": m_eff, "s_eff": s_eff,
                  "hook": head._hook, "real": None, "loss": None, "stats": None}
        return r

    def __repr__(self):
        return f"LazyArcLogits(shape={tuple(self.shape)}, materialised={self._arc['real'] is not None})"

    # ---- the two ways out -------------------------------------------------------------------------------------------
    def materialise(self):
        """The stored logits (compatibility path), made once."""
        a = self._arc
        if a["real"] is None:
            head = a["head"]
            cfg = _head_cfg(a["m_eff"], a["s_eff"], 0.0, head.easy_margin, head.out_feats, head.engine)
            # the logits-storing kernels are the CUDA-core engine's: a tensor-engine head gets a shadow of x.dtype
            x = a["x"]
            w = a["w"] if a["w"].dtype == x.dtype else a["w"].to(x.dtype)
            head.last_stats = HeadStats()
            a["real"] = _ArcLogitsFn.apply(x, a["weight"], w.contiguous(), a["label"], cfg, a["hook"], head.last_stats)
        return a["real"]

    def fused_loss(self, label_smoothing):
        a = self._arc
        head = a["head"]
        if head.validate_labels:
            c_tot = head._num_classes_total or head.out_feats
            lab = a["label"]
            if lab.numel() and (int(lab.min()) < 0 or int(lab.max()) >= c_tot):
                raise IndexError(f"label out of range [0, {c_tot})")      # the reference's scatter_ raises here (:381)
        head.last_stats = HeadStats(want_dw_sqnorm=bool(getattr(head, "track_dw_norm", False)))
        a["stats"] = head.last_stats
        use_cache = head.cache_weight_prep and (not head.training or head._w_prep.get("optimizer_current", False))
        w = a["weight"].detach() if use_tcgen05(a["x"], head.engine) else a["w"]
        a["loss"] = arcface_loss(a["x"], a["weight"], a["label"], m_eff=a["m_eff"], s_eff=a["s_eff"],
                                 label_smoothing=label_smoothing, easy_margin=head.easy_margin, hook=a["hook"],
                                 stats=head.last_stats, engine=head.engine, compute_weight=w,
                                 class_offset=head._class_offset, num_classes_total=head._num_classes_total,
                                 group=head._group, weight_cache=head._w_prep if use_cache else None)
        return a["loss"]

    def _same_labels(self, target):
        lab = self._arc["label"]
        if not isinstance(target, torch.Tensor) or target.dim() != 1 or target.shape != lab.shape or target.is_floating_point():
            return False
        if target.data_ptr() == lab.data_ptr() and target.dtype == lab.dtype:
            return True
        return target.device == lab.device and bool(torch.equal(target.to(torch.int64), lab))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # below __torch_function__ (which sees every Python-level call first): only reached by C++ callers -- the real logits
        from torch.utils._pytree import tree_map
        swap = lambda v: v.materialise() if isinstance(v, LazyArcLogits) else v
        return func(*tree_map(swap, args), **tree_map(swap, kwargs or {}))

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        lazy = next((a for a in args if isinstance(a, LazyArcLogits)), None)
        if lazy is None:
            lazy = next((v for v in kwargs.values() if isinstance(v, LazyArcLogits)), None)
        name = getattr(func, "__name__", "")
        with torch._C.DisableTorchFunctionSubclass():
            if lazy is not None and lazy._arc["real"] is None:
                if func is F.cross_entropy and args and args[0] is lazy:
                    names = ("input", "target", "weight", "size_average", "ignore_index", "reduce", "reduction", "label_smoothing")
                    kw = dict(zip(names, args)); kw.update(kwargs)
                    if (kw.get("weight") is None and kw.get("reduction", "mean") == "mean" and kw.get("size_average") is None
                            and kw.get("reduce") is None and lazy._same_labels(kw.get("target"))):
                        return lazy.fused_loss(float(kw.get("label_smoothing", 0.0)))
                elif name in ("max", "argmax") and args and args[0] is lazy and lazy._arc["stats"] is not None \
                        and lazy._arc["head"]._group is None:
                    dim = kwargs.get("dim", args[1] if len(args) > 1 else None)
                    keep = kwargs.get("keepdim", args[2] if len(args) > 2 else False)
                    st = lazy._arc["stats"]
                    if isinstance(dim, int) and dim in (1, -1) and not keep and st.row_argmax is not None:
                        if name == "argmax":
                            return st.row_argmax
                        return torch.return_types.max((st.row_best, st.row_argmax))
                elif name in ("detach", "__get__") and (name == "detach" or func == torch.Tensor.data.__get__):
                    return lazy
                elif name in ("size", "dim", "numel", "__len__") or func in (torch.Tensor.shape.__get__, torch.Tensor.dtype.__get__,
                                                                   torch.Tensor.device.__get__, torch.Tensor.requires_grad.__get__,
                                                                   torch.Tensor.is_cuda.__get__, torch.Tensor.ndim.__get__):
                    return func(*args, **kwargs)
            # everything else: on the real logits
            swap = lambda v: v.materialise() if isinstance(v, LazyArcLogits) else v
            return func(*[swap(v) for v in args], **{k: swap(v) for k, v in kwargs.items()})


class GraphedHeadStep:
    """One fused head step (K1-K3: loss = arcface_loss(x, weight, y); loss.backward()) captured once into a CUDA
    graph and replayed per batch.  Why: the step is ~20 short kernels; launched from Python they cost ~0.75 ms of
    host time against ~0.4 ms of GPU time at cfg3 (512 x 100k x 512), so an eager step is launch-bound.

    ``step(x, y)`` copies the batch into the graph's static buffers, replays, and returns the loss (0-dim device
    tensor, no sync).  After the call ``step.dx`` holds dL/dx ([B,D], x's dtype) and ``weight.grad`` (= ``step.dw``)
    dL/dweight -- the same tensors every call, overwritten in place, so an optimizer can keep pointing at them.
    Everything the kernels take by value (m_eff, s_eff, label smoothing, hook state, shapes) is baked in:
    build a new step when the schedule changes them (ArcMarginProduct.graphed_step does that for you)."""

    def __init__(self, weight: torch.Tensor, B: int, D: int, *, dtype=torch.bfloat16, warmup: int = 2, **loss_kw):
        require_cuda(weight)
        dev = weight.device
        self.weight = weight
        loss_kw = dict(loss_kw, unit_upstream=True)   # the root gradient below is the constant 1
        self.loss_kw = loss_kw
        self.x = torch.zeros(B, D, dtype=dtype, device=dev).requires_grad_(True)
        self.y = torch.zeros(B, dtype=torch.int64, device=dev)
        self.stats = HeadStats()
        # A private leaf that shares the parameter's storage: autograd ties a leaf's gradient accumulation to the
        # stream the leaf was first used on; the parameter itself has usually been used on the (uncapturable)
        # default stream already.  Warm-up and capture both run on `side`, so no cross-stream edge is recorded.
        self.w_leaf = weight.detach().requires_grad_(True)
        # static inputs of the capture that torch would otherwise create with a fill kernel inside the graph (each one
        # a launch without the programmatic-serialization attribute between two of ours): the NaN flag K2 sets, and
        # the root gradient of loss.backward()
        self.stats.sticky_nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        self._root_grad = torch.ones((), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                 # warm-up off the capture: library load, workspaces, attributes
            for _ in range(max(1, warmup)):
                self.x.grad = None
                self.w_leaf.grad = None
                arcface_loss(self.x, self.w_leaf, self.y, stats=HeadStats(), **loss_kw).backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.x.grad = None
        self.w_leaf.grad = None                       # so that capture allocates the grads from the graph's pool
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.loss = arcface_loss(self.x, self.w_leaf, self.y, stats=self.stats, **loss_kw)
            torch.autograd.backward(self.loss, grad_tensors=self._root_grad)
        self.dw = self.w_leaf.grad
        self.dx = self.x.grad

    def __call__(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if x.device.type == "cpu":
            return self._call_from_host(x, y)
        self.x.data.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        return self.replay()

    def _call_from_host(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """Host (pinned) batch: the H2D copies run on a private copy stream into one of two staging slots, so the
        copy of batch i+1 overlaps the replay of batch i; the replay stream only adds a device-to-device copy of the
        0.5 MB slot into the graph's static inputs.  As with any non_blocking copy the caller must not overwrite
        the host tensors of a batch before that batch's copy has run (keep two host buffers, or sync)."""
        dev = self.weight.device
        cur = torch.cuda.current_stream(dev)
        if getattr(self, "_stage", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = []
            for _ in range(2):
                free = torch.cuda.Event(); free.record(cur)
                self._stage.append((torch.empty_like(self.x.data), torch.empty_like(self.y), torch.cuda.Event(), free))
            self._slot = 0
        xs, ys, ready, free = self._stage[self._slot]
        self._slot ^= 1
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(free)            # the replay that consumed this slot two batches ago
            xs.copy_(x, non_blocking=True)
            ys.copy_(y, non_blocking=True)
            ready.record(self._copy_stream)
        cur.wait_event(ready)
        self.x.data.copy_(xs, non_blocking=True)
        self.y.copy_(ys, non_blocking=True)
        free.record(cur)
        return self.replay()

    def replay(self) -> torch.Tensor:
        """Re-run on whatever the static buffers hold (inputs already resident)."""
        self.graph.replay()
        self.weight.grad = self.dw                    # the same tensor every call, refreshed in place by the replay
        return self.loss

    def close(self):
        """Wait for the replays in flight and destroy the captured graph.  With a process group in the step the graph
        holds NCCL kernels of that communicator: torch.distributed.destroy_process_group() must not run before this
        (round 1's bench left with os._exit because the teardown blocked)."""
        if getattr(self, "graph", None) is not None:
            torch.cuda.synchronize(self.weight.device)
            self.graph.reset()
            self.graph = None


class ArcMarginProduct(nn.Module):
    """Drop-in for the reference ArcMarginProduct (face_models.py:297-445): same constructor, the
    same externally mutated attributes (s, m, easy_margin, use_warm_up, warm_up_epochs, margin_factor,
    scale_factor, current_epoch -- hyperparameter_tuning.py:829-840 pokes them), same state_dict
    (``weight`` [C,D] fp32, buffer ``u`` [1])."""

    def __init__(self, in_feats, out_feats, s=32.0, m=0.5, use_warm_up=True, easy_margin=False):
        super().__init__()
        self.in_feats = in_feats
        self.out_feats = out_feats
        self.s = s
        self.m = m
        self.easy_margin = easy_margin
        self.use_warm_up = use_warm_up
        self.warm_up_epochs = 10
        self.margin_factor = 0.0
        self.scale_factor = 0.3
        self.current_epoch = 0
        self.weight = nn.Parameter(torch.empty(out_feats, in_feats, dtype=torch.float32))
        nn.init.xavier_normal_(self.weight, gain=math.sqrt(2))
        self.register_buffer('u', torch.zeros(1))
        self.easy_margin_used = False
        self.engine = _lib.ENGINE_AUTO
        self.compute_dtype: Optional[torch.dtype] = None   # None: follow the input's dtype
        self.last_stats = HeadStats()
        self._hook = _Hook()
        self._w_shadow = None
        self._w_prep = {}                  # K1 output for the current weight version
        # Reuse K1(weight) while the parameter is unchanged.  Only in eval mode (or when an optimizer keeps the operands
        # current in place, optim.HeadAdamW): training changes the rows every step, and the (data_ptr, _version) key
        # cannot see writes through ``weight.data`` (EMA updates, manual renormalisation) -- a stale normalised copy
        # would silently train on old weights.  load_state_dict clears the cache (hook below).
        self.cache_weight_prep = True
        self.validate_labels = False       # debug: check 0 <= label < C on the host (one sync), like scatter_ would
        self.lazy_logits = True            # forward() returns a LazyArcLogits: criterion(output, target) runs the fused loss
        # class-parallel placement (parallel.ShardedArcMarginProduct sets these): this module owns the class rows
        # [_class_offset, _class_offset + out_feats) of a [_num_classes_total, D] matrix sharded over _group
        self._class_offset = 0
        self._num_classes_total = None
        self._group = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._drop_weight_caches())

    def _drop_weight_caches(self):
        """Forget every derived copy of the weights (K1 operands, bf16 shadow).  Called after load_state_dict; call it
        yourself after writing through ``weight.data``."""
        static = self._w_prep.get("static_refresh")
        self._w_prep.pop("key", None); self._w_prep.pop("val", None)
        self._w_shadow = None
        if static is not None:
            static()                       # an attached HeadAdamW re-derives its in-place operands

    # -- schedule + effective parameters (host logic, stateful exactly like the reference) --------
    def _step_schedule(self):
        self.margin_factor, self.scale_factor = head_schedule(
            self.current_epoch, self.warm_up_epochs, self.use_warm_up, self.training,
            self.margin_factor, self.scale_factor)
        m_eff, s_eff = effective_margin_scale(self.s, self.m, self.margin_factor, self.scale_factor,
                                              self.training)
        # face_models.py:415-420: the warning branches draw from the CPU generator only when the
        # first two conditions hold; the "rescale" multiplies by 20/20 (a no-op) -- both kept.
        if s_eff > 20.0 and self.training and torch.rand(1).item() < 0.05:
            print(f"Warning: Scale value too high ({s_eff:.1f}) - reducing to prevent instability")
        elif s_eff < 3.0 and self.training and torch.rand(1).item() < 0.05:
            print(f"Warning: Scale value too low ({s_eff:.1f}) - might train slowly")
        if self.easy_margin:
            self.easy_margin_used = True
        return m_eff, s_eff

    def _operands(self, input, wants_logits=False):
        """(x, weight, w_compute): x in the compute dtype, the differentiated parameter, and the tensor K1
        reads.  tcgen05 engine: K1 normalises the fp32 master straight into fp16 operands, no shadow copy.
        CUDA-core engine with a bf16 x and an fp32 master: a bf16 shadow, re-made when the parameter changes."""
        require_cuda(input, self.weight)
        dt = self.compute_dtype or (input.dtype if input.dtype in (torch.float32, torch.bfloat16)
                                    else torch.float32)
        x = input.to(dt).contiguous()
        if self.weight.dtype == dt or use_tcgen05(x, self.engine, wants_logits):
            return x, self.weight, self.weight.detach()
        key = (self.weight._version, self.weight.data_ptr(), dt)
        if self._w_shadow is None or self._w_shadow[0] != key:
            self._w_shadow = (key, self.weight.detach().to(dt))
        return x, self.weight, self._w_shadow[1]

    def forward(self, input, label):
        """Scaled logits [B,C] fp32 -- as a LazyArcLogits (``lazy_logits``, default on): the reference trainer's
        ``criterion(output, target)`` then runs the fused loss and nothing B x C is stored; any other use of the result
        computes and stores the logits (compatibility path)."""
        m_eff, s_eff = self._step_schedule()
        if getattr(self, "lazy_logits", True) and input.is_cuda and input.dim() == 2:
            x, weight, w = self._operands(input)
            return LazyArcLogits(self, x, weight, w.contiguous(), label.contiguous().to(torch.int64), m_eff, s_eff)
        x, weight, w = self._operands(input, wants_logits=True)
        cfg = _head_cfg(m_eff, s_eff, 0.0, self.easy_margin, self.out_feats, self.engine)
        self.last_stats = HeadStats()
        return _ArcLogitsFn.apply(x, weight, w.contiguous(), label.contiguous().to(torch.int64), cfg, self._hook,
                                  self.last_stats)

    def wants_tensor_engine(self, input_dtype=torch.bfloat16) -> bool:
        """Would forward_loss run on the tcgen05 engine for an input of this dtype (D permitting)?"""
        if self.engine == _lib.ENGINE_SIMT or self.in_feats % 8 != 0 or self.in_feats > TCGEN05_MAX_D:
            return False
        if not _lib.load_library().b200f_has_tcgen05():
            return False
        dt = self.compute_dtype or input_dtype
        return self.engine == _lib.ENGINE_TCGEN05 or dt == torch.bfloat16

    def forward_loss(self, input, label, label_smoothing=0.05, return_pred=False, x_operands=None):
        """Fused criterion(forward(input,label), label) with nn.CrossEntropyLoss(label_smoothing):
        the logits never reach HBM.  return_pred: also return outputs.max(1) indices
        (hyperparameter_tuning.py:1001).  x_operands: see arcface_loss (ArcFaceNet's fused tail passes them)."""
        m_eff, s_eff = self._step_schedule()
        if x_operands is not None:
            require_cuda(input, self.weight)
            x, weight, w = input.contiguous(), self.weight, self.weight.detach()
        else:
            x, weight, w = self._operands(input)
        if self.validate_labels:
            c_tot = self._num_classes_total or self.out_feats
            if label.numel() and (int(label.min()) < 0 or int(label.max()) >= c_tot):
                raise IndexError(f"label out of range [0, {c_tot})")      # the reference's scatter_ raises here (:381)
        self.last_stats = HeadStats(want_dw_sqnorm=bool(getattr(self, "track_dw_norm", False)))
        use_cache = self.cache_weight_prep and (not self.training or self._w_prep.get("optimizer_current", False))
        loss = arcface_loss(x, weight, label, m_eff=m_eff, s_eff=s_eff, label_smoothing=label_smoothing,
                            easy_margin=self.easy_margin, hook=self._hook, stats=self.last_stats,
                            engine=self.engine, compute_weight=w, class_offset=self._class_offset,
                            num_classes_total=self._num_classes_total, group=self._group,
                            weight_cache=self._w_prep if use_cache else None, x_operands=x_operands)
        if return_pred:
            if self._group is not None:
                from . import parallel
                return loss, parallel.merge_row_argmax(self.last_stats.row_best, self.last_stats.row_argmax, self._group)[1]
            return loss, self.last_stats.row_argmax
        return loss

    def graphed_step(self, B, label_smoothing=0.05, dtype=torch.bfloat16, optimizer=None):
        """CUDA-graph version of ``loss = forward_loss(x, y, label_smoothing); loss.backward()`` for a fixed batch
        size: returns ``step`` with ``loss = step(x, y)``, ``step.dx`` and ``self.weight.grad`` filled in place.
        The graph bakes in the schedule's (m_eff, s_eff) and the hook state; it is rebuilt when they change
        (once per epoch during the warm-up, never afterwards).
        optimizer: an ``optim.HeadAdamW`` on this head -- the graph then reads the normalised operands that optimizer
        refreshes with every ``step()`` and contains no K1 over the weights."""
        cache = self.__dict__.setdefault("_graphed", {})
        extra = {} if optimizer is None else {"weight_cache": optimizer.weight_cache}

        def step(x, y):
            m_eff, s_eff = self._step_schedule()
            hook = self._hook
            key = (B, float(label_smoothing), dtype, m_eff, s_eff, bool(self.easy_margin), self.engine, hook.enabled,
                   hook.max_grad_norm, hook.phase, hook.epoch, self.weight.data_ptr(), id(optimizer),
                   self._class_offset, self._num_classes_total, id(self._group))
            g = cache.get("step")
            if g is None or cache.get("key") != key:
                if g is not None:
                    g.close()
                g = GraphedHeadStep(self.weight, B, self.in_feats, dtype=dtype, m_eff=m_eff, s_eff=s_eff,
                                    label_smoothing=label_smoothing, easy_margin=self.easy_margin,
                                    hook=_Hook(hook.enabled, hook.max_grad_norm, hook.phase, hook.epoch),
                                    engine=self.engine, class_offset=self._class_offset,
                                    num_classes_total=self._num_classes_total, group=self._group, **extra)
                cache["step"], cache["key"] = g, key
            self.last_stats = g.stats
            step.dx = g.dx
            return g(x, y)

        def close():
            g = cache.pop("step", None)
            cache.pop("key", None)
            if g is not None:
                g.close()

        step.dx = None
        step.close = close
        return step

    def release_graphs(self):
        """Destroy every CUDA graph this head captured (graphed_step).  Call before
        torch.distributed.destroy_process_group(): a captured NCCL all-reduce keeps the communicator busy."""
        cache = self.__dict__.get("_graphed", {})
        g = cache.pop("step", None)
        cache.pop("key", None)
        if g is not None:
            g.close()

    def update_epoch(self, epoch):
        self.current_epoch = epoch

    # reference attributes that used to be filled with .item() syncs on every forward (:358-360):
    # now read lazily from the device
    @property
    def max_cos_theta(self):
        t = self.last_stats.cos_minmax
        return float(t[1].item()) if t is not None else 0.0

    @property
    def min_cos_theta(self):
        t = self.last_stats.cos_minmax
        return float(t[0].item()) if t is not None else 0.0

    @property
    def nan_seen(self):
        t = self.last_stats.nan_flag
        seen = bool(t.item()) if t is not None else False
        if seen:
            print("Uh oh! NaN or Inf in ArcFace output!")      # face_models.py:427
            if t is self.last_stats.sticky_nan_flag:
                t.zero_()                                      # a captured step's flag: set until read
        return seen

    def get_margin_stats(self):
        return {
            'margin_factor': self.margin_factor,
            'scale_factor': self.scale_factor,
            'effective_margin': self.m * self.margin_factor,
            'effective_scale': self.s * self.scale_factor,
            'max_cos_theta': self.max_cos_theta,
            'min_cos_theta': self.min_cos_theta,
            'easy_margin_used': self.easy_margin_used if self.easy_margin else False,
        }


class _FusedTailFn(torch.autograd.Function):
    """y = dropout(BatchNorm1d(z)) -- and, in the same pass over z, what the head's K1 would make of y: 1 / |y| and the
    fp16 operand rows (kernels bn_stats / tail_fwd / tail_bwd, include/b200face.h).  src/face_models.py:516-525."""

    @staticmethod
    def forward(ctx, z, gamma, beta, running_mean, running_var, training, momentum, eps, mask, keep_scale, want_ops,
                want_emb, side):
        lib = _lib.load_library()
        z = z.contiguous()
        if z.dtype not in (torch.float32, torch.bfloat16):
            z = z.float()
        B, D = z.shape
        dev = z.device
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        if training:
            if B < 2:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(z.shape)}")
            mean = torch.empty(D, dtype=torch.float32, device=dev)
            invstd = torch.empty(D, dtype=torch.float32, device=dev)
            check(lib.b200f_bn_stats(ptr(z), dtype_code(z), B, D, float(eps), float(momentum), ptr(running_mean),
                                     ptr(running_var), ptr(mean), ptr(invstd), stream_ptr(dev)), "b200f_bn_stats")
            stat, stat_is_var = invstd, 0
        else:
            mean, stat, stat_is_var = running_mean.float().contiguous(), running_var.float().contiguous(), 1
        y = torch.empty(B, D, dtype=torch.float32, device=dev)
        xo = torch.empty(B, D, dtype=torch.float16, device=dev) if want_ops else None
        emb = torch.empty(B, D, dtype=torch.float32, device=dev) if want_emb else None
        inv = torch.empty(B, dtype=torch.float32, device=dev)
        check(lib.b200f_tail_fwd(ptr(z), dtype_code(z), B, D, ptr(g32), ptr(b32), ptr(mean), ptr(stat), stat_is_var,
                                 float(eps), ptr(mask), float(keep_scale), NORM_EPS, OPERAND_SCALE, ptr(y), ptr(xo), ptr(emb),
                                 ptr(inv), stream_ptr(dev)), "b200f_tail_fwd")
        side["x_operands"] = (xo, inv) if want_ops else None
        side["emb"] = emb
        side["inv_norm"] = inv
        invstd_b = stat if not stat_is_var else torch.rsqrt(stat + eps)
        ctx.save_for_backward(z, g32, mean, invstd_b, mask if mask is not None else torch.empty(0, device=dev))
        ctx.has_mask, ctx.keep_scale, ctx.training, ctx.eps = mask is not None, float(keep_scale), bool(training), float(eps)
        ctx.z_dtype, ctx.g_dtype = z.dtype, gamma.dtype
        ctx.mark_non_differentiable(*(t for t in ()))
        return y

    @staticmethod
    def backward(ctx, dy):
        z, g32, mean, invstd, mask = ctx.saved_tensors
        lib = _lib.load_library()
        B, D = z.shape
        dev = z.device
        dy = dy.float().contiguous()
        dgamma = torch.empty(D, dtype=torch.float32, device=dev)
        dbeta = torch.empty(D, dtype=torch.float32, device=dev)
        dz = torch.empty(B, D, dtype=torch.float32, device=dev)
        check(lib.b200f_tail_bwd(ptr(dy), ptr(mask) if ctx.has_mask else None, ctx.keep_scale, ptr(z), dtype_code(z), ptr(mean),
                                 ptr(invstd), 0, ctx.eps, ptr(g32), int(ctx.training), B, D, ptr(dgamma), ptr(dbeta), ptr(dz),
                                 stream_ptr(dev)), "b200f_tail_bwd")
        return (dz.to(ctx.z_dtype), dgamma.to(ctx.g_dtype), dbeta.to(ctx.g_dtype), None, None, None, None, None, None, None,
                None, None, None)


def fused_tail(z: torch.Tensor, bn: nn.BatchNorm1d, p_drop: float = 0.0, training: bool = True,
               mask: Optional[torch.Tensor] = None, want_operands: bool = False, want_emb: bool = False):
    """The embedding tail of ArcFaceNet behind the Linear layer (src/face_models.py:516-525) as ONE fused pass:
    BatchNorm1d (train: batch statistics + running-stat update; eval: running statistics) -> dropout (train; `mask` =
    uint8 keep mask [B, D], drawn here from torch's CUDA generator when None) -> row L2 norm.
    Returns (y, side): y [B, D] fp32 = dropout(bn(z)), differentiable w.r.t. z and the BatchNorm affine parameters;
    side = {"x_operands": (y_hat16, inv_norm) for ArcMarginProduct.forward_loss, "emb": normalised rows, "inv_norm"}."""
    require_cuda(z, bn.weight, bn.bias, bn.running_mean, bn.running_var)
    keep_scale = 1.0
    if training and p_drop > 0.0:
        if mask is None:
            mask = (torch.rand(z.shape, device=z.device) >= p_drop).to(torch.uint8)
        keep_scale = 1.0 / (1.0 - p_drop) if p_drop < 1.0 else 0.0
    else:
        mask = None
    if mask is not None:
        if mask.dtype != torch.uint8 or tuple(mask.shape) != tuple(z.shape) or not mask.is_cuda:
            raise ValueError("mask: uint8 CUDA tensor of z's shape")
        mask = mask.contiguous()
    momentum = 0.1 if bn.momentum is None else bn.momentum
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    side = {}
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    y = _FusedTailFn.apply(z, bn.weight, bn.bias, rm, rv, training or rm is None, momentum, bn.eps, mask, keep_scale,
                           want_operands, want_emb, side)
    return y, side


class ArcFaceNet(nn.Module):
    """Drop-in for the reference ArcFaceNet (face_models.py:447-613).  The ResNet18 trunk is
    torchvision's (out of scope); the head tail -- embedding, bn, dropout, normalise, ArcFace, and
    the backward-hook gradient renormalisation (:538-570) -- is mirrored over the fused kernels.
    state_dict keys match the reference so its checkpoints load strictly."""

    def __init__(self, num_classes=18, dropout_rate=0.2, s=32.0, m=0.5, easy_margin=False,
                 backbone: Optional[nn.Module] = None, pretrained: bool = False):
        super().__init__()
        if backbone is None:
            import torchvision.models as models
            weights = models.ResNet18_Weights.IMAGENET1K_V1 if pretrained else None
            backbone = models.resnet18(weights=weights)
        self.backbone = backbone
        self.features = nn.Sequential(*list(self.backbone.children())[:-1])
        self.embedding = nn.Linear(512, 512, bias=False)
        self.bn = nn.BatchNorm1d(512, eps=1e-5)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.arcface = ArcMarginProduct(512, num_classes, s=s, m=m, use_warm_up=True, easy_margin=easy_margin)
        self.max_grad_norm = 1.0
        self.current_epoch = 0
        self.phase = 1
        self.backbone_frozen = False
        self.val_classifier = nn.Linear(512, num_classes)
        nn.init.xavier_normal_(self.val_classifier.weight, gain=math.sqrt(2))
        self._hook_armed = False          # the reference registers its hook after the 1st forward
        self.fused_tail = True            # CUDA: BatchNorm + dropout + row norm (+ the head's K1) as one fused kernel

    def freeze_backbone(self):
        self.backbone_frozen = True
        self.phase = 1
        for param_name, param in self.named_parameters():
            if 'backbone' in param_name or 'features' in param_name:
                param.requires_grad = False

    def unfreeze_backbone(self):
        self.backbone_frozen = False
        self.phase = 2
        for param in self.parameters():
            param.requires_grad = True

    def set_max_grad_norm(self, max_norm):
        self.max_grad_norm = max_norm

    def _tail(self, x, training, normalize=True):
        """features -> Linear(512,512) -> BatchNorm1d -> dropout (train) -> F.normalize (src/face_models.py:516-525).
        normalize=False returns the row BEFORE the L2 normalise: the head's K1 normalises its input anyway
        (face_models.py:351), and normalising a unit vector again is the identity in value and in gradient, so the
        training paths hand the head the un-normalised row and the reference's double normalise (:525 + :351)
        costs one fused kernel each way instead of a chain of eight torch kernels."""
        x = self.features(x)
        x = x.view(x.size(0), -1)
        x = self.embedding(x)
        x = self.bn(x)
        if training:
            x = self.dropout(x)
        if not normalize:
            return x
        if not torch.is_grad_enabled() and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):
            return l2_normalize(x)[0]                         # K1 (no autograd needed)
        return F.normalize(x, p=2, dim=1, eps=1e-12)

    def _tail_fused(self, x, training, want_operands=False, want_emb=False):
        """CUDA path of the tail: trunk and Linear stay torch modules, BatchNorm + dropout + row norm run as the fused
        kernel."""
        z = self.embedding(self.features(x).view(x.size(0), -1))
        return fused_tail(z, self.bn, self.dropout.p, training, want_operands=want_operands, want_emb=want_emb)

    def _arm_hook(self):
        # hook state as the reference's closure would read it at backward time (:544-556)
        self.arcface._hook = _Hook(enabled=self._hook_armed, max_grad_norm=self.max_grad_norm,
                                   phase=self.phase, epoch=self.current_epoch)
        self._hook_armed = True

    def forward(self, x, labels=None):
        if self.training:
            if labels is None:
                raise ValueError("Labels must be provided during training")
            if x.is_cuda and self.fused_tail:
                pre = self._tail_fused(x, True)[0]               # CUDA: fused tail; the head's K1 does the (single) normalise
            else:
                pre = self._tail(x, True, normalize=not x.is_cuda)
            self.arcface.update_epoch(self.current_epoch)
            self._arm_hook()
            return self.arcface(pre, labels)
        emb = self._tail(x, False)
        self.val_classifier.weight.data = F.normalize(self.val_classifier.weight.data, p=2, dim=1, eps=1e-12)
        if labels is not None:
            return self.val_classifier(emb)
        return emb

    def forward_loss(self, x, labels, label_smoothing=0.05, return_pred=False):
        """Fused training step head: == criterion(self(x, labels), labels) incl. the hook."""
        if labels is None:
            raise ValueError("Labels must be provided during training")
        ops = None
        if x.is_cuda and self.fused_tail:
            # the head's K1 over x is fused into the tail when the head will run on the tcgen05 engine
            want = self.arcface.wants_tensor_engine(torch.float32) and self.arcface._group is None
            pre, side = self._tail_fused(x, self.training, want_operands=want)
            ops = side["x_operands"]
        else:
            pre = self._tail(x, self.training, normalize=not x.is_cuda)
        self.arcface.update_epoch(self.current_epoch)
        self._arm_hook()
        return self.arcface.forward_loss(pre, labels, label_smoothing, return_pred, x_operands=ops)

    def get_embedding(self, x):
        if x.is_cuda and self.fused_tail and not torch.is_grad_enabled():
            return self._tail_fused(x, False, want_emb=True)[1]["emb"]
        return self._tail(x, False)                           # no dropout (src/face_models.py:584-590)

    def update_epoch(self, epoch):
        self.current_epoch = epoch
        self.arcface.update_epoch(epoch)

    @property
    def last_grad_norm(self):
        t = self.arcface.last_stats.hook_out
        return float(t[1].item()) if (t is not None and self.arcface._hook.enabled) else 0.0

    def get_arcface_stats(self):
        stats = self.arcface.get_margin_stats()
        stats['grad_norm'] = self.last_grad_norm
        stats['max_grad_norm'] = self.max_grad_norm
        stats['phase'] = self.phase
        stats['backbone_frozen'] = self.backbone_frozen
        return stats

    def get_training_phase(self):
        return {'phase': self.phase, 'backbone_frozen': self.backbone_frozen, 'epoch': self.current_epoch}
