"""ctypes binding of libb200face.so (include/b200face.h).

There is NO fallback: if the library is missing, or a call returns an error, this raises.
The reference has no FFI of its own (pure Python, src/face_models.py, src/app.py); these are
the entry points a ctypes binding of its head / gallery-match surface calls."""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_void_p)

import torch

F32, BF16, F16N = 0, 1, 2
METRIC_L2EPS, METRIC_COS = 0, 1
OPERAND_BF16, OPERAND_FP16 = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2
STAT_SUMEXP, STAT_SUMEXP2, STAT_ZTARGET, STAT_SUMZ, STAT_COLS = 0, 1, 2, 3, 4

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libb200face.so"


class LibraryMissingError(RuntimeError):
    pass


class HookCfg(Structure):
    """struct b200f_hook_cfg"""
    _fields_ = [("enabled", c_int32), ("max_grad_norm", c_float), ("phase", c_int32), ("epoch", c_int32)]


class HeadCfg(Structure):
    """struct b200f_head_cfg"""
    _fields_ = [("m_eff", c_float), ("s_eff", c_float), ("label_smoothing", c_float),
                ("easy_margin", c_int32), ("num_classes_total", c_int64), ("engine", c_int32),
                ("operand_scale", c_float)]


def lib_path() -> str:
    """The in-tree library, or the one B200FACE_LIB names (tools/ load instrumented builds of the same sources)."""
    return os.environ.get("B200FACE_LIB") or os.path.join(_HERE, _LIB_NAME)


_lib = None

# name -> (restype, argtypes): exactly the declarations of include/b200face.h
PROTOTYPES = {
    "b200f_version": (c_int, []),
    "b200f_last_error": (c_char_p, []),
    "b200f_launch_count": (ctypes.c_ulonglong, []),
    "b200f_has_tcgen05": (c_int, []),
    "b200f_l2norm_rows": (c_int, [c_void_p, c_int, c_int64, c_int, c_float, c_void_p, c_void_p, c_int,
                                  c_float, c_void_p]),
    "b200f_head_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "b200f_arcface_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_int64, c_int64, c_int, POINTER(HeadCfg),
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "b200f_arcface_loss": (c_int, [c_void_p, c_int64, POINTER(HeadCfg), c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "b200f_arcface_hook_scale": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_int, c_float, c_int,
                                         c_int, c_void_p, c_void_p]),
    "b200f_arcface_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, POINTER(HeadCfg),
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200f_arcface_bwd_phase": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int64, c_int64, c_int64, c_int, POINTER(HeadCfg), c_void_p, c_void_p, c_int,
                                        c_void_p, c_size_t, c_void_p]),
    "b200f_arcface_bwd_parts_ok": (c_int, [c_int64, c_int64, c_int, c_int]),
    "b200f_arcface_bwd_part": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int64, c_int64, c_int, POINTER(HeadCfg), c_void_p, c_void_p,
                                       c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "b200f_head_request_dw_sqnorm": (c_int, [c_void_p]),
    "b200f_l2norm_bwd": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p, c_int64, c_int, c_void_p,
                                 c_void_p, c_void_p]),
    "b200f_l2norm_rows_pair": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                       c_int, c_int, c_float, c_int, c_float, c_void_p]),
    "b200f_arcface_fwd_loss": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int64, c_int64, c_int, POINTER(HeadCfg), POINTER(HookCfg),
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200f_arcface_fwd_raw": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float,
                                      c_void_p, c_int64, c_int64, c_int64, c_int, POINTER(HeadCfg), POINTER(HookCfg),
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200f_arcface_loss_hook": (c_int, [c_void_p, c_int64, POINTER(HeadCfg), POINTER(HookCfg), c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "b200f_arcface_bwd_dx": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, POINTER(HeadCfg),
                                     c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "b200f_bn_stats": (c_int, [c_void_p, c_int, c_int64, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "b200f_tail_fwd": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float,
                               c_void_p, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200f_tail_bwd": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p,
                               c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200f_gallery_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int, c_int]),
    "b200f_gallery_topk": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64,
                                   c_int64, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200f_gallery_has_tc": (c_int, [c_int]),
    "b200f_gallery_prepare": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b200f_gallery_tc_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "b200f_gallery_topk_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                      c_int64, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "b200f_umma_selftest": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200f_umma_timeout_flag": (c_int, [c_int]),
    "b200f_umma_xw_selftest": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200f_umma_set_pair": (c_int, [c_int]),
    "b200f_umma_xw_probe": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200f_set_tunable": (c_int, [c_char_p, c_int]),
    "b200f_stage_ms": (c_int, [c_char_p, c_void_p]),
    "b200f_head_adamw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_double, c_double,
                                 c_double, c_double, c_double, c_int64, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                 c_void_p]),
    "b200f_gallery_merge": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
}


def load_library():
    """dlopen libb200face.so (built by __graft_entry__.build()).  Raises LibraryMissingError if it is
    not there -- there is no CPU or PyTorch fallback for the hot path."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise LibraryMissingError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  b200face has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)       # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().b200f_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16N        # fp16 only ever appears as K1's normalised operand format
    raise TypeError(f"b200face kernels take float32 or bfloat16, got {t.dtype}")


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b200face kernels need CUDA tensors (no CPU fallback); got a "
                               f"{t.device} tensor")


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


# Per-call CUDA-event timing of the C-ABI entry points (bench.py's roofline measurement).  Events are
# recorded on the launching stream; nothing synchronises until the caller reads them.
PROFILE = False
TIMERS = {}


class timed:
    def __init__(self, name, device):
        self.name, self.device = name, device

    def __enter__(self):
        if PROFILE:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if PROFILE:
            self.e1.record(torch.cuda.current_stream(self.device))
            TIMERS.setdefault(self.name, []).append((self.e0, self.e1))
        return False


_workspaces = {}


def workspace(nbytes: int, device, tag: str = "ws") -> torch.Tensor:
    """Caller-owned scratch, cached per (device, stream, tag) so steady-state calls allocate
    nothing.  (Stream-keyed: two streams never share scratch.)"""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf
