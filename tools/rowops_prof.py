#!/usr/bin/env python
"""The HBM-bound row kernels at their production sizes, alone: K1 (W bf16 -> w_hat fp16 + 1/||w||, cfg3), K1 inverse
norms only (fp32 W), K5 (AdamW/AMSGrad + next step's K1, cfg3), the normalise-backward of rows, and the gallery
prepare (1 M x 512 fp32 -> fp16 scan operand + bias).  Prints, per kernel, the algorithmic bytes, the live launch
duration (CUDA events on the launching stream, L2 flushed before every launch) and GB/s against MEASURED_PEAKS.json;
under `ncu --set full -k regex:"l2norm|adamw|gallery_prepare"` the same launches give the DRAM bytes.
Development / profiling aid: no oracle, no checks (tests/test_gpu_head.py, test_gpu_adamw.py hold parity)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import b200face
from b200face import _lib
from b200face import head as H

dev = torch.device("cuda:0")
C, D, N = 100_000, 512, int(os.environ.get("GN", 1_000_000))
REPS = int(os.environ.get("REPS", 10))
peak = 6548.2
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
g = torch.Generator(device=dev).manual_seed(7)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)      # 256 MB > the 126 MB L2


def timed(fn):
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def timed_back_to_back(fn, inner=5):
    """`inner` launches per event pair, no flush in between: every working set here is >= 205 MB against a 126 MB L2,
    and the ~8 us an event pair adds around a single short launch (launch latency behind the flush) is amortised."""
    ts = []
    for _ in range(max(2, REPS // 2)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    ts.sort()
    return ts[len(ts) // 2]


out = {}


def report(name, nbytes, fn):
    fn(); torch.cuda.synchronize()
    med, mn = timed(fn)
    b2b = timed_back_to_back(fn) if REPS > 1 else med          # REPS=1: the ncu capture run
    out[name] = {"algorithmic_bytes": int(nbytes), "median_us": round(med * 1e3, 1), "min_us": round(mn * 1e3, 1),
                 "GBps": round(nbytes / (med * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(nbytes / (med * 1e-3) / 1e9 / peak, 4),
                 "back_to_back_us": round(b2b * 1e3, 1), "back_to_back_GBps": round(nbytes / (b2b * 1e-3) / 1e9, 1),
                 "back_to_back_frac": round(nbytes / (b2b * 1e-3) / 1e9 / peak, 4)}
    print(name, out[name], flush=True)


w32 = torch.randn(C, D, generator=g, device=dev) * 0.006
w16 = w32.to(torch.bfloat16)
# K1: bf16 rows in, fp16 normalised rows (x 2^8) + inverse norms out
report("k1_w_bf16_to_f16n", C * D * (2 + 2) + C * 4, lambda: H._k1(w16, True))
# K1 on the fp32 master
report("k1_w_f32_to_f16n", C * D * (4 + 2) + C * 4, lambda: H._k1(w32, True))
# K1, inverse norms only (CUDA-core engine / gallery 'cos')
report("k1_w_f32_inv_only", C * D * 4 + C * 4, lambda: H._k1(w32, False))
# normalise-backward of rows: v (fp16 normalised) + inv + dv_hat fp32 in, dv fp32 out
wh, inv = H._k1(w16, True)
dwh = torch.randn(C, D, generator=g, device=dev)
report("l2norm_bwd_rows", C * D * (2 + 4 + 4) + C * 4, lambda: H._normalize_bwd(wh, inv, dwh))
del dwh, wh, inv
# K5: W, dW, m, v, vmax in; W, m, v, vmax, w_hat16 out (+ inverse norms)
wp = w32.clone()
opt = b200face.HeadAdamW(wp, lr=1e-3, weight_decay=1e-4, amsgrad=True)
grad = torch.randn(C, D, generator=g, device=dev) * 1e-3
report("k5_adamw_amsgrad_k1", C * D * 38 + C * 4, lambda: opt.step(grad))
del opt, grad, wp
# gallery prepare: fp32 rows in, fp16 rows + bias out
G = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device=dev), dim=1)
prep = [None]


def do_prepare():
    prep[0] = b200face.PreparedGallery(G, "l2eps", operand_fmt=_lib.OPERAND_FP16)


report("gallery_prepare_f32_to_f16", N * D * (4 + 2) + N * 4, do_prepare)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
