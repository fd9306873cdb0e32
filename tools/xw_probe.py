#!/usr/bin/env python
"""Probe the X-stationary tcgen05 kernel (b200f_umma_xw_selftest) on a real B200: single-CTA and CTA-pair
(cta_group::2) variants against torch fp32 matmul on the same fp16 operands.  Each case runs in its own
subprocess under a timeout so a hung or faulting kernel cannot take the probe (or the box) down.
Development aid, not part of the product.  Results: gpurun_out/xw_probe.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(c):
    import torch
    import b200face
    from b200face import _lib
    lib = b200face.load_library()
    dev = torch.device("cuda:0")
    B, C, D, pair = c["B"], c["C"], c["D"], c["pair"]
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, D, generator=g, device=dev).half()
    w = torch.randn(C, D, generator=g, device=dev).half()
    ref = x.float() @ w.float().t()
    out = torch.full((B, C), float("nan"), device=dev)
    rc = lib.b200f_umma_xw_selftest(_lib.ptr(x), _lib.ptr(w), _lib.ptr(out), B, C, D, pair, _lib.stream_ptr(dev))
    res = dict(c, rc=rc)
    if rc != 0:
        res["msg"] = (lib.b200f_last_error() or b"").decode()
        return res
    try:
        torch.cuda.synchronize()
        res["err"] = float((out - ref).norm() / ref.norm())
        res["nan"] = int(torch.isnan(out).sum())
        if c.get("time"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                lib.b200f_umma_xw_selftest(_lib.ptr(x), _lib.ptr(w), _lib.ptr(out), B, C, D, pair, _lib.stream_ptr(dev))
            e1.record(); torch.cuda.synchronize()
            res["ms"] = e0.elapsed_time(e1) / 5
    except RuntimeError as e:
        res["fault"] = str(e)[:300]
    res["timeout_flag"] = lib.b200f_umma_timeout_flag(1)
    return res


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        print("RESULT " + json.dumps(run_case(json.loads(sys.argv[2]))))
        return
    cases = []
    for pair in (1, 2):
        cases += [dict(B=128 * pair, C=128 * pair, D=64, pair=pair), dict(B=128 * pair, C=128 * pair, D=512, pair=pair),
                  dict(B=512, C=4096, D=512, pair=pair), dict(B=300, C=1000, D=200, pair=pair),
                  dict(B=1024, C=148 * 300, D=512, pair=pair, time=1), dict(B=8, C=8, D=8, pair=pair)]
    results = []
    for c in cases:
        try:
            p = subprocess.run([sys.executable, __file__, "--case", json.dumps(c)], capture_output=True, text=True, timeout=120)
            lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
            r = json.loads(lines[-1][7:]) if lines else dict(c, crashed=p.returncode, stderr=p.stderr[-400:])
        except subprocess.TimeoutExpired:
            r = dict(c, hung=True)
        print(r, flush=True)
        results.append(r)
        if r.get("hung"):
            break                      # the GPU may be wedged: stop here
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "xw_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
