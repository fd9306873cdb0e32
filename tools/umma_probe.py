#!/usr/bin/env python
"""Probe the tcgen05 GEMM core on a real B200 (development aid, not part of the product).

Stage 1 runs the self-test GEMM with the default shared-memory descriptors for every operand layout
(K-major / MN-major A and B, bf16 x bf16 and fp16 x bf16) and reports the error against torch fp32
matmul.  Stage 2, only for layouts that fail, sweeps the descriptor byte offsets (LBO / SBO / k-step)
to find the encoding the hardware expects.  Every stage runs in a subprocess so a faulting kernel
cannot take the probe down.  Results: gpurun_out/umma_probe.json
"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_cases(cases):
    import torch
    import b200face
    from b200face import _lib
    lib = b200face.load_library()
    dev = torch.device("cuda:0")
    out = []
    for c in cases:
        M, N, K = c["M"], c["N"], c["K"]
        g = torch.Generator(device=dev).manual_seed(1)
        a = torch.randn(M, K, generator=g, device=dev)
        b = torch.randn(N, K, generator=g, device=dev)
        fmt = c.get("fmt", 0)                               # 0 bf16 x bf16, 1 fp16 x bf16 (faults), 2 fp16 x fp16
        a16 = a.half() if fmt >= 1 else a.bfloat16()
        b16 = b.half() if fmt == 2 else b.bfloat16()
        ref = a16.float() @ b16.float().t()
        a_store = a16.t().contiguous() if c["a_mn"] else a16.contiguous()
        b_store = b16.t().contiguous() if c["b_mn"] else b16.contiguous()
        ks = c.get("k_splits", 1)
        res = torch.full((ks, M, N), float("nan"), device=dev)
        rc = lib.b200f_umma_selftest(_lib.ptr(a_store), _lib.ptr(b_store), _lib.ptr(res), M, N, K, c["a_mn"], c["b_mn"],
                                     fmt, ks, c.get("a_lbo", -1), c.get("a_sbo", -1),
                                     c.get("a_kstep", -1), c.get("b_lbo", -1), c.get("b_sbo", -1), c.get("b_kstep", -1),
                                     _lib.stream_ptr(dev))
        err = None
        if rc == 0:
            try:
                torch.cuda.synchronize()
                got = res.sum(0)
                err = float((got - ref).norm() / ref.norm())
            except RuntimeError as e:                       # sticky CUDA error: stop this subprocess
                out.append(dict(c, rc=rc, err=None, fault=str(e)[:200]))
                break
        flag = lib.b200f_umma_timeout_flag(1)
        out.append(dict(c, rc=rc, err=err, timeout=flag,
                        msg=(lib.b200f_last_error() or b"").decode() if rc else ""))
    return out


def sub(cases, timeout=240):
    p = subprocess.run([sys.executable, __file__, "--cases", json.dumps(cases)], capture_output=True, text=True,
                       timeout=timeout)
    for ln in p.stdout.splitlines():
        if ln.startswith("RESULT "):
            return json.loads(ln[7:])
    return [dict(c, rc=None, err=None, fault=(p.stderr or p.stdout)[-400:]) for c in cases]


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--cases":
        print("RESULT " + json.dumps(run_cases(json.loads(sys.argv[2]))), flush=True)
        return
    shapes = [dict(M=128, N=256, K=64), dict(M=128, N=256, K=256), dict(M=296, N=704, K=192), dict(M=512, N=1024, K=512)]
    layouts = [dict(a_mn=0, b_mn=0), dict(a_mn=0, b_mn=1), dict(a_mn=1, b_mn=1)]
    stage1 = [dict(s, **l) for l in layouts for s in shapes]
    stage1 += [dict(M=512, N=512, K=4096, a_mn=0, b_mn=1, k_splits=8)]
    report = {"stage1": sub(stage1)}
    # fp16 x fp16 (the format the head uses); fp16 x bf16 was probed once: illegal instruction on B200
    report["mixed_fp16"] = sub([dict(M=256, N=512, K=256, a_mn=0, b_mn=0, fmt=2),
                                dict(M=256, N=512, K=256, a_mn=0, b_mn=1, fmt=2),
                                dict(M=256, N=512, K=256, a_mn=1, b_mn=1, fmt=2)])
    ok = lambda r: r.get("err") is not None and r["err"] < 1e-2
    bad_layouts = []
    for l in layouts:
        rs = [r for r in report["stage1"] if r["a_mn"] == l["a_mn"] and r["b_mn"] == l["b_mn"]]
        if not all(ok(r) for r in rs):
            bad_layouts.append(l)
    report["bad_layouts"] = bad_layouts
    sweeps = {}
    base = dict(M=256, N=512, K=256)
    for l in bad_layouts:
        cands = []
        if l["b_mn"]:
            for lbo, sbo, ks in itertools.product([8192, 1024, 128, 16], [1024, 8192, 128], [2048, 32, 256]):
                c = dict(base, **l, b_lbo=lbo, b_sbo=sbo, b_kstep=ks)
                if l["a_mn"]:
                    c.update(a_lbo=lbo, a_sbo=sbo, a_kstep=ks)
                cands.append(c)
        else:
            for lbo, sbo, ks in itertools.product([0, 16, 1024], [1024, 128, 8192], [32, 16, 64]):
                cands.append(dict(base, **l, a_lbo=lbo, a_sbo=sbo, a_kstep=ks, b_lbo=lbo, b_sbo=sbo, b_kstep=ks))
        res = []
        for i in range(0, len(cands), 12):
            res += sub(cands[i:i + 12])
        sweeps[f"a_mn{l['a_mn']}_b_mn{l['b_mn']}"] = sorted(res, key=lambda r: (r.get("err") is None, r.get("err") or 9e9))[:6]
    report["sweeps"] = sweeps
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "umma_probe.json"), "w") as f:
        json.dump(report, f, indent=1)
    for k in ("stage1", "mixed_fp16"):
        for r in report[k]:
            print(k, {x: r[x] for x in r if x in ("M", "N", "K", "a_mn", "b_mn", "fmt", "k_splits", "rc", "err", "timeout", "fault", "msg")})
    print("bad layouts:", bad_layouts)
    for k, v in sweeps.items():
        print("sweep", k)
        for r in v:
            print("   ", {x: r[x] for x in r if x.endswith(("lbo", "sbo", "kstep")) or x in ("err", "fault")})


if __name__ == "__main__":
    main()
