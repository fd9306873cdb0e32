#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json on B200: fused ArcFace head (fwd+bwd) samples/s, with
the gallery-match queries/s beside it.  Prints ONE JSON line (rank 0).

    python bench.py --gpus 1 --steps K --warmup W                 # cfg3: 512-d x 100k classes, batch 512, bf16
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
                                                                  # cfg4: 1M classes class-sharded, batch 4096
    python bench.py --impl reference ...                          # CPU arm: the reference algorithm (torch port)

A "step" = one fused head forward+backward (K1 x2, K2, loss, hook scalar, K3, normalise-backward) on one
synthetic batch.  `value` times the steps with x / labels already in HBM; `e2e` times the same step
through the public API (ArcMarginProduct.forward_loss + backward) with x and labels coming from pinned
host memory every step and the loss read back every step.  Inputs (W bf16 102 MB + dW fp32 205 MB) exceed
the 126 MB L2, so no explicit flush is needed between iterations (config.l2 = "inputs_exceed_l2").
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG3 = dict(B=512, C=100_000, D=512)
CFG4 = dict(B=4096, C=1_000_000, D=512)
CFG2 = dict(Q=1000, N=10_000, D=512, k=1)
STREAM = dict(Q=128, N=1_000_000, D=512, k=5)
CFG5 = dict(Q=8192, N=1_000_000, D=512, k=5)
EPOCH = 10            # post-warm-up schedule state: m_eff = 0.45, s_eff = 6.72
LS = 0.05


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)       # ~0.3 s of GPU time: long enough for the power cap to bite
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every stage from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--pair", type=int, default=0, help="1 = single CTAs, 2 = cta_group::2 CTA pairs (default)")
    ap.add_argument("--g-chunk-mb", type=int, default=0, help="logit-gradient buffer budget per class chunk (MB)")
    ap.add_argument("--pdl", type=int, default=-1, help="0 / 1: programmatic dependent launch between the step's kernels (default on)")
    ap.add_argument("--tune", action="append", default=[], metavar="NAME=INT",
                    help="b200f_set_tunable(NAME, INT) before the run (experiments; see csrc/umma_head.cu)")
    ap.add_argument("--no-gallery", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the head + optimizer (K5) measurement")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the 1-GPU run of cfg4 (the scaling anchor) on the N = 1 line")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln.split(",") for ts, ln in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or \
               [ln.split(",") for _, ln in self.lines[-3:]]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def claim_stdout():
    """stdout must carry exactly ONE JSON line, and libraries print there too (with N > 1 NCCL writes
    "NCCL version ..." on file descriptor 1 whatever NCCL_DEBUG_FILE says).  Returns a file object on the original
    stdout and points descriptor 1 -- for this process, its C libraries and its children -- at stderr."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


W_BLOCK = 4096


def synth_w_rows(lo, hi, C_total, D, dev, seed=4321):
    """Rows [lo, hi) of THE class-weight matrix of the run: xavier_normal(gain sqrt 2) of the full [C_total, D] matrix
    (SURVEY 8d), bf16, drawn block by block from generators seeded with (seed, block) -- every world size sees the same
    matrix (losses of the N = 1, 2, 4, 8 lines are comparable), any rank can materialise any row."""
    std = (2.0 ** 0.5) * (2.0 / (C_total + D)) ** 0.5
    out = torch.empty(hi - lo, D, dtype=torch.bfloat16, device=dev)
    b = lo // W_BLOCK
    while b * W_BLOCK < hi:
        r0, r1 = b * W_BLOCK, min((b + 1) * W_BLOCK, C_total)
        g = torch.Generator(device=dev).manual_seed(seed * 1_000_003 + b)
        blk = (torch.randn(r1 - r0, D, generator=g, device=dev) * std).to(torch.bfloat16)
        s0, s1 = max(r0, lo), min(r1, hi)
        out[s0 - lo:s1 - lo] = blk[s0 - r0:s1 - r0]
        b += 1
    return out


def synth_head(B, C_total, D, dev, lo, hi):
    """SURVEY 8d recipe: x ~ N(0,1) -> bf16, W xavier_normal(gain sqrt2) -> bf16 compute copy, y ~ U{0..C-1}; 12.5 % of
    the rows planted near their class centre (cos ~ 0.95) so that the margin is exercised.  x and y are the same on every
    rank; w = rows [lo, hi) of the global matrix."""
    g = torch.Generator(device=dev).manual_seed(1234)
    std = (2.0 ** 0.5) * (2.0 / (C_total + D)) ** 0.5
    x = torch.randn(B, D, generator=g, device=dev)
    y = torch.randint(0, C_total, (B,), generator=g, device=dev)
    n = B // 8
    centres = torch.cat([synth_w_rows(int(c), int(c) + 1, C_total, D, dev) for c in y[:n].tolist()]).float()
    x[:n] = 3.0 * centres + 1.0 * std * torch.randn(n, D, generator=g, device=dev)
    return x.to(torch.bfloat16), synth_w_rows(lo, hi, C_total, D, dev), y


def rel_err(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / b.norm())


def engine_code(name):
    from b200face import _lib
    return {"auto": _lib.ENGINE_AUTO, "simt": _lib.ENGINE_SIMT, "tcgen05": _lib.ENGINE_TCGEN05}[name]


def run_b200(args):
    import b200face
    from b200face import _lib, parallel
    from b200face import head as H
    from b200face.head import GraphedHeadStep, HeadStats, arcface_loss, effective_margin_scale, head_schedule
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    result_out = sys.stdout
    if world > 1:
        result_out = claim_stdout()                           # NCCL prints its version banner on stdout
        torch.distributed.init_process_group("nccl", device_id=dev)
        group = torch.distributed.group.WORLD
    lib = b200face.load_library()
    if args.pair:
        lib.b200f_set_tunable(b"pair", args.pair)
    if args.g_chunk_mb:
        lib.b200f_set_tunable(b"g_chunk_mb", args.g_chunk_mb)
    if args.pdl in (0, 1):
        lib.b200f_set_tunable(b"pdl", args.pdl)
    for kv in args.tune:
        name, _, val = kv.partition("=")
        lib.b200f_set_tunable(name.encode(), int(val))
    cfgw = CFG3 if world == 1 else CFG4
    B, C_total, D = cfgw["B"], cfgw["C"], cfgw["D"]
    c_lo, c_hi = parallel.shard_bounds(C_total, world, rank)
    C_local = c_hi - c_lo
    mf, sf = head_schedule(EPOCH, 10, True, True, 0.0, 0.3)
    m_eff, s_eff = effective_margin_scale(32.0, 0.5, mf, sf, True)
    eng = engine_code(args.engine)
    use_graph = not args.eager

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- N > 1 first: (1) sharded-vs-unsharded parity on a reduced class count, (2) cfg4 UNSHARDED on rank 0 = the
    # 1-GPU anchor of the scaling line, measured in this job (same box, same clocks)
    parity, anchor = None, None
    if world > 1:
        parity = parity_sharded(dev, world, rank, group, eng, not args.no_cpu_baseline)
        sync_all()
        if rank == 0 and not args.no_cfg4:
            anchor = bench_cfg4_single_gpu(dev, eng, m_eff, s_eff, steps=5)
        sync_all()
    x, w, y = synth_head(B, C_total, D, dev, c_lo, c_hi)
    w_master = w.float().requires_grad_(True)        # fp32 master receives the fp32 dW; kernels read the bf16 copy
    loss_kw = dict(compute_weight=w, m_eff=m_eff, s_eff=s_eff, label_smoothing=LS, class_offset=c_lo,
                   num_classes_total=C_total, group=group, engine=eng)

    def eager_step(xin, yin):
        xin = xin.detach().requires_grad_(True)
        w_master.grad = None
        loss = arcface_loss(xin, w_master, yin, **loss_kw)
        loss.backward()
        return loss

    launches_per_step = None
    if use_graph:
        l0 = lib.b200f_launch_count()
        gstep = GraphedHeadStep(w_master, B, D, dtype=torch.bfloat16, warmup=1, **loss_kw)
        launches_per_step = (lib.b200f_launch_count() - l0) // 2          # one warm-up step + the captured step
        gstep(x, y)
        step = lambda: gstep.replay()
    else:
        step = lambda: eager_step(x, y)
    for _ in range(max(args.warmup, 3)):
        step()
    # ---- per-kernel durations: each C-ABI stage alone between CUDA events on the launching stream, L2 flushed
    # (a 256 MB write) before every launch; these are what the roofline line reports.  Taken BEFORE the long timed
    # region, i.e. at the clocks of a kernel timed alone (the roofline uses the burst peaks).
    sync_all()
    kern = time_stages(H, _lib, x, w, y, m_eff, s_eff, c_lo, C_total, eng, dev, reps=max(5, min(args.steps, 20)))
    # the gallery lines are per-call figures too: measured here, before the long region heats the board
    gal_single = bench_gallery(dev, peaks(), eng) if (world == 1 and not args.no_gallery and rank == 0) else None
    # ---- burst figure: 20 steps from an idle GPU (full clocks); the timed region below is long enough for the
    # board power cap to pull the SM clock down (sw_power_cap: 1965 -> ~1670 MHz after 1 s of this step)
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None      # started here: nvidia-smi needs ~0.3 s before its first line
    time.sleep(0.5)
    eb0, eb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eb0.record()
    for _ in range(20):
        step()
    eb1.record()
    sync_all()
    ms_burst = torch.tensor([eb0.elapsed_time(eb1) / 20], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms_burst, op=torch.distributed.ReduceOp.MAX)
    ms_burst = float(ms_burst)
    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    _lib.TIMERS.clear(); _lib.PROFILE = not use_graph
    sync_all()
    l0 = lib.b200f_launch_count()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    sync_all()
    t_wall1 = time.time()
    launches = (launches_per_step * args.steps) if use_graph else (lib.b200f_launch_count() - l0)
    _lib.PROFILE = False
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(ms)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    loss_val = float(loss.detach())
    # ---- timed region 2: end to end through the public API, host buffers ------------------------
    # the public modules: ArcMarginProduct on one GPU, parallel.ShardedArcMarginProduct (class-parallel) on several
    xh = x.cpu().pin_memory(); yh = y.cpu().pin_memory()
    if use_graph:
        gstep.close()                                                 # frees the first graph's buffers (and its NCCL nodes)
        del gstep
    if world == 1:
        head = b200face.ArcMarginProduct(D, C_local).to(dev)
    else:
        head = parallel.ShardedArcMarginProduct(D, C_total, group=group).to(dev)
    head.update_epoch(EPOCH); head.train()
    hl = head if world == 1 else head.local
    hl.engine = eng
    hl.compute_dtype = torch.bfloat16
    hl.cache_weight_prep = False                                      # training changes W every step: K1(W) is timed
    with torch.no_grad():
        hl.weight.copy_(w.float())
    api_step = head.graphed_step(B, LS, torch.bfloat16) if use_graph else None
    def e2e_step():
        if use_graph:
            l = api_step(xh, yh)                                      # H2D into the static buffers, then replay
        else:
            xd = xh.to(dev, non_blocking=True); yd = yh.to(dev, non_blocking=True)
            hl.zero_grad(set_to_none=True)
            l = head.forward_loss(xd.requires_grad_(True), yd, LS)
            l.backward()
        return l
    # The loss of EVERY step is read back to the host, through pinned memory with a one-step lag: the D2H copy of
    # step i is enqueued behind it and read while step i+1 runs, so the host never idles the GPU (a training loop
    # that logs its loss does exactly this).
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]
    losses_read = []
    def e2e_loop(n):
        for i in range(n):
            l = e2e_step()
            loss_host[i & 1].copy_(l.detach(), non_blocking=True)
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                losses_read.append(float(loss_host[(i - 1) & 1]))
        loss_ev[(n - 1) & 1].synchronize()
        losses_read.append(float(loss_host[(n - 1) & 1]))
    e2e_loop(3)
    sync_all()
    losses_read.clear()
    e0.record()
    e2e_loop(args.steps)
    e1.record()
    sync_all()
    assert len(losses_read) == args.steps
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms2, op=torch.distributed.ReduceOp.MAX)
    e2e_val = B * args.steps / (float(ms2) * 1e-3)
    e2e_loss = losses_read[-1]
    if use_graph:
        api_step.close()
    gal_sharded = None
    if world > 1 and not args.no_gallery:
        gal_sharded = bench_gallery_sharded(dev, world, rank, group, eng, sync_all, peaks())

    if rank != 0:
        finish(world)
        return
    pk = peaks()
    value = B * args.steps / (ms_total * 1e-3)
    flops_step = 6.0 * B * C_total * D
    # Roofline of the DOMINANT kernel of the step (the longest of the four GEMM kernels, each timed alone between
    # CUDA events the library records on the launching stream, L2 flushed before the stage).  Algorithmic work per
    # launch: K2 / K3a / K3c 2*B*C*D FLOP; K3b is bounded by bytes: it reads G^T (2 B per
    # batch row and class, rows padded to 64) and w_hat16 (2*D B per class) and writes dW fp32 (4*D B per class).
    n_cls = C_local                                           # event pairs are summed over the class chunks of a call
    roofs = {}
    gemm_flop = 2.0 * B * C_local * D
    for k, label in (("k2", "K2 arcface_fwd: cosine GEMM + margin + softmax-CE statistics"),
                     ("k3a", "K3a logit gradient: cosine GEMM recompute + G^T (fp16) + column sums"),
                     ("k3c", "K3c dX = G . w_hat (split-K)"),
                     ("k3b", "K3b dW_hat = G^T . x_hat (batch > 512: both operands streamed, normalise-backward separate)")):
        if k == "k3b" and B <= 512:
            continue                                           # fused, byte-bound form: below
        if kern.get(k) and n_cls:
            ach = gemm_flop / (kern[k] * 1e-3) / 1e12
            roofs[k] = {"kernel": label, "bound": "tensor", "achieved": round(ach, 2), "peak": pk["tf_burst"], "unit": "TFLOP/s",
                        "frac": round(ach / pk["tf_burst"], 4), "traffic": load_traffic(k),
                        "algorithmic_flop_per_launch": gemm_flop, "avg_launch_ms": round(kern[k], 4)}
    if kern.get("k3b") and n_cls and B <= 512:
        ldg = (B + 63) // 64 * 64
        k3b_bytes = C_local * (2.0 * ldg + 2.0 * D + 4.0 * D) + 2.0 * B * D
        ach = k3b_bytes / (kern["k3b"] * 1e-3) / 1e9
        roofs["k3b"] = {"kernel": "K3b dW = G^T . x_hat with the normalise-backward of W fused (reads G^T + w_hat16, writes dW fp32)",
                        "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s",
                        "frac": round(ach / pk["hbm"], 4), "traffic": load_traffic("k3b"),
                        "algorithmic_bytes_per_launch": k3b_bytes, "avg_launch_ms": round(kern["k3b"], 4)}
    roof = None
    if roofs:
        dom = max(roofs, key=lambda k: roofs[k]["avg_launch_ms"])
        roof = dict(roofs[dom])
        roof["peak_source"] = pk["source"] + (", burst figure (kernel timed alone)" if roof["bound"] == "tensor" else ", measured copy bandwidth")
        roof["share_of_step"] = round(roof["avg_launch_ms"] / ms_burst, 3)      # both at full clocks
    elif kern.get("arcface_fwd"):                              # CUDA-core engine: the fused forward stage as a whole
        ach = gemm_flop / (kern["arcface_fwd"] * 1e-3) / 1e12
        roof = {"kernel": "arcface_fwd stage", "bound": "tensor", "achieved": round(ach, 2), "peak": pk["tf_burst"],
                "unit": "TFLOP/s", "frac": round(ach / pk["tf_burst"], 4), "traffic": None,
                "avg_launch_ms": round(kern["arcface_fwd"], 4)}
    out = {
        "metric": "arcface_head_samples_per_sec", "value": round(value, 1), "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world, args, m_eff, s_eff),
        "algorithmic_tflops": round(flops_step * args.steps / (ms_total * 1e-3) / 1e12, 2),
        "frac_of_bf16_peak": round(flops_step * args.steps / (ms_total * 1e-3) / 1e12 / (pk["tf_sustained"] * world), 4),
        "e2e": {"value": round(e2e_val, 1), "unit": "samples/s", "h2d_bytes_per_step": int(xh.numel() * 2 + yh.numel() * 8),
                "d2h_bytes_per_step": 4, "d2h": "loss of every step, pinned buffer, read with a one-step lag"},
        "burst": {"steps": 20, "ms_per_step": round(ms_burst, 4), "value": round(B / (ms_burst * 1e-3), 1),
                  "note": "20 steps from an idle GPU at full SM clock; `value` above is the sustained figure of the timed region"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "roofline_all": {k: {kk: v[kk] for kk in ("bound", "achieved", "peak", "unit", "frac", "avg_launch_ms")} for k, v in roofs.items()},
        "kernel_ms": {k: round(v, 4) for k, v in kern.items()},
        "loss": round(loss_val, 5), "e2e_loss": round(e2e_loss, 5),
    }
    if anchor is not None:
        # the driver's efficiency formula divides by the N = 1 LINE, which is cfg3 (another workload): this is cfg4 against
        # cfg4 on one GPU, measured by rank 0 of this very job before the sharded run
        out["anchor_1gpu_ms"] = anchor["ms_per_step"]
        out["anchor_1gpu"] = anchor
        out["efficiency_vs_cfg4_1gpu"] = round(anchor["ms_per_step"] / (world * ms_burst), 4)
        out["efficiency_vs_cfg4_1gpu_sustained"] = round(anchor["ms_per_step"] / (world * ms_total / args.steps), 4)
    if parity is not None:
        out["parity"] = parity
    if world == 1 and not args.no_train_step and use_graph and H.use_tcgen05(x, eng):
        out["train_step"] = bench_train_step(dev, pk, eng, w, x, y)
    if world == 1 and not args.no_cfg4 and use_graph and H.use_tcgen05(x, eng):
        torch.cuda.empty_cache()
        out["cfg4_single_gpu"] = bench_cfg4_single_gpu(dev, eng, m_eff, s_eff)
    out["scaling_note"] = ("N = 1 runs cfg3 (100k classes, batch 512) and N >= 2 run cfg4 (1M classes class-sharded, batch 4096), "
                           "as BASELINE.json's configs name them; cfg4's own 1-GPU anchor is `cfg4_single_gpu` on the N = 1 line")
    if gal_single is not None:
        out["gallery"] = gal_single
    if world > 1 and gal_sharded is not None:
        out["gallery"] = gal_sharded
    if world == 1 and not args.no_cpu_baseline:
        # CPU arm on the SAME inputs as the GPU arm: its first step is also the parity reference of this line
        gpu_out = public_api_step(hl, x, y)
        out["cpu_baseline"], cpu_out = cpu_head_baseline(CFG3, x.cpu(), w.cpu(), y.cpu())
        out["parity"] = {"what": "this line's GPU step (ArcMarginProduct.forward_loss + backward, tcgen05 engine) against the CPU "
                                 "port of the reference (fp32) on the identical bf16-rounded x, W, y",
                         "loss_rel": abs(gpu_out[0] - cpu_out[0]) / abs(cpu_out[0]),
                         "dx_rel": rel_err(gpu_out[1], cpu_out[1]), "dw_rel": rel_err(gpu_out[2], cpu_out[2]),
                         "loss_gpu": gpu_out[0], "loss_cpu": cpu_out[0], "tolerance": 1e-3,
                         "graph_replay_loss": round(loss_val, 6)}
        del gpu_out, cpu_out
        if not args.no_gallery:
            out.setdefault("gallery", {})["cpu_baseline"] = cpu_gallery_baselines()
        out["cfg1_cpu"] = cpu_cfg1_baseline()
    print(json.dumps(out), file=result_out, flush=True)
    finish(world)


def finish(world):
    """Every captured graph has been closed by now (GraphedHeadStep.close releases the NCCL kernels it holds), so the
    process group can be torn down normally."""
    if world > 1:
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def public_api_step(head, x, y):
    """One eager step through the public module on device tensors -> (loss, dx fp32, dW) on the CPU."""
    head.zero_grad(set_to_none=True)
    xg = x.detach().clone().requires_grad_(True)
    loss = head.forward_loss(xg, y, LS)
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), head.last_stats.dx_f32.cpu(), head.weight.grad.cpu()


def parity_sharded(dev, world, rank, group, eng, with_cpu):
    """N > 1: the class-parallel head across the REAL NCCL ranks against (a) the unsharded head on rank 0's GPU and (b) the
    CPU port of the reference, on a reduced class count (B = 4096 x C = 8192 per rank) so that both references finish
    in seconds.  Every rank takes part; rank 0 returns the record."""
    import b200face
    from b200face import parallel
    B, D = CFG4["B"], CFG4["D"]
    C = 8192 * world
    lo, hi = parallel.shard_bounds(C, world, rank)
    x, w, y = synth_head(B, C, D, dev, lo, hi)
    sh = parallel.ShardedArcMarginProduct(D, C, group=group).to(dev)
    sh.update_epoch(EPOCH); sh.train(); sh.local.engine = eng; sh.local.compute_dtype = torch.bfloat16
    with torch.no_grad():
        sh.local.weight.copy_(w.float())
    l_s, dx_s, dw_s = public_api_step(sh.local, x, y)
    full_w = sh.gather_weight()                                # the reference's [C, D] layout (all ranks take part)
    rec = None
    if rank == 0:
        full = b200face.ArcMarginProduct(D, C).to(dev)
        full.update_epoch(EPOCH); full.train(); full.engine = eng; full.compute_dtype = torch.bfloat16
        full.load_state_dict({"weight": full_w, "u": torch.zeros(1)}, strict=True)
        l_u, dx_u, dw_u = public_api_step(full, x, y)
        rec = {"what": f"class-parallel head over {world} NCCL ranks vs the unsharded head on one GPU vs the CPU port; "
                       f"B = {B}, C = {C} ({hi - lo} per rank), D = {D}, same seeded bf16 inputs",
               "sharded_vs_unsharded": {"loss_rel": abs(l_s - l_u) / abs(l_u), "dx_rel": rel_err(dx_s, dx_u),
                                        "dw_rel_rank0_rows": rel_err(dw_s, dw_u[lo:hi])},
               "loss_sharded": l_s, "loss_unsharded": l_u, "tolerance": 1e-3}
        if with_cpu:
            _, cpu = cpu_head_baseline(dict(B=B, C=C, D=D), x.cpu(), full_w.cpu().bfloat16(), y.cpu(), budget_s=0.0)
            rec["sharded_vs_cpu_port"] = {"loss_rel": abs(l_s - cpu[0]) / abs(cpu[0]), "dx_rel": rel_err(dx_s, cpu[1]),
                                          "dw_rel_rank0_rows": rel_err(dw_s, cpu[2][lo:hi])}
            rec["loss_cpu"] = cpu[0]
        del full
    del sh, full_w
    torch.cuda.empty_cache()
    return rec


def time_stages(H, _lib, x, w, y, m_eff, s_eff, c_lo, C_total, eng, dev, reps=10):
    """Average duration of each stage of the step launched alone (CUDA events on the launching stream, a 256 MB
    L2 flush before every launch): K1 on the weights, K2 (+ its partial reduction), K3 (all backward kernels)."""
    lib = _lib.load_library()
    cfg = H._head_cfg(m_eff, s_eff, LS, False, C_total, eng)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    B = x.shape[0]
    acc = {"l2norm_rows_w": [], "arcface_fwd": [], "arcface_bwd": [], "k2": [], "k3a": [], "k3b": [], "k3c": []}
    have_events = H.use_tcgen05(x, eng)                       # the tensor engine records per-kernel event pairs on request
    ms_c = ctypes.c_float()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(reps + 2):
        f16n = H.use_tcgen05(x, eng)
        xo, inv_nx = H._k1(x, f16n)
        flush.zero_()
        a0, a1 = ev(), ev(); a0.record(); wo, inv_nw = H._k1(w, f16n); a1.record()
        flush.zero_()
        _lib.PROFILE = True; _lib.TIMERS.clear()
        if have_events:
            lib.b200f_set_tunable(b"stage_events", 1)
        out = H._fwd_kernels(x, w, y, cfg, c_lo, False)
        row_stats = out[4]
        lse = torch.empty(B, dtype=torch.float32, device=dev); out2 = torch.empty(2, dtype=torch.float32, device=dev)
        _lib.check(lib.b200f_arcface_loss(_lib.ptr(row_stats), B, cfg, _lib.ptr(lse), _lib.ptr(out2), _lib.ptr(out2[1:]),
                                          _lib.stream_ptr(dev)), "loss")
        out4 = torch.empty(4, dtype=torch.float32, device=dev)
        _lib.check(lib.b200f_arcface_hook_scale(_lib.ptr(out2[1:]), None, B, s_eff, 0, 1.0, 1, 0, _lib.ptr(out4),
                                                _lib.stream_ptr(dev)), "hook")
        flush.zero_()
        H._bwd_kernels(out[0], out[1], y, out[2], out[3], lse, out4, cfg, c_lo)
        _lib.PROFILE = False
        torch.cuda.synchronize()
        if have_events:
            lib.b200f_set_tunable(b"stage_events", 0)
            for k in ("k2", "k3a", "k3b", "k3c"):
                _lib.check(lib.b200f_stage_ms(k.encode(), ctypes.byref(ms_c)), "stage_ms")
                acc[k].append(float(ms_c.value))
        acc["l2norm_rows_w"].append(a0.elapsed_time(a1))
        for k in ("arcface_fwd", "arcface_bwd"):
            acc[k].append(statistics.mean(a.elapsed_time(b) for a, b in _lib.TIMERS[k]))
    return {k: statistics.mean(v[2:]) for k, v in acc.items() if v}


def load_traffic(kernel):
    """dram bytes per launch of the named kernel from the committed ncu capture (profiles/), else null."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        rec = json.load(open(p)).get(kernel)
        if isinstance(rec, dict):
            return rec.get("dram_bytes")
        return rec
    return None


def bench_cfg4_single_gpu(dev, eng, m_eff, s_eff, steps=20):
    """cfg4 (1 M classes, batch 4096) on ONE GPU: the 1-GPU anchor of the class-parallel scaling line (the N = 1 bench
    line itself is cfg3, as BASELINE.json's configs prescribe).  `steps` graph replays after 3, device-resident inputs."""
    from b200face.head import GraphedHeadStep
    c = CFG4
    x, w, y = synth_head(c["B"], c["C"], c["D"], dev, 0, c["C"])
    w_master = w.float().requires_grad_(True)
    g = GraphedHeadStep(w_master, c["B"], c["D"], dtype=torch.bfloat16, warmup=1, compute_weight=w, m_eff=m_eff, s_eff=s_eff,
                        label_smoothing=LS, class_offset=0, num_classes_total=c["C"], group=None, engine=eng)
    g(x, y)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = steps
    e0.record()
    for _ in range(n):
        loss = g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    out = {"workload": "cfg4 on 1 GPU: ArcFace head 512-d, 1M classes (unsharded), batch 4096, bf16 fwd+bwd",
           "ms_per_step": round(ms, 4), "value": round(c["B"] / (ms * 1e-3), 1), "unit": "samples/s", "steps": n,
           "algorithmic_tflops": round(6.0 * c["B"] * c["C"] * c["D"] / (ms * 1e-3) / 1e12, 1), "loss": round(float(loss), 5)}
    g.close()
    del g, w_master, w, x
    torch.cuda.empty_cache()
    return out


def bench_train_step(dev, pk, eng, w_bf16, x, y):
    """SURVEY 8f rank 3 -- the optimizer step right behind the head.  One training step of the head's weights =
    fwd + bwd (CUDA-graph replay) + AdamW(amsgrad) on the fp32 master [C, D]:
      unfused: graph with K1 over W inside + torch.optim.AdamW(fused=True) (the reference's optimizer, torch's best kernel)
      fused  : graph WITHOUT K1 over W + b200f_head_adamw, which also writes next step's normalised fp16 operands.
    Optimizer kernels are also timed alone (L2 flushed); bytes per launch = C*D*(9*4 + 2) + C*4."""
    import b200face
    B, D = x.shape
    C = w_bf16.shape[0]

    def make_head():
        h = b200face.ArcMarginProduct(D, C).to(dev)
        h.update_epoch(EPOCH); h.train(); h.engine = eng; h.compute_dtype = torch.bfloat16
        with torch.no_grad():
            h.weight.copy_(w_bf16.float())
        return h
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    res = {}

    def timed_loop(body, n=60):
        for _ in range(5):
            body()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            body()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def timed_alone(fn, n=10):
        ts = []
        for _ in range(n + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.mean(ts[2:])
    # ---- unfused
    h = make_head(); h.cache_weight_prep = False
    opt_t = torch.optim.AdamW([h.weight], lr=1e-3, weight_decay=1e-4, amsgrad=True, fused=True)
    step = h.graphed_step(B, LS, torch.bfloat16)
    def unfused():
        step(x, y); opt_t.step()
    res["unfused_ms"] = round(timed_loop(unfused), 4)
    res["torch_fused_adamw_ms"] = round(timed_alone(opt_t.step), 4)
    del step, opt_t, h
    torch.cuda.empty_cache()
    # ---- fused
    h = make_head()
    opt_f = b200face.HeadAdamW(h, lr=1e-3, weight_decay=1e-4, amsgrad=True)
    step = h.graphed_step(B, LS, torch.bfloat16, optimizer=opt_f)
    def fused():
        step(x, y); opt_f.step()
    res["fused_ms"] = round(timed_loop(fused), 4)
    k5 = timed_alone(opt_f.step)
    res["b200f_head_adamw_ms"] = round(k5, 4)
    byt = C * D * (9 * 4 + 2) + C * 4
    res["k5_algorithmic_bytes"] = byt
    res["k5_GBps"] = round(byt / (k5 * 1e-3) / 1e9, 1)
    res["k5_frac_of_hbm_peak"] = round(byt / (k5 * 1e-3) / 1e9 / pk["hbm"], 4)
    res["samples_per_sec_fused"] = round(B / (res["fused_ms"] * 1e-3), 1)
    res["samples_per_sec_unfused"] = round(B / (res["unfused_ms"] * 1e-3), 1)
    res["note"] = ("head fwd+bwd (graph replay) + AdamW(amsgrad) of the fp32 class weights, inputs resident; "
                   "unfused = K1(W) in the graph + torch.optim.AdamW(fused=True)")
    return res


def bench_gallery(dev, pk, eng):
    """Gallery match beside the head: cfg2 (1k x 10k, top-1 + threshold; latency-bound, fits L2) and the streaming
    regime (Q=128 vs 1M x 512: HBM-bound).  The gallery is resident with its scan operand prepared once
    (b200face.PreparedGallery, what GalleryIndex keeps): the tensor engine then streams the bf16 rows (2 B / element),
    re-ranks its candidates exactly in fp32 and proves the top-k per query; `redo` counts queries sent to the exact
    engine.  Roofline of the streaming run: bytes actually streamed per pass / time against the HBM peak."""
    import b200face
    g = torch.Generator(device=dev).manual_seed(1234)
    res = {}
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for name, c in (("cfg2", CFG2), ("stream_q128_n1m", STREAM)):
        G = torch.nn.functional.normalize(torch.randn(c["N"], c["D"], generator=g, device=dev), dim=1)
        Q = torch.nn.functional.normalize(torch.randn(c["Q"], c["D"], generator=g, device=dev), dim=1)
        h = c["Q"] // 2
        src = torch.randint(0, c["N"], (h,), generator=g, device=dev)
        tau = 0.5 + 2.0 * torch.rand(h, 1, generator=g, device=dev)
        Q[:h] = torch.nn.functional.normalize(G[src] + tau / c["D"] ** 0.5 * torch.randn(h, c["D"], generator=g, device=dev), dim=1)
        tc = b200face.gallery.tensor_engine_ok(Q, G, eng)
        prep = b200face.PreparedGallery(G, "l2eps") if tc else None
        redo = torch.zeros(1, dtype=torch.int32, device=dev)
        call = lambda qq: b200face.gallery_topk(qq, G, c["k"], 1.0, "l2eps", engine=eng, prepared=prep, redo_count=redo)
        for _ in range(3):
            call(Q)
        torch.cuda.synchronize()
        reps = 10
        times = []
        for _ in range(reps):
            if name != "cfg2":
                flush.zero_()                                   # the 1 GB operand exceeds L2 anyway; keep the rule explicit
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx, score, acc = call(Q); e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.mean(times)
        redo_per_call = int(redo) / (reps + 3)
        Qh = Q.cpu().pin_memory()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            i2, s2, a2 = call(Qh.to(dev, non_blocking=True))
            host = (i2.cpu(), s2.cpu(), a2.cpu())
        e1.record(); torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1) / reps
        eg = 2 if tc else 4
        byt = c["N"] * c["D"] * eg + c["N"] * 4 + c["Q"] * c["D"] * 4 + c["Q"] * c["k"] * 12
        res[name] = {"queries_per_sec": round(c["Q"] / (ms * 1e-3), 1), "ms": round(ms, 4),
                     "e2e_queries_per_sec": round(c["Q"] / (ms_e2e * 1e-3), 1),
                     "engine": "tcgen05 bf16 scan + exact fp32 re-rank + proof" if tc else "fp32 CUDA cores (exact)",
                     "redo_per_call": redo_per_call,
                     "accepted_frac": round(float(acc.float().mean()), 3),
                     "streamed_bytes": byt, "gallery_elem_bytes": eg,
                     "algorithmic_GBps": round(byt / (ms * 1e-3) / 1e9, 1),
                     "frac_of_hbm_peak": round(byt / (ms * 1e-3) / 1e9 / pk["hbm"], 4),
                     "fp32_equivalent_GBps": round((c["N"] * c["D"] * 4) / (ms * 1e-3) / 1e9, 1),
                     "algorithmic_tflops": round(2.0 * c["Q"] * c["N"] * c["D"] / (ms * 1e-3) / 1e12, 2)}
        if name != "cfg2" and tc:
            # several batches in flight (b200face.gallery_topk_batches, private streams): the latency-bound head and
            # tail of one call overlap the HBM-bound main scan of another.  Every batch streams the 1 GB operand
            # again (8x the L2), so no flush between them.
            depth, nb = b200face.gallery.PIPELINE_DEPTH, 24
            perms = [Q[torch.randperm(c["Q"], generator=g, device=dev)] for _ in range(6)]
            batches = [perms[i % 6] for i in range(nb)]
            pcall = lambda: b200face.gallery_topk_batches(batches, G, c["k"], 1.0, "l2eps", depth=depth, engine=eng, prepared=prep,
                                                          redo_count=redo)
            pcall()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); outs = pcall(); e1.record()
            torch.cuda.synchronize()
            ms_b = e0.elapsed_time(e1) / nb
            res[name]["pipelined"] = {"depth": depth, "batches": nb, "ms_per_batch": round(ms_b, 4),
                                      "queries_per_sec": round(c["Q"] / (ms_b * 1e-3), 1),
                                      "algorithmic_GBps": round(byt / (ms_b * 1e-3) / 1e9, 1),
                                      "frac_of_hbm_peak": round(byt / (ms_b * 1e-3) / 1e9 / pk["hbm"], 4),
                                      "note": "same kernels and results as the serial calls; up to `depth` batches in flight"}
            del outs, batches, perms
        del G, Q, prep
    return res


def bench_gallery_sharded(dev, world, rank, group, eng, sync_all, pk):
    """cfg5: 1M x 512 gallery rows split contiguously over the ranks, 8192 queries on every rank, top-5:
    per-shard tensor-engine top-k (exact), ONE all-gather of [Q,5] x 12 B per rank, merge (lowest global index wins
    ties).  Device time, max over ranks."""
    import b200face
    from b200face import parallel
    c = CFG5
    lo, hi = parallel.shard_bounds(c["N"], world, rank)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    G = torch.nn.functional.normalize(torch.randn(hi - lo, c["D"], generator=g, device=dev), dim=1)
    gq = torch.Generator(device=dev).manual_seed(1234)
    Q = torch.nn.functional.normalize(torch.randn(c["Q"], c["D"], generator=gq, device=dev), dim=1)
    prep = b200face.PreparedGallery(G, "l2eps")
    local = lambda q, gs, k, thr, metric, index_offset=0: b200face.gallery_topk(q, gs, k, thr, metric, index_offset=index_offset,
                                                                          engine=eng, prepared=prep)[:2]
    call = lambda: parallel.sharded_gallery_topk(Q, G, c["k"], 1.0, "l2eps", index_offset=lo, group=group, local_topk=local,
                                                 merge=b200face.gallery.merge_topk)
    for _ in range(3):
        call()
    sync_all()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        idx, score, acc = call()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(ms)
    tf = 2.0 * c["Q"] * c["N"] * c["D"] / (ms * 1e-3) / 1e12
    return {"cfg5": {"queries_per_sec": round(c["Q"] / (ms * 1e-3), 1), "ms": round(ms, 4), "Q": c["Q"], "N_total": c["N"],
                     "N_per_rank": hi - lo, "k": c["k"],
                     "algorithmic_tflops": round(tf, 2),
                     "roofline": {"bound": "tensor", "achieved": round(tf, 2), "peak": pk["tf_burst"] * world, "unit": "TFLOP/s",
                                  "frac": round(tf / (pk["tf_burst"] * world), 4),
                                  "note": "Q = 8192 is tensor-bound (intensity 2Q/2 B = 8192 FLOP/B): 2 Q N D against the "
                                          "burst bf16 peak of the N GPUs; the HBM-bound regime is stream_q128_n1m on the N = 1 line"},
                     "comm": "one all-gather of [Q,k] (score fp32, index int64) per rank + merge kernel"}}


def cpu_head_baseline(c, x, w, y, budget_s=20.0):
    """The reference algorithm on this box's host cores (oracle/torch_port.py, kind 'port': the reference is pure Python
    and is not on the GPU box), on the SAME x / W / y as the GPU arm.  Bounded: full steps until ~budget_s.
    Returns (cpu_baseline record, (loss, dx, dW) of the first step: the parity reference)."""
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    head = torch_port.HeadPort(c["D"], c["C"])
    head.current_epoch = EPOCH
    head.train()
    with torch.no_grad():
        head.weight.copy_(w.float())
    xf = x.float()
    loss, dx, dw = torch_port.head_step(head, xf, y, LS)     # warm-up = the parity reference
    first = (float(loss), dx.clone(), dw.clone())
    t0 = time.time(); n = 0
    while budget_s > 0:
        torch_port.head_step(head, xf, y, LS); n += 1
        if time.time() - t0 > budget_s or n >= 10:
            break
    if n == 0:
        return None, first
    dt = (time.time() - t0) / n
    rec = {"value": round(c["B"] / dt, 1), "unit": "samples/s", "cores": threads, "kind": "port",
           "sample": f"{n} full cfg3 head steps (B={c['B']}, C={c['C']}, fp32 torch CPU, fwd+bwd, the GPU arm's inputs), {dt:.2f} s each"}
    return rec, first


def cpu_gallery_baselines():
    """SURVEY 8d / BASELINE.md 4.2: (i) the reference's own compare_faces loop (src/app.py:50-64: one F.pairwise_distance
    + .item() per reference, a Python loop, 1 thread by construction) and (ii) the vectorised stand-in the north star
    names, torch.cdist + topk on all cores -- for cfg2 and for ONE rank's shard of cfg5 (125 k rows; the whole of cfg5 is
    8 such shards).  Bounded samples, stated."""
    from oracle import torch_port
    threads = os.cpu_count() or 1
    g = torch.Generator().manual_seed(1234)
    out = {"cores": threads, "kind": "port"}
    for name, N, Qv, nq_loop, k in (("cfg2", CFG2["N"], CFG2["Q"], 8, 1), ("cfg5_one_shard_of_8", CFG5["N"] // 8, 2048, 2, 5)):
        G = torch.nn.functional.normalize(torch.randn(N, 512, generator=g), dim=1)
        Q = torch.nn.functional.normalize(torch.randn(Qv, 512, generator=g), dim=1)
        refs = [{"name": str(i), "embedding": G[i:i + 1]} for i in range(N)]
        torch.set_num_threads(1)
        t0 = time.time()
        for i in range(nq_loop):
            torch_port.compare_faces_loop(Q[i:i + 1], refs, 1.0)
        t_loop = (time.time() - t0) / nq_loop
        torch.set_num_threads(threads)
        torch_port.gallery_vectorised(Q[:64], G, k, 1.0)
        t0 = time.time()
        torch_port.gallery_vectorised(Q, G, k, 1.0)
        t_vec = time.time() - t0
        out[name] = {"compare_faces_loop_queries_per_sec": round(1.0 / t_loop, 2), "loop_threads": 1,
                     "loop_sample": f"{nq_loop} queries x {N} references, verbatim Python loop",
                     "cdist_topk_queries_per_sec": round(Qv / t_vec, 1), "cdist_threads": threads,
                     "cdist_sample": f"one call, {Qv} queries x {N} rows, top-{k}"}
        del refs, G, Q
    out["cfg5_note"] = "cfg5 = 8 shards of 125 k rows: whole-gallery CPU throughput is 1/8 of the per-shard figures (extrapolated)"
    return out


def cpu_cfg1_baseline(steps=4):
    """cfg1, the reference's own CPU-runnable case: ResNet18 + ArcFace head, 36 classes, 512-d, batch 32, 224 x 224, one
    full training step (forward, CE with label smoothing 0.05, backward, AdamW-amsgrad) -- oracle/torch_port.ArcFaceNetPort
    (random-init trunk: the ImageNet weights cannot be downloaded; same FLOPs) -- and the head alone at that shape."""
    from oracle import torch_port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    m = torch_port.ArcFaceNetPort(36).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=1e-4, amsgrad=True)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(32, 3, 224, 224, generator=g); y = torch.randint(0, 36, (32,), generator=g)
    torch_port.arcfacenet_train_step(m, opt, x, y)
    t0 = time.time()
    for _ in range(steps):
        torch_port.arcfacenet_train_step(m, opt, x, y)
    dt = (time.time() - t0) / steps
    head = torch_port.HeadPort(512, 36).train()
    emb = torch.randn(32, 512, generator=g)
    torch_port.head_step(head, emb, y, LS)
    t0 = time.time()
    for _ in range(50):
        torch_port.head_step(head, emb, y, LS)
    dth = (time.time() - t0) / 50
    return {"workload": "cfg1: ResNet18 + ArcFace head, 36 classes, 512-d, batch 32, 224x224, full train step on CPU (port)",
            "samples_per_sec": round(32 / dt, 2), "ms_per_step": round(dt * 1e3, 1), "steps": steps, "cores": threads, "kind": "port",
            "head_only_ms": round(dth * 1e3, 3), "head_only_samples_per_sec": round(32 / dth, 1),
            "head_share_of_step": round(dth / dt, 5)}


def workload_config(world, args, m_eff, s_eff):
    """The `config` object of the JSON line: the workload BASELINE.json names, identical in both arms (the reference arm
    times the reference's CPU path on THIS configuration; what it samples of it is stated in its cpu_baseline.sample)."""
    cfgw = CFG3 if world == 1 else CFG4
    c_local = cfgw["C"] // world + (1 if cfgw["C"] % world else 0)
    return {"workload": ("cfg3: ArcFace head 512-d, 100k classes, batch 512 bf16 fwd+bwd, 1 B200" if world == 1 else
                         f"cfg4: ArcFace partial-FC 512-d, 1M classes class-sharded over {world} B200 "
                         f"({c_local} per rank), batch 4096, one all-reduce each way"),
            "B": cfgw["B"], "C_total": cfgw["C"], "C_per_rank": c_local, "D": cfgw["D"], "label_smoothing": LS,
            "m_eff": round(m_eff, 4), "s_eff": round(s_eff, 4), "engine": args.engine,
            "launch": "cuda-graph replay of the captured step" if not args.eager else "eager (one C-ABI call per stage)",
            "grads": "dx fp32, dW fp32", "l2": "inputs_exceed_l2 (W bf16 + dW fp32 = 3 x C x D x 2 B per rank)",
            "parallelism": "single GPU" if world == 1 else f"class-parallel x{world} (NCCL all-reduce [B,4] fwd, [B,D] bwd)"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (torch port of
    src/face_models.py:334-429 + CrossEntropyLoss + autograd), all host threads.
    N = 1: cfg3 as is.  N > 1: the b200 arm runs cfg4 (1 M classes, batch 4096); the CPU arm times ONE of its 8 class
    shards at the FULL batch (B = 4096 x C = 125 000: like-for-like per row and per class) and reports
    samples/s of the whole of cfg4 as 4096 / (8 x shard time) -- an extrapolation, labelled as such."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import torch_port
    from b200face.head import effective_margin_scale, head_schedule      # pure Python (the schedule); loads no library
    mf, sf = head_schedule(EPOCH, 10, True, True, 0.0, 0.3)
    m_eff, s_eff = effective_margin_scale(32.0, 0.5, mf, sf, True)
    multi = args.gpus > 1
    c = dict(CFG3) if not multi else dict(CFG4, C=CFG4["C"] // 8)
    shards = 8 if multi else 1
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    head = torch_port.HeadPort(c["D"], c["C"])
    head.current_epoch = EPOCH
    head.train()
    x = torch.randn(c["B"], c["D"], generator=g)
    y = torch.randint(0, c["C"], (c["B"],), generator=g)
    # bounded: the whole run must end within a few minutes whatever K is asked
    t0 = time.time(); torch_port.head_step(head, x, y, LS); probe = time.time() - t0
    steps = max(1, min(args.steps, int(120.0 / max(probe, 1e-3))))
    warm = max(0, min(args.warmup, int(30.0 / max(probe, 1e-3))))
    for _ in range(warm):
        torch_port.head_step(head, x, y, LS)
    t0 = time.time()
    for _ in range(steps):
        loss, _, _ = torch_port.head_step(head, x, y, LS)
    dt = (time.time() - t0) / steps
    val = c["B"] / (dt * shards)
    sample = (f"{steps} full cfg3 head steps (B={c['B']}, C={c['C']}, fp32, fwd+bwd) of the reference algorithm on CPU"
              if not multi else
              f"{steps} steps of ONE of cfg4's 8 class shards at the full batch (B={c['B']}, C={c['C']}, fp32, fwd+bwd), {dt:.2f} s each; "
              f"value = {c['B']} / (8 x that): the whole of cfg4 extrapolated from one shard")
    print(json.dumps({
        "impl": "reference", "metric": "arcface_head_samples_per_sec", "value": round(val, 1), "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm + 1, "ms_per_step": round(dt * shards * 1e3, 2),
        "higher_is_better": True, "scaling": "strong" if multi else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args, m_eff, s_eff),
        "timed": {"B": c["B"], "C_timed": c["C"], "shards_extrapolated": shards},
        "cpu_baseline": {"value": round(val, 1), "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 1), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": round(float(loss), 5)}), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
