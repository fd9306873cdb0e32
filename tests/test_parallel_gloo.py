"""World-size-2 gloo tests (CPU) of the multi-GPU HOST logic in b200face.parallel: the collectives,
their layouts and the merge rules.  The per-shard kernel outputs are produced by the oracle here (the
CUDA kernels need a GPU; their shard arithmetic is covered by tests/test_gpu_head.py::test_class_shards*),
so what is under test is exactly the code path between the kernels on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, cfg_from_golden, golden
import oracle
from oracle import arcface_oracle as ao


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, fn(rank, world)))
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def _head_job(rank, world):
    from b200face import parallel
    d = golden("head_postwarm.npz")
    cfg = cfg_from_golden(d)
    x, w, y = d["x"].astype(np.float64), d["w"].astype(np.float64), d["y"]
    B, C = x.shape[0], w.shape[0]
    lo, hi = parallel.shard_bounds(C, world, rank)
    m_eff, s_eff = oracle.effective_margin_scale(cfg)
    owned = (y >= lo) & (y < hi)
    loc = np.where(owned, y - lo, 0)
    z = ao._shard_logits(x, w[lo:hi], loc, owned, cfg, np.float64)           # stand-in for K2 on this shard
    rows = np.arange(B)
    stats = np.stack([np.exp(z - s_eff).sum(1), np.exp(2 * (z - s_eff)).sum(1),
                      np.where(owned, z[rows, loc], 0.0), z.sum(1)], axis=1)
    t = torch.tensor(stats)
    parallel.reduce_row_stats(t)                                              # product code under test
    st = t.numpy()
    eps = cfg.label_smoothing
    lse = s_eff + np.log(st[:, 0])
    loss = float((lse - (1 - eps) * st[:, 2] - eps / C * st[:, 3]).mean())
    best = torch.tensor(z.max(1)); arg = torch.tensor(z.argmax(1) + lo)
    top, pred = parallel.merge_row_argmax(best, arg)
    # backward exchange: partial dx_hat of this shard
    full = oracle.head_forward_backward(x, w, y, cfg)
    wh, _ = oracle.l2_normalize_rows(w)
    part = torch.tensor(full["g_cos"][:, lo:hi] @ wh[lo:hi])
    parallel.reduce_dxhat(part)
    cmm = parallel.reduce_cos_minmax(torch.tensor([float(z.min()), float(z.max())]))
    return dict(loss=loss, lse=lse, pred=pred.numpy(), top=top.numpy(), dxhat=part.numpy(), cmm=cmm.numpy())


def test_class_sharded_head_exchange():
    out = _run(_head_job, 2)
    d = golden("head_postwarm.npz")
    cfg = cfg_from_golden(d)
    full = oracle.head_forward_backward(d["x"], d["w"], d["y"], cfg)
    wh, _ = oracle.l2_normalize_rows(d["w"].astype(np.float64))
    for r in (0, 1):
        assert out[r]["loss"] == pytest.approx(float(full["loss"]), rel=1e-12)
        np.testing.assert_allclose(out[r]["lse"], full["lse"], rtol=1e-12)
        assert np.array_equal(out[r]["pred"], full["argmax"])
        np.testing.assert_allclose(out[r]["top"], full["logits"].max(1), rtol=1e-12)
        np.testing.assert_allclose(out[r]["dxhat"], full["g_cos"] @ wh, rtol=1e-9, atol=1e-15)
        assert out[r]["cmm"][0] == pytest.approx(full["logits"].min()) and out[r]["cmm"][1] == pytest.approx(full["logits"].max())


def _gallery_job(rank, world):
    from b200face import parallel
    d = golden("gallery_synth.npz")
    g, q = d["g"], d["q"]
    lo, hi = parallel.shard_bounds(g.shape[0], world, rank)
    res = {}
    for metric, largest in (("l2eps", False), ("cos", True)):
        def local(q_, g_, k_, t_, m_, off):                   # stand-in for K4 on this shard
            i, s, _ = oracle.gallery_topk(q_.numpy(), g_.numpy(), k_, t_, m_)
            return torch.tensor(np.where(i >= 0, i + off, -1)), torch.tensor(s)

        def merge(idx_all, score_all, t_, m_):                # stand-in for the merge kernel
            i, s = oracle.merge_topk_shards(list(idx_all.numpy()), list(score_all.numpy()), idx_all.shape[2], largest)
            acc = (s[:, 0] >= t_) if largest else (s[:, 0] <= t_)
            return torch.tensor(i), torch.tensor(s), torch.tensor(acc)

        idx, score, acc = parallel.sharded_gallery_topk(torch.tensor(q), torch.tensor(g[lo:hi]), 5, 1.0, metric,
                                                        index_offset=lo, local_topk=local, merge=merge)
        res[metric] = (idx.numpy(), score.numpy(), acc.numpy())
    return res


def test_gallery_sharded_allgather_merge():
    out = _run(_gallery_job, 2)
    d = golden("gallery_synth.npz")
    for metric in ("l2eps", "cos"):
        gi, gs, ga = oracle.gallery_topk(d["q"], d["g"], 5, 1.0, metric)
        for r in (0, 1):
            idx, score, acc = out[r][metric]
            assert np.array_equal(idx, gi)                    # incl. the planted duplicate rows 17/150
            np.testing.assert_allclose(score, gs, rtol=1e-6)
            assert np.array_equal(acc, ga)


def test_full_matrix_init_is_world_size_independent():
    """ShardedArcMarginProduct draws its rows from ONE full-matrix xavier_normal(gain sqrt 2) distribution
    (src/face_models.py:324): std uses the TOTAL class count, shards of any partition tile the same matrix."""
    from b200face import parallel
    C, D = 10_001, 64
    full = parallel.full_matrix_init_rows(0, C, C, D, seed=3)
    assert float(full.std()) == pytest.approx((2.0 ** 0.5) * (2.0 / (C + D)) ** 0.5, rel=0.02)
    for world in (2, 3, 8):
        parts = [parallel.full_matrix_init_rows(*parallel.shard_bounds(C, world, r), C, D, seed=3) for r in range(world)]
        assert torch.equal(torch.cat(parts), full)
    assert not torch.equal(full[:4096], full[4096:8192])              # blocks are not copies of each other
    assert not torch.equal(full, parallel.full_matrix_init_rows(0, C, C, D, seed=4))
