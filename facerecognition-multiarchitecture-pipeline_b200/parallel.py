"""Multi-GPU sharding of the two parts of the hot path (one process per GPU, torch.distributed).

The reference is single-process (no distributed code anywhere); this is the B200 scale-out named by
BASELINE.json: classes partitioned partial-FC style with ONE all-reduce of the per-row softmax
statistics forward and ONE all-reduce of dx_hat backward; gallery rows partitioned with a per-shard
top-k, ONE all-gather and a merge.  Collectives go through torch.distributed (NCCL over
NVLink/NVSwitch on the GPU box; gloo in the CPU tests of this host logic).  Messages are tiny
([B,4] fp32, [B,D] fp32, [Q,k] x 12 B) so they are latency-bound; nothing here is per-link sized.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split [lo, hi) of `total` rows over `world` ranks."""
    return (total * rank) // world, (total * (rank + 1)) // world


def reduce_row_stats(row_stats: torch.Tensor, group=None) -> torch.Tensor:
    """SUM the [B,4] forward statistics over class shards.  Every column is additive because the
    softmax shift is the constant s_eff (|logit| <= s_eff), so no MAX exchange is needed."""
    dist.all_reduce(row_stats, op=dist.ReduceOp.SUM, group=group)
    return row_stats


def reduce_dxhat(dxhat: torch.Tensor, group=None) -> torch.Tensor:
    """SUM the per-shard partial dx_hat = G_shard . w_hat_shard  ([B,D] fp32)."""
    dist.all_reduce(dxhat, op=dist.ReduceOp.SUM, group=group)
    return dxhat


def merge_row_argmax(row_best: torch.Tensor, row_argmax: torch.Tensor, group=None):
    """Global outputs.max(1): gather every shard's (best logit, global class id), keep the largest
    logit, lowest class id on ties (shards own ascending class ranges, so the first maximum wins)."""
    world = dist.get_world_size(group)
    bests = [torch.empty_like(row_best) for _ in range(world)]
    idxs = [torch.empty_like(row_argmax) for _ in range(world)]
    dist.all_gather(bests, row_best, group=group)
    dist.all_gather(idxs, row_argmax, group=group)
    best = torch.stack(bests)                       # [P,B]
    idx = torch.stack(idxs)
    valid = idx >= 0
    key = torch.where(valid, best, torch.full_like(best, float("-inf")))
    top = key.max(dim=0).values
    is_top = (key == top.unsqueeze(0)) & valid
    big = torch.iinfo(torch.int64).max
    arg = torch.where(is_top, idx, torch.full_like(idx, big)).min(dim=0).values
    return top, arg


def reduce_cos_minmax(cos_minmax: torch.Tensor, group=None) -> torch.Tensor:
    """Global {min,max} cosine for get_margin_stats (only when somebody asks for it)."""
    lo = cos_minmax[0:1].clone()
    hi = cos_minmax[1:2].clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return torch.cat([lo, hi])


def gather_topk(idx: torch.Tensor, score: torch.Tensor, group=None):
    """all-gather the per-shard [Q,k] lists -> [P,Q,k] (global ids)."""
    world = dist.get_world_size(group)
    idx_all = [torch.empty_like(idx) for _ in range(world)]
    score_all = [torch.empty_like(score) for _ in range(world)]
    dist.all_gather(idx_all, idx.contiguous(), group=group)
    dist.all_gather(score_all, score.contiguous(), group=group)
    return torch.stack(idx_all), torch.stack(score_all)


def sharded_gallery_topk(q: torch.Tensor, g_shard: torch.Tensor, k: int, thresh: float, metric: str = "l2eps",
                         *, index_offset: int, group=None,
                         local_topk: Optional[Callable] = None, merge: Optional[Callable] = None):
    """Gallery-parallel match: every rank holds all queries and rows [index_offset, +N_local) of the
    gallery.  local_topk / merge default to the CUDA kernels (K4 and its merge); the CPU tests of this
    host logic inject stand-ins."""
    if local_topk is None or merge is None:
        from . import gallery
        local_topk = local_topk or (lambda q_, g_, k_, t_, m_, off: gallery.gallery_topk(
            q_, g_, k_, t_, m_, index_offset=off)[:2])
        merge = merge or gallery.merge_topk
    idx, score = local_topk(q, g_shard, k, thresh, metric, index_offset)
    idx_all, score_all = gather_topk(idx, score, group)
    return merge(idx_all, score_all, thresh, metric)


INIT_BLOCK_ROWS = 4096


def full_matrix_init_rows(lo: int, hi: int, out_feats: int, in_feats: int, seed: int = 0) -> torch.Tensor:
    """Rows [lo, hi) of a [out_feats, in_feats] matrix initialised like the reference's FULL head
    (``nn.init.xavier_normal_(weight, gain=sqrt 2)``, src/face_models.py:324: std = sqrt(2) * sqrt(2 / (C + D)) with
    the TOTAL class count as fan-out), drawn block by block from generators seeded with (seed, block): the matrix is
    the same whatever the number of ranks, no rank draws rows it does not own, and no two ranks hold duplicate rows."""
    std = (2.0 ** 0.5) * (2.0 / (out_feats + in_feats)) ** 0.5
    out = torch.empty(hi - lo, in_feats, dtype=torch.float32)
    b = lo // INIT_BLOCK_ROWS
    while b * INIT_BLOCK_ROWS < hi:
        r0, r1 = b * INIT_BLOCK_ROWS, min((b + 1) * INIT_BLOCK_ROWS, out_feats)
        g = torch.Generator().manual_seed(seed * 1_000_003 + b)
        blk = torch.randn(r1 - r0, in_feats, generator=g) * std
        s0, s1 = max(r0, lo), min(r1, hi)
        out[s0 - lo:s1 - lo] = blk[s0 - r0:s1 - r0]
        b += 1
    return out


class ShardedArcMarginProduct(torch.nn.Module):
    """Class-parallel ArcMarginProduct: rank r owns weight rows [lo_r, hi_r) of the [C,D] matrix.
    Every rank sees the full batch of embeddings and labels (global class ids).  ``forward_loss`` / ``graphed_step``
    run K1-K3 on the shard with one all-reduce each way (``self.local`` is an ordinary ArcMarginProduct placed at
    class offset lo_r).  Interop with the reference's checkpoint layout (``arcface.weight`` [C,512], strict load at
    src/testing.py:124): ``gather_weight()`` / ``load_full_weight()`` / ``full_state_dict()``.
    Call ``close()`` before ``torch.distributed.destroy_process_group()`` when graphed steps were used."""

    def __init__(self, in_feats, out_feats, s=32.0, m=0.5, use_warm_up=True, easy_margin=False, group=None, seed=0):
        super().__init__()
        from .head import ArcMarginProduct
        group = group if group is not None else dist.group.WORLD       # the head reads "no group" as "not sharded"
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.out_feats_total = out_feats
        self.lo, self.hi = shard_bounds(out_feats, self.world, self.rank)
        self.local = ArcMarginProduct(in_feats, self.hi - self.lo, s=s, m=m, use_warm_up=use_warm_up,
                                      easy_margin=easy_margin)
        with torch.no_grad():                      # the shard of the FULL matrix's init, not an init of a small matrix
            self.local.weight.copy_(full_matrix_init_rows(self.lo, self.hi, out_feats, in_feats, seed))
        self.local._class_offset = self.lo
        self.local._num_classes_total = out_feats
        self.local._group = group

    def update_epoch(self, epoch):
        self.local.update_epoch(epoch)

    @property
    def weight(self):
        return self.local.weight

    def forward_loss(self, input, label, label_smoothing=0.05, return_pred=False):
        return self.local.forward_loss(input, label, label_smoothing, return_pred)

    def graphed_step(self, B, label_smoothing=0.05, dtype=torch.bfloat16, optimizer=None):
        """CUDA-graph form of forward_loss + backward for a fixed batch size, both all-reduces captured inside."""
        return self.local.graphed_step(B, label_smoothing, dtype, optimizer)

    def close(self):
        self.local.release_graphs()

    def gather_weight(self) -> torch.Tensor:
        """Full [C,D] weight in the reference's layout (checkpoint save)."""
        parts = [torch.empty(shard_bounds(self.out_feats_total, self.world, r)[1]
                             - shard_bounds(self.out_feats_total, self.world, r)[0],
                             self.local.in_feats, dtype=self.local.weight.dtype,
                             device=self.local.weight.device) for r in range(self.world)]
        dist.all_gather(parts, self.local.weight.detach().contiguous(), group=self.group)
        return torch.cat(parts, dim=0)

    def full_state_dict(self) -> dict:
        """state_dict of the equivalent unsharded reference module: {'weight': [C,D], 'u': [1]}."""
        return {"weight": self.gather_weight(), "u": self.local.u.detach().clone()}

    def load_full_weight(self, weight: torch.Tensor):
        """Take this rank's rows of a reference-layout [C,D] weight (checkpoint load)."""
        if tuple(weight.shape) != (self.out_feats_total, self.local.in_feats):
            raise ValueError(f"expected a [{self.out_feats_total}, {self.local.in_feats}] weight, got {tuple(weight.shape)}")
        with torch.no_grad():
            self.local.weight.copy_(weight[self.lo:self.hi].to(self.local.weight.device))
        self.local._drop_weight_caches()
