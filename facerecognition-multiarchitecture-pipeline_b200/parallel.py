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


class ShardedArcMarginProduct(torch.nn.Module):
    """Class-parallel ArcMarginProduct: rank r owns weight rows [lo_r, hi_r) of the [C,D] matrix.
    Every rank sees the full batch of embeddings and labels (global class ids).  forward_loss runs
    K1-K3 on the shard with one all-reduce each way.  state_dict interop with the reference layout:
    gather_weight() / load_full_weight()."""

    def __init__(self, in_feats, out_feats, s=32.0, m=0.5, use_warm_up=True, easy_margin=False, group=None):
        super().__init__()
        from .head import ArcMarginProduct
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.out_feats_total = out_feats
        self.lo, self.hi = shard_bounds(out_feats, self.world, self.rank)
        self.local = ArcMarginProduct(in_feats, self.hi - self.lo, s=s, m=m, use_warm_up=use_warm_up,
                                      easy_margin=easy_margin)

    def update_epoch(self, epoch):
        self.local.update_epoch(epoch)

    def forward_loss(self, input, label, label_smoothing=0.05, return_pred=False):
        from .head import arcface_loss, HeadStats
        hd = self.local
        m_eff, s_eff = hd._step_schedule()
        x, weight, w = hd._operands(input)
        hd.last_stats = HeadStats()
        loss = arcface_loss(x, weight, label, compute_weight=w, m_eff=m_eff, s_eff=s_eff, label_smoothing=label_smoothing,
                            easy_margin=hd.easy_margin, class_offset=self.lo,
                            num_classes_total=self.out_feats_total, group=self.group, hook=hd._hook,
                            stats=hd.last_stats, engine=hd.engine)
        if return_pred:
            _, pred = merge_row_argmax(hd.last_stats.row_best, hd.last_stats.row_argmax, self.group)
            return loss, pred
        return loss

    def gather_weight(self) -> torch.Tensor:
        """Full [C,D] weight in the reference's layout (checkpoint save)."""
        parts = [torch.empty(shard_bounds(self.out_feats_total, self.world, r)[1]
                             - shard_bounds(self.out_feats_total, self.world, r)[0],
                             self.local.in_feats, dtype=self.local.weight.dtype,
                             device=self.local.weight.device) for r in range(self.world)]
        dist.all_gather(parts, self.local.weight.detach().contiguous(), group=self.group)
        return torch.cat(parts, dim=0)

    def load_full_weight(self, weight: torch.Tensor):
        """Take this rank's rows of a reference-layout [C,D] weight (checkpoint load)."""
        with torch.no_grad():
            self.local.weight.copy_(weight[self.lo:self.hi].to(self.local.weight.device))
