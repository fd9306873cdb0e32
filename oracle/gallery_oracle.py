"""numpy restatement of the reference gallery match (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/app.py:50-64 (``compare_faces``: Euclidean
``F.pairwise_distance`` with its default eps=1e-6 added element-wise to the
difference, strict ``<`` so the first index wins ties, accept iff d_min <= thresh)
and the cosine class-centre siblings
/root/reference/src/hyperparameter_tuning.py:1036-1047,1076 and
/root/reference/src/face_models.py:891-893
(``normalize(emb) @ normalize(W).T [* s]`` then ``max(1)``).

Nothing here is imported by the product.  Pinned against the reference's own
``compare_faces`` on its ``face_references.pkl`` fixture by
tests/golden/make_golden.py -> tests/golden/gallery_*.npz.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

PAIRWISE_EPS = 1e-6      # torch.nn.functional.pairwise_distance default
COS_NORM_EPS = 1e-12     # F.normalize default


def pairwise_distance_eps(q: np.ndarray, g: np.ndarray, dtype=np.float32) -> np.ndarray:
    """d[i,j] = || q_i - g_j + 1e-6 ||_2   (app.py:59).  q [Q,D], g [N,D] -> [Q,N]."""
    q = np.asarray(q, dtype=dtype)
    g = np.asarray(g, dtype=dtype)
    eps = np.asarray(PAIRWISE_EPS, dtype=dtype)
    out = np.empty((q.shape[0], g.shape[0]), dtype=dtype)
    for i in range(q.shape[0]):
        diff = (q[i][None, :] - g) + eps
        out[i] = np.sqrt(np.sum(diff * diff, axis=1, dtype=dtype))
    return out


def compare_faces(emb: Optional[np.ndarray], refs: Sequence[dict], thresh: float):
    """Verbatim control flow of app.py:50-64 on numpy arrays ([1,D] embeddings)."""
    if emb is None or not refs:
        return "Unknown", float("inf"), None
    min_dist = float("inf")
    best_match = "Unknown"
    best_ref_idx = None
    for i, ref in enumerate(refs):
        dist = float(pairwise_distance_eps(np.asarray(emb).reshape(1, -1),
                                           np.asarray(ref["embedding"]).reshape(1, -1))[0, 0])
        if dist < min_dist:
            min_dist = dist
            best_match = ref["name"]
            best_ref_idx = i
    return (best_match, min_dist, best_ref_idx) if min_dist <= thresh else ("Unknown", min_dist, None)


def _stable_topk(score: np.ndarray, k: int, largest: bool) -> Tuple[np.ndarray, np.ndarray]:
    """Row-wise top-k with lowest-index tie-break."""
    key = -score if largest else score
    idx = np.argsort(key, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int64), np.take_along_axis(score, idx, axis=1)


def gallery_topk(q: np.ndarray, g: np.ndarray, k: int, thresh: float, metric: str = "l2eps",
                 dtype=np.float32):
    """Batched form of compare_faces: per query the k best gallery rows.
    metric 'l2eps': ascending eps-distance, accept iff best <= thresh (app.py:60,64).
    metric 'cos'  : descending cosine of the row-normalised vectors, accept iff best >= thresh.
    Returns (idx [Q,k] int64, score [Q,k], accept [Q] bool); missing slots (N<k): idx -1,
    score +inf / -inf."""
    q = np.asarray(q, dtype=dtype)
    g = np.asarray(g, dtype=dtype)
    Q, N = q.shape[0], g.shape[0]
    if metric == "l2eps":
        s = pairwise_distance_eps(q, g, dtype)
        largest, fill = False, np.inf
    elif metric == "cos":
        s = cosine_scores(q, g, dtype)
        largest, fill = True, -np.inf
    else:
        raise ValueError(metric)
    kk = min(k, N)
    idx = np.full((Q, k), -1, dtype=np.int64)
    sc = np.full((Q, k), fill, dtype=dtype)
    if kk > 0:
        idx[:, :kk], sc[:, :kk] = _stable_topk(s, kk, largest)
    if N == 0:
        accept = np.zeros(Q, dtype=bool)
    else:
        accept = (sc[:, 0] >= thresh) if largest else (sc[:, 0] <= thresh)
    return idx, sc, accept


def cosine_scores(emb: np.ndarray, w: np.ndarray, dtype=np.float32) -> np.ndarray:
    emb = np.asarray(emb, dtype=dtype)
    w = np.asarray(w, dtype=dtype)
    en = emb / np.maximum(np.sqrt((emb * emb).sum(1, keepdims=True)), np.asarray(COS_NORM_EPS, dtype))
    wn = w / np.maximum(np.sqrt((w * w).sum(1, keepdims=True)), np.asarray(COS_NORM_EPS, dtype))
    return en @ wn.T


def cosine_class_match(emb: np.ndarray, w: np.ndarray, s: float = 1.0, dtype=np.float32):
    """hyperparameter_tuning.py:1039-1046,1076: logits = normalize(emb) @ normalize(W).T * s;
    pred = logits.max(1).indices (first max wins).  Returns (pred [B], best_logit [B])."""
    logits = cosine_scores(emb, w, dtype) * np.asarray(s, dtype=dtype)
    pred = logits.argmax(axis=1)
    return pred.astype(np.int64), logits[np.arange(logits.shape[0]), pred]


def merge_topk_shards(idx_list: List[np.ndarray], score_list: List[np.ndarray], k: int,
                      largest: bool):
    """Merge per-shard top-k lists (global indices) into the global top-k, lowest global
    index first on ties (all-gather then k-way merge)."""
    idx = np.concatenate(idx_list, axis=1)
    sc = np.concatenate(score_list, axis=1)
    bad = idx < 0
    key = np.where(bad, np.inf, -sc if largest else sc)
    order = np.lexsort((np.where(bad, np.iinfo(np.int64).max, idx), key), axis=1)[:, :k]
    return np.take_along_axis(idx, order, axis=1), np.take_along_axis(sc, order, axis=1)
