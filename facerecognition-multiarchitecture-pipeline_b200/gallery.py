"""Host side of the gallery match over libb200face.so (kernel K4).

Reference surface mirrored here:
  compare_faces(emb, refs, thresh) -> (name, dist, idx)      /root/reference/src/app.py:50-64
  cosine class-centre match (normalize @ normalize.T * s, max(1))
      /root/reference/src/hyperparameter_tuning.py:1039-1046,1076 ; src/face_models.py:891-893
Batched entry: gallery_topk(Q, G, k, thresh, metric).  GalleryIndex keeps the gallery resident on the
GPU so the Streamlit loop (src/app.py:631-639) does not re-upload it every frame.

No CPU path: the kernels run on CUDA or the call raises."""
from __future__ import annotations

import collections
import weakref
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, dtype_code, ptr, require_cuda, stream_ptr

_METRICS = {"l2eps": _lib.METRIC_L2EPS, "cos": _lib.METRIC_COS}
MAX_K = 16


def _default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("b200face gallery match needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _inv_norm(t: torch.Tensor) -> torch.Tensor:
    lib = _lib.load_library()
    inv = torch.empty(t.shape[0], dtype=torch.float32, device=t.device)
    if t.shape[0]:
        check(lib.b200f_l2norm_rows(ptr(t), dtype_code(t), t.shape[0], t.shape[1], 1e-12, ptr(inv), None, 0, 1.0,
                                    stream_ptr(t.device)), "b200f_l2norm_rows")
    return inv


TC_MIN_WORK = 1 << 22        # Q*N below this stays on the exact CUDA-core engine (launch-bound anyway)


class PreparedGallery:
    """Scan operand of the tensor engine for one gallery and one metric (b200f_gallery_prepare): the bf16 rows
    (L2-normalised for 'cos') and the per-row bias.  Build it once per gallery version; every query batch then
    streams 2 bytes per gallery element instead of 4."""

    def __init__(self, g: torch.Tensor, metric: str, operand_fmt: Optional[int] = None):
        """operand_fmt: _lib.OPERAND_FP16 (8x tighter proof margin, values must stay far inside +-65504) or
        OPERAND_BF16 (any range).  None: fp16 when max|g| <= 1024 -- embeddings and class centres -- else bf16
        (one host read of the maximum, at build time only)."""
        require_cuda(g)
        lib = _lib.load_library()
        self.metric = metric
        self.N, self.D = g.shape
        self.key = (g.data_ptr(), g._version, tuple(g.shape), metric)
        if operand_fmt is None:
            small = metric == "cos" or (self.N > 0 and float(g.abs().max()) <= 1024.0)
            operand_fmt = _lib.OPERAND_FP16 if small else _lib.OPERAND_BF16
        self.operand_fmt = operand_fmt
        self.g16 = torch.empty(self.N, self.D, dtype=torch.float16 if operand_fmt == _lib.OPERAND_FP16 else torch.bfloat16,
                               device=g.device)
        self.bias = torch.empty(self.N + 2, dtype=torch.float32, device=g.device)
        check(lib.b200f_gallery_prepare(ptr(g), dtype_code(g), self.N, self.D, _METRICS[metric], operand_fmt,
                                        ptr(self.g16), ptr(self.bias), stream_ptr(g.device)), "b200f_gallery_prepare")

    def matches(self, g: torch.Tensor, metric: str) -> bool:
        return self.key == (g.data_ptr(), g._version, tuple(g.shape), metric)


# Callers that pass the same gallery TENSOR again and again without keeping a PreparedGallery (compare_faces in a loop,
# gallery_topk(q, g)) would pay a full pass over g plus a host read of its maximum per call.  The last few prepared
# operands are therefore remembered per tensor OBJECT (a weak reference: the entry dies with the tensor, and a new tensor
# that happens to land on the same address never matches) and per version counter.  As with every version-keyed cache,
# writes through `.data` are invisible to it -- GalleryIndex, which owns its rows, is the interface for galleries that change.
_PREPARED_KEEP = 2
_prepared_cache: "collections.OrderedDict[int, tuple]" = collections.OrderedDict()


def _prepared_for(g: torch.Tensor, metric: str) -> PreparedGallery:
    ent = _prepared_cache.get(id(g))
    if ent is not None:
        ref, prep = ent
        if ref() is g and prep.matches(g, metric):
            _prepared_cache.move_to_end(id(g))
            return prep
        del _prepared_cache[id(g)]
    prep = PreparedGallery(g, metric)
    key = id(g)
    _prepared_cache[key] = (weakref.ref(g, lambda _r, k=key: _prepared_cache.pop(k, None)), prep)
    while len(_prepared_cache) > _PREPARED_KEEP:
        _prepared_cache.popitem(last=False)
    return prep


def tensor_engine_ok(q: torch.Tensor, g: torch.Tensor, engine: int) -> bool:
    """The bf16 tcgen05 scan + exact re-rank takes fp32 inputs with D % 8 == 0, D <= 512 on sm_100; AUTO uses it
    once the scan is big enough to matter."""
    if engine == _lib.ENGINE_SIMT or q.dtype != torch.float32 or g.shape[0] == 0:
        return False
    if not _lib.load_library().b200f_gallery_has_tc(q.shape[1]):
        if engine == _lib.ENGINE_TCGEN05:
            raise RuntimeError("the tensor gallery engine needs sm_100, D % 8 == 0 and D <= 512")
        return False
    return engine == _lib.ENGINE_TCGEN05 or q.shape[0] * g.shape[0] >= TC_MIN_WORK


def gallery_topk(q: torch.Tensor, g: torch.Tensor, k: int = 1, thresh: float = 1.0, metric: str = "l2eps",
                 *, index_offset: int = 0, g_inv: Optional[torch.Tensor] = None,
                 engine: int = _lib.ENGINE_AUTO, prepared: Optional[PreparedGallery] = None,
                 redo_count: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per query the k best gallery rows.  q [Q,D], g [N,D] CUDA, same dtype (fp32 / bf16).
    metric 'l2eps': score = ||q - g + 1e-6||_2 ascending, accept = best <= thresh  (app.py:59-64)
    metric 'cos'  : score = cosine of the row-normalised vectors, descending, accept = best >= thresh
    Returns (idx [Q,k] int64 global row ids, -1 = no such neighbour; score [Q,k] fp32; accept [Q] bool).
    Ties go to the lowest index (strict '<', app.py:60).
    engine: AUTO = tensor engine for big fp32 scans (results identical to the exact engine), SIMT = exact fp32
    CUDA-core engine, TCGEN05 = force the tensor engine.  prepared: a PreparedGallery of g to reuse."""
    if metric not in _METRICS:
        raise ValueError(f"metric must be one of {list(_METRICS)}")
    if not 1 <= k <= MAX_K:
        raise ValueError(f"k must be in [1, {MAX_K}]")
    require_cuda(q, g)
    if q.dtype != g.dtype:
        raise TypeError("queries and gallery must share a dtype")
    q = q.contiguous()
    g = g.contiguous()
    Q, D = q.shape
    N = g.shape[0]
    if N and g.shape[1] != D:
        raise ValueError("dimension mismatch between queries and gallery")
    lib = _lib.load_library()
    dev = q.device
    idx = torch.empty(Q, k, dtype=torch.int64, device=dev)
    score = torch.empty(Q, k, dtype=torch.float32, device=dev)
    accept = torch.empty(Q, dtype=torch.uint8, device=dev)
    q_inv = gi = None
    if metric == "cos":
        q_inv = _inv_norm(q)
        gi = g_inv if g_inv is not None else _inv_norm(g)
    if tensor_engine_ok(q, g, engine):
        # tensor engine: bf16 tcgen05 scan of the prepared gallery, exact fp32 re-rank, proof of exactness per query
        # (unproven queries are recomputed by the exact engine on the device; redo_count counts them)
        if prepared is None or not prepared.matches(g, metric):
            prepared = _prepared_for(g, metric)
        nbytes = lib.b200f_gallery_tc_workspace_bytes(Q, N, D, k)
        ws = _lib.workspace(nbytes, dev, "gallery_tc")
        check(lib.b200f_gallery_topk_tc(ptr(q), ptr(g), ptr(prepared.g16), ptr(prepared.bias), ptr(q_inv), ptr(gi), Q, N,
                                        int(index_offset), D, k, _METRICS[metric], prepared.operand_fmt, float(thresh),
                                        ptr(idx), ptr(score),
                                        ptr(accept), ptr(redo_count), ptr(ws), ws.numel(), stream_ptr(dev)),
              "b200f_gallery_topk_tc")
        return idx, score, accept.view(torch.bool)
    nbytes = lib.b200f_gallery_workspace_bytes(Q, N, D, k, dtype_code(q), engine)
    ws = _lib.workspace(nbytes, dev, "gallery")
    check(lib.b200f_gallery_topk(ptr(q), ptr(g), dtype_code(q), ptr(q_inv), ptr(gi), Q, N, int(index_offset), D,
                                 k, _METRICS[metric], float(thresh), engine, ptr(idx), ptr(score), ptr(accept),
                                 ptr(ws), ws.numel(), stream_ptr(dev)), "b200f_gallery_topk")
    return idx, score, accept.view(torch.bool)


_pipe_streams = {}
PIPELINE_DEPTH = 3      # measured at Q = 128 vs 1 M x 512: 270 us per call serial, 239 at depth 2, 233 at depth 3


def gallery_topk_batches(batches: Sequence[torch.Tensor], g: torch.Tensor, k: int = 1, thresh: float = 1.0,
                         metric: str = "l2eps", *, depth: int = PIPELINE_DEPTH, index_offset: int = 0,
                         engine: int = _lib.ENGINE_AUTO, prepared: Optional[PreparedGallery] = None,
                         redo_count: Optional[torch.Tensor] = None) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """gallery_topk for several query batches against ONE gallery with up to `depth` calls in flight on private CUDA
    streams.  One call is a chain of dependent kernels (query prepare -> sample scan -> bound -> main scan -> select);
    with several batches in flight the latency-bound head and tail of one call run beside the HBM-bound main scan of
    another.  The results are those of the serial calls, element for element (each call owns its scratch: the
    workspace cache is keyed by stream).  Returns [(idx, score, accept)] in the order of `batches`, ready for use on
    the caller's current stream."""
    batches = list(batches)
    if not batches:
        return []
    require_cuda(g, *batches)
    dev = g.device
    g = g.contiguous()
    if depth <= 1 or len(batches) == 1:
        return [gallery_topk(q, g, k, thresh, metric, index_offset=index_offset, engine=engine, prepared=prepared,
                             redo_count=redo_count) for q in batches]
    # what every call shares is built once, on the caller's stream, before the side streams fork from it
    if prepared is None and any(tensor_engine_ok(q, g, engine) for q in batches if q.dtype == g.dtype):
        prepared = _prepared_for(g, metric)
    g_inv = _inv_norm(g) if metric == "cos" else None
    cur = torch.cuda.current_stream(dev)
    depth = min(depth, len(batches))
    key = (str(dev), depth)
    streams = _pipe_streams.get(key)
    if streams is None:
        streams = _pipe_streams[key] = [torch.cuda.Stream(dev) for _ in range(depth)]
    for s in streams:
        s.wait_stream(cur)
    outs = []
    for i, q in enumerate(batches):
        with torch.cuda.stream(streams[i % depth]):
            res = gallery_topk(q, g, k, thresh, metric, index_offset=index_offset, g_inv=g_inv, engine=engine,
                               prepared=prepared, redo_count=redo_count)
        for t in res:
            t.record_stream(cur)                  # allocated on a side stream, consumed on the caller's
        outs.append(res)
    for s in streams:
        cur.wait_stream(s)
    return outs


def merge_topk(idx_all: torch.Tensor, score_all: torch.Tensor, thresh: float, metric: str):
    """Merge P per-shard lists [P,Q,k] (global ids) into the global top-k (lowest id wins ties)."""
    require_cuda(idx_all, score_all)
    lib = _lib.load_library()
    P, Q, k = idx_all.shape
    dev = idx_all.device
    idx = torch.empty(Q, k, dtype=torch.int64, device=dev)
    score = torch.empty(Q, k, dtype=torch.float32, device=dev)
    accept = torch.empty(Q, dtype=torch.uint8, device=dev)
    check(lib.b200f_gallery_merge(ptr(idx_all.contiguous()), ptr(score_all.contiguous()), P, Q, k,
                                  _METRICS[metric], float(thresh), ptr(idx), ptr(score), ptr(accept),
                                  stream_ptr(dev)), "b200f_gallery_merge")
    return idx, score, accept.view(torch.bool)


def compare_faces(emb, refs, thresh):
    """Drop-in for src/app.py:50-64.  emb: Tensor [1,D] (any device) or None; refs: list of dicts with
    'name' and 'embedding' (Tensor [1,D]); returns (name, min_dist, index) or ("Unknown", d, None)."""
    if emb is None or not refs:
        return "Unknown", float('inf'), None
    dev = emb.device if emb.is_cuda else _default_device()
    q = emb.reshape(1, -1).to(device=dev, dtype=torch.float32)
    g = torch.cat([r['embedding'].reshape(1, -1) for r in refs], dim=0).to(device=dev, dtype=torch.float32)
    idx, score, accept = gallery_topk(q, g, 1, thresh, "l2eps")
    i, d, ok = int(idx[0, 0].item()), float(score[0, 0].item()), bool(accept[0].item())
    if i < 0:
        return "Unknown", float('inf'), None
    return (refs[i]['name'], d, i) if ok else ("Unknown", d, None)


def cosine_class_match(emb: torch.Tensor, weight: torch.Tensor, s: float = 1.0):
    """pred = (normalize(emb) @ normalize(weight).T * s).max(1)   (hyperparameter_tuning.py:1039-1046,1076)
    without the [B,C] matrix: K4 with k=1, metric=cos, gallery = class centres.
    Returns (pred [B] int64, best_logit [B] fp32)."""
    idx, score, _ = gallery_topk(emb.to(weight.dtype) if emb.dtype != weight.dtype else emb, weight, 1,
                                 -2.0, "cos")
    return idx[:, 0], score[:, 0] * s


class GalleryIndex:
    """Device-resident reference store for the recognition loop (src/app.py: refs list,
    add/rename/delete at :428-433,477-513; pickle format of save_refs/load_refs :67-123)."""

    def __init__(self, dim: int = 512, device=None, dtype=torch.float32, capacity: int = 1024):
        self.device = torch.device(device) if device is not None else _default_device()
        self.dim, self.dtype = dim, dtype
        self.names: List[str] = []
        self._buf = torch.empty(capacity, dim, dtype=dtype, device=self.device)

    def __len__(self):
        return len(self.names)

    @property
    def embeddings(self) -> torch.Tensor:
        return self._buf[:len(self.names)]

    def add(self, name: str, embedding: torch.Tensor) -> int:
        n = len(self.names)
        if n == self._buf.shape[0]:
            grown = torch.empty(2 * n, self.dim, dtype=self.dtype, device=self.device)
            grown[:n] = self._buf
            self._buf = grown
        self._buf[n] = embedding.reshape(-1).to(device=self.device, dtype=self.dtype)
        self.names.append(name)
        return n

    def rename(self, index: int, name: str):
        self.names[index] = name

    def delete(self, index: int):
        n = len(self.names)
        if index < n - 1:
            self._buf[index:n - 1] = self._buf[index + 1:n].clone()
        del self.names[index]

    @classmethod
    def from_refs(cls, refs: Sequence[dict], device=None, dtype=torch.float32):
        dim = refs[0]['embedding'].numel() if refs else 512
        gi = cls(dim, device, dtype, capacity=max(16, len(refs)))
        for r in refs:
            gi.add(r['name'], r['embedding'])
        return gi

    @classmethod
    def from_saved(cls, saved: Sequence[dict], device=None, dtype=torch.float32):
        """From the unpickled list save_refs writes ({'name','embedding_numpy','image_path'}, app.py:82-86)."""
        refs = [{'name': r['name'], 'embedding': torch.as_tensor(r['embedding_numpy'])} for r in saved]
        return cls.from_refs(refs, device, dtype)

    def save(self, path: str, image_paths: Optional[Sequence[str]] = None):
        """Write the store as the pickle the reference's save_refs writes (src/app.py:82-91): a list of
        {'name', 'embedding_numpy' (1,D) float32, 'image_path'}; load_refs (:104-123) and this class read it back."""
        import pickle
        with open(path, "wb") as f:
            pickle.dump(self.to_saved(image_paths), f)

    @classmethod
    def load(cls, path: str, device=None, dtype=torch.float32):
        """Read a face_references.pkl written by the reference (or by save)."""
        import pickle
        with open(path, "rb") as f:
            saved = pickle.load(f)
        return cls.from_saved(saved, device, dtype)

    def to_saved(self, image_paths: Optional[Sequence[str]] = None) -> List[dict]:
        emb = self.embeddings.float().cpu().numpy()
        return [{'name': n, 'embedding_numpy': emb[i:i + 1].copy(),
                 'image_path': image_paths[i] if image_paths else None} for i, n in enumerate(self.names)]

    def prepared(self, metric: str = "l2eps") -> PreparedGallery:
        """The tensor engine's scan operand for the current contents (rebuilt after add / delete)."""
        g = self.embeddings
        cache = self.__dict__.setdefault("_prepared", {})
        pg = cache.get(metric)
        if pg is None or not pg.matches(g, metric):
            pg = PreparedGallery(g, metric)
            cache[metric] = pg
        return pg

    def match(self, emb: torch.Tensor, thresh: float = 1.0, k: int = 1, metric: str = "l2eps",
              engine: int = _lib.ENGINE_AUTO):
        """Batched compare_faces: emb [Q,D].  Returns (idx, score, accept) device tensors."""
        q = emb.reshape(-1, self.dim).to(device=self.device, dtype=self.dtype)
        g = self.embeddings
        pg = self.prepared(metric) if tensor_engine_ok(q, g, engine) else None
        return gallery_topk(q, g, k, thresh, metric, engine=engine, prepared=pg)

    def match_batches(self, embs: Sequence[torch.Tensor], thresh: float = 1.0, k: int = 1, metric: str = "l2eps",
                      engine: int = _lib.ENGINE_AUTO, depth: int = PIPELINE_DEPTH):
        """match() for several query batches with up to `depth` of them in flight (gallery_topk_batches)."""
        qs = [e.reshape(-1, self.dim).to(device=self.device, dtype=self.dtype) for e in embs]
        g = self.embeddings
        pg = self.prepared(metric) if any(tensor_engine_ok(q, g, engine) for q in qs) else None
        return gallery_topk_batches(qs, g, k, thresh, metric, depth=depth, engine=engine, prepared=pg)

    def compare_faces(self, emb: Optional[torch.Tensor], thresh: float):
        """compare_faces(emb, refs, thresh) against the resident gallery."""
        if emb is None or not self.names:
            return "Unknown", float('inf'), None
        idx, score, accept = self.match(emb, thresh, 1)
        i, d, ok = int(idx[0, 0].item()), float(score[0, 0].item()), bool(accept[0].item())
        if i < 0:
            return "Unknown", float('inf'), None
        return (self.names[i], d, i) if ok else ("Unknown", d, None)
