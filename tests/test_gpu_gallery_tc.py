"""K4 on the tensor cores (bf16 tcgen05 scan + exact fp32 re-rank + per-query proof, b200f_gallery_topk_tc) against
the exact fp32 CUDA-core engine and the oracle.  Bar: identical top-k identities and accept decisions; identities may
differ only at score ties within 1e-6 (north star); scores to fp32 summation-order accuracy."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _case(Q, N, D, seed, dev, dup=True):
    g = torch.Generator().manual_seed(seed)
    G = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    Qm = torch.nn.functional.normalize(torch.randn(Q, D, generator=g), dim=1)
    h = Q // 2
    src = torch.randint(0, N, (h,), generator=g)
    tau = 0.5 + 2.0 * torch.rand(h, 1, generator=g)
    Qm[:h] = torch.nn.functional.normalize(G[src] + tau / D ** 0.5 * torch.randn(h, D, generator=g), dim=1)
    if dup and N > 200:
        G[150] = G[17]                                        # exact duplicate rows: the first index must win
        Qm[min(1, Q - 1)] = torch.nn.functional.normalize(G[17] + 0.01 * torch.randn(D, generator=g), dim=0)
    return Qm.to(dev), G.to(dev)


def _check_same(a, b, tie=1e-6, rtol=3e-6):
    (i1, s1, a1), (i2, s2, a2) = a, b
    i1, i2 = i1.cpu().numpy(), i2.cpu().numpy()
    s1, s2 = s1.cpu().numpy().astype(np.float64), s2.cpu().numpy().astype(np.float64)
    assert np.array_equal(a1.cpu().numpy(), a2.cpu().numpy())
    np.testing.assert_allclose(s1, s2, rtol=rtol, atol=1e-9)
    diff = i1 != i2
    if diff.any():
        assert np.all(np.abs(s1[diff] - s2[diff]) <= tie), "identities differ away from a score tie"


@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
@pytest.mark.parametrize("metric", ["l2eps", "cos"])
@pytest.mark.parametrize("k", [1, 5, 16])
@pytest.mark.parametrize("Q,N,D", [(64, 5000, 512), (300, 20000, 512), (128, 70000, 128), (7, 300, 64), (1, 9000, 512),
                                   (300, 140000, 512)])   # the last two N exceed 4x the sample: pre-pass bound active
def test_tensor_engine_equals_exact_engine(cuda_device, Q, N, D, k, metric, fmt):
    import b200face
    from b200face import _lib
    q, g = _case(Q, N, D, Q + N + k, cuda_device)
    thr = 1.0 if metric == "l2eps" else 0.5
    redo = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    prep = b200face.PreparedGallery(g, metric, _lib.OPERAND_FP16 if fmt == "fp16" else _lib.OPERAND_BF16)
    tc = b200face.gallery_topk(q, g, k, thr, metric, engine=_lib.ENGINE_TCGEN05, redo_count=redo, prepared=prep)
    ex = b200face.gallery_topk(q, g, k, thr, metric, engine=_lib.ENGINE_SIMT)
    torch.cuda.synchronize()
    assert _lib.load_library().b200f_umma_timeout_flag(1) == 0
    _check_same(tc, ex)
    if N >= 5000 and k <= 5 and fmt == "fp16":
        assert int(redo) <= max(1, Q // 50), "the proof of exactness should hold for almost every query"


def test_tensor_engine_vs_oracle(cuda_device):
    import b200face
    from b200face import _lib
    q, g = _case(200, 4000, 512, 3, cuda_device)
    for metric, thr in (("l2eps", 1.0), ("cos", 0.5)):
        idx, score, acc = b200face.gallery_topk(q, g, 5, thr, metric, engine=_lib.ENGINE_TCGEN05)
        ridx, rscore, racc = oracle.gallery_topk(q.cpu().numpy(), g.cpu().numpy(), 5, thr, metric)
        assert np.array_equal(acc.cpu().numpy(), racc)
        np.testing.assert_allclose(score.cpu().numpy(), rscore, rtol=1e-5, atol=1e-7)
        i = idx.cpu().numpy()
        diff = i != ridx
        assert np.all(np.abs(score.cpu().numpy()[diff] - rscore[diff]) <= 1e-6)


def test_crowded_gallery_falls_back_to_exact(cuda_device):
    """A gallery of near-identical rows: the bf16 scan cannot separate them, the proof fails, and the exact engine
    recomputes those queries on the device -- the results still equal the exact engine's."""
    import b200face
    from b200face import _lib
    g0 = torch.Generator().manual_seed(5)
    D, N, Q = 512, 6000, 40
    centre = torch.nn.functional.normalize(torch.randn(D, generator=g0), dim=0)
    G = torch.nn.functional.normalize(centre + 1e-4 * torch.randn(N, D, generator=g0), dim=1).to(cuda_device)
    Qm = torch.nn.functional.normalize(centre + 1e-4 * torch.randn(Q, D, generator=g0), dim=1).to(cuda_device)
    redo = torch.zeros(1, dtype=torch.int32, device=cuda_device)
    tc = b200face.gallery_topk(Qm, G, 5, 1.0, "l2eps", engine=_lib.ENGINE_TCGEN05, redo_count=redo)
    ex = b200face.gallery_topk(Qm, G, 5, 1.0, "l2eps", engine=_lib.ENGINE_SIMT)
    torch.cuda.synchronize()
    assert int(redo) > 0                                      # the fallback did run
    _check_same(tc, ex, tie=0.0, rtol=0.0)                    # flagged queries are bitwise the exact engine's


def test_gallery_index_reuses_prepared_operand(cuda_device):
    import b200face
    from b200face import _lib
    q, g = _case(96, 8000, 512, 11, cuda_device)
    index = b200face.GalleryIndex(512, device=cuda_device, capacity=8000)
    index._buf[:8000] = g
    index.names = [f"id{i}" for i in range(8000)]
    r1 = index.match(q, 1.0, 5, engine=_lib.ENGINE_TCGEN05)
    pg = index.prepared("l2eps")
    r2 = index.match(q, 1.0, 5, engine=_lib.ENGINE_TCGEN05)
    assert index.prepared("l2eps") is pg                      # not rebuilt
    _check_same(r1, r2, tie=0.0, rtol=0.0)
    _check_same(r1, b200face.gallery_topk(q, g, 5, 1.0, "l2eps", engine=_lib.ENGINE_SIMT))
    index.add("new", g[5])                                    # contents changed -> operand rebuilt
    assert index.prepared("l2eps") is not pg


def test_streaming_regime_properties(cuda_device):
    """Q = 128 against 1M x 512 (BASELINE's gallery size): the oracle is too slow, so check properties -- planted
    copies are found at rank 1 with their exact distance, results are invariant to the engine on a query subset."""
    import b200face
    from b200face import _lib
    dev = cuda_device
    g0 = torch.Generator(device=dev).manual_seed(9)
    N, D, Q = 1_000_000, 512, 128
    G = torch.nn.functional.normalize(torch.randn(N, D, generator=g0, device=dev), dim=1)
    src = torch.randint(0, N, (Q,), generator=g0, device=dev)
    Qm = torch.nn.functional.normalize(G[src] + 0.3 / D ** 0.5 * torch.randn(Q, D, generator=g0, device=dev), dim=1)
    redo = torch.zeros(1, dtype=torch.int32, device=dev)
    idx, score, acc = b200face.gallery_topk(Qm, G, 5, 1.0, "l2eps", engine=_lib.ENGINE_TCGEN05, redo_count=redo)
    assert torch.equal(idx[:, 0], src)
    d = torch.linalg.vector_norm(Qm - G[src] + 1e-6, dim=1)
    torch.testing.assert_close(score[:, 0], d, rtol=1e-5, atol=1e-7)
    assert bool(acc.all())
    assert bool((score[:, 1:] >= score[:, :-1]).all())        # ascending
    sub = slice(0, 8)
    ex = b200face.gallery_topk(Qm[sub].contiguous(), G, 5, 1.0, "l2eps", engine=_lib.ENGINE_SIMT)
    _check_same((idx[sub], score[sub], acc[sub]), ex)
    assert int(redo) <= 4


def test_unscaled_gallery_uses_bf16_and_stays_exact(cuda_device):
    """Rows far outside fp16's range: PreparedGallery picks bf16 operands by itself; forcing fp16 overflows in the
    scan, which must only cost speed (the affected queries go to the exact engine), never correctness."""
    import b200face
    from b200face import _lib
    g0 = torch.Generator().manual_seed(2)
    G = (torch.randn(3000, 256, generator=g0) * 3.0e4).to(cuda_device)
    Qm = (G[:50] + 100.0 * torch.randn(50, 256, generator=g0).to(cuda_device)).contiguous()
    ex = b200face.gallery_topk(Qm, G, 5, 1.0e9, "l2eps", engine=_lib.ENGINE_SIMT)
    auto = b200face.PreparedGallery(G, "l2eps")
    assert auto.operand_fmt == _lib.OPERAND_BF16
    _check_same(b200face.gallery_topk(Qm, G, 5, 1.0e9, "l2eps", engine=_lib.ENGINE_TCGEN05, prepared=auto), ex, rtol=1e-5)
    forced = b200face.PreparedGallery(G, "l2eps", _lib.OPERAND_FP16)
    _check_same(b200face.gallery_topk(Qm, G, 5, 1.0e9, "l2eps", engine=_lib.ENGINE_TCGEN05, prepared=forced), ex, rtol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fmt", ["fp16", "bf16"])
@pytest.mark.parametrize("D", [512, 64, 100])                 # 100: not a multiple of 8 -> the scalar prepare kernel
def test_prepared_operand_matches_torch(cuda_device, D, fmt, dtype):
    """b200f_gallery_prepare (vector and scalar kernels): the 16-bit rows are the round-to-nearest cast of the fp32
    values (L2EPS) / of the normalised rows (COS); bias = |g|^2 - 2e-6 sum(g); the two trailing slots hold the largest
    row norm and the count of rows the 16-bit format lost."""
    import b200face
    from b200face import _lib
    gen = torch.Generator().manual_seed(D)
    N = 1003
    g = (torch.randn(N, D, generator=gen) * 0.3).to(dtype).to(cuda_device)
    t16 = torch.float16 if fmt == "fp16" else torch.bfloat16
    code = _lib.OPERAND_FP16 if fmt == "fp16" else _lib.OPERAND_BF16
    gf = g.float()
    pg = b200face.PreparedGallery(g, "l2eps", code)
    torch.cuda.synchronize()
    assert torch.equal(pg.g16, gf.to(t16))
    ref_bias = (gf.double() ** 2).sum(1) - 2e-6 * gf.double().sum(1)
    np.testing.assert_allclose(pg.bias[:N].cpu().numpy(), ref_bias.cpu().numpy(), rtol=3e-6, atol=1e-7)
    np.testing.assert_allclose(float(pg.bias[N]), float(gf.double().norm(dim=1).max()), rtol=3e-6)
    assert int(pg.bias[N + 1:].view(torch.int32)[0]) == 0
    pc = b200face.PreparedGallery(g, "cos", code)
    torch.cuda.synchronize()
    ghat = (gf / gf.norm(dim=1, keepdim=True).clamp_min(1e-12))
    ulp = 2.0 ** -10 if fmt == "fp16" else 2.0 ** -7          # one unit in the last place of values below 1
    assert float((pc.g16.float() - ghat).abs().max()) <= ulp
    assert float(pc.bias[:N].abs().max()) == 0.0


def test_prepared_operand_flags_fp16_overflow(cuda_device):
    import b200face
    from b200face import _lib
    g = torch.randn(64, 512, device=cuda_device)
    g[5, 7] = 1.0e5                                           # finite in fp32, infinite in fp16
    g[9, 500] = -7.0e4
    pg = b200face.PreparedGallery(g, "l2eps", _lib.OPERAND_FP16)
    torch.cuda.synchronize()
    assert int(pg.bias[64 + 1:].view(torch.int32)[0]) == 2


@pytest.mark.parametrize("metric", ["l2eps", "cos"])
@pytest.mark.parametrize("depth", [2, 3])
def test_batches_in_flight_equal_serial_calls(cuda_device, metric, depth):
    """gallery_topk_batches / GalleryIndex.match_batches: several query batches in flight on private streams give,
    element for element, what the serial calls give (tensor engine and exact engine)."""
    import b200face
    from b200face import _lib
    q, g = _case(5 * 96, 30000, 512, 11, cuda_device)
    thr = 1.0 if metric == "l2eps" else 0.5
    batches = [q[i * 96:(i + 1) * 96] for i in range(5)] + [q[:7]]
    for engine in (_lib.ENGINE_TCGEN05, _lib.ENGINE_SIMT):
        serial = [b200face.gallery_topk(b, g, 5, thr, metric, engine=engine) for b in batches]
        for _ in range(3):                                    # repeated: scratch of a stream is reused across rounds
            piped = b200face.gallery_topk_batches(batches, g, 5, thr, metric, depth=depth, engine=engine)
        torch.cuda.synchronize()
        assert len(piped) == len(batches)
        for a, b in zip(piped, serial):
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert b200face.gallery_topk_batches([], g, 5, thr, metric) == []
    index = b200face.GalleryIndex(512, cuda_device, capacity=g.shape[0])
    index._buf[:] = g
    index.names = [f"id{i}" for i in range(g.shape[0])]
    got = index.match_batches(batches, thr, 5, metric, depth=depth)
    want = [index.match(b, thr, 5, metric) for b in batches]
    torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])


def test_repeated_calls_reuse_the_prepared_operand_of_the_same_tensor(cuda_device):
    """gallery_topk(q, g) without a PreparedGallery: the operand of the SAME tensor object is built once (weak reference +
    version counter), rebuilt after an in-place change, never shared with another tensor, and dropped with the tensor."""
    import gc
    from b200face import gallery as G
    from b200face import _lib
    g0 = torch.Generator().manual_seed(3)
    gal = torch.randn(20000, 256, generator=g0).to(cuda_device)
    q = (gal[:300] + 0.01 * torch.randn(300, 256, generator=g0).to(cuda_device)).contiguous()
    G._prepared_cache.clear()
    i1, s1, a1 = G.gallery_topk(q, gal, 3, 1.0, engine=_lib.ENGINE_TCGEN05)
    prep = G._prepared_cache[id(gal)][1]
    i2, s2, a2 = G.gallery_topk(q, gal, 3, 1.0, engine=_lib.ENGINE_TCGEN05)
    assert G._prepared_cache[id(gal)][1] is prep                   # reused
    assert torch.equal(i1, i2) and torch.equal(s1, s2)
    assert torch.equal(i1[:, 0].cpu(), torch.arange(300))
    gal.mul_(-1.0)                                                  # in place: the version counter moves
    i3, _, _ = G.gallery_topk(q, gal, 3, 1.0, engine=_lib.ENGINE_TCGEN05)
    assert G._prepared_cache[id(gal)][1] is not prep
    ref3, _, _ = G.gallery_topk(q, gal, 3, 1.0, engine=_lib.ENGINE_SIMT)
    assert torch.equal(i3, ref3)
    other = gal.clone()
    G.gallery_topk(q, other, 3, 1.0, engine=_lib.ENGINE_TCGEN05)
    assert G._prepared_cache[id(other)][1] is not G._prepared_cache[id(gal)][1]
    n = len(G._prepared_cache)
    del other
    gc.collect()
    assert len(G._prepared_cache) == n - 1                          # the entry died with its tensor
