"""GPU parity of the gallery match (kernel K4 + merge through the C ABI) against the reference-pinned
oracle and the committed golden vectors.  Bar (north star): top-k identities and accept/reject
decisions bit-exact except at score ties within 1e-6."""
import numpy as np
import pytest
import torch

from conftest import golden
import oracle

pytestmark = pytest.mark.gpu


def _ids_equal_up_to_ties(idx, score, ridx, rscore, tie=1e-6):
    """Identities must match; where they differ the two scores involved must be within `tie`."""
    idx, ridx = np.asarray(idx), np.asarray(ridx)
    diff = idx != ridx
    if not diff.any():
        return True
    return bool(np.all(np.abs(np.asarray(score, dtype=np.float64)[diff] - np.asarray(rscore, dtype=np.float64)[diff]) <= tie))


def test_reference_fixture_7x7(cuda_device):
    """face_references.pkl: the reference's only real data, frozen with the reference's own results."""
    import b200face
    d = golden("gallery_fixture.npz")
    emb = torch.tensor(d["emb"])
    refs = [{"name": str(n), "embedding": emb[i:i + 1], "image": None} for i, n in enumerate(d["names"])]
    for i in range(7):
        name, dist, idx = b200face.compare_faces(emb[i:i + 1], refs, 1.0)
        assert (name, idx) == (str(d["self_name"][i]), int(d["self_idx"][i]))
        assert dist == pytest.approx(float(d["self_dist"][i]), rel=2e-6)
    t = 0
    for thr in d["loo_thr"]:
        for i in range(7):
            sub = refs[:i] + refs[i + 1:]
            name, dist, idx = b200face.compare_faces(emb[i:i + 1].to(cuda_device), sub, float(thr))
            assert name == str(d["loo_name"][t])
            assert (-1 if idx is None else idx) == int(d["loo_idx"][t])
            assert dist == pytest.approx(float(d["loo_dist"][t]), rel=2e-6)
            t += 1
    # full distance matrix through the batched entry
    idx, score, acc = b200face.gallery_topk(emb.to(cuda_device), emb.to(cuda_device), 7, 1.0, "l2eps")
    order = np.argsort(d["dist"], axis=1, kind="stable")
    assert np.array_equal(idx.cpu().numpy(), order)
    np.testing.assert_allclose(score.cpu().numpy(), np.take_along_axis(d["dist"], order, 1), rtol=2e-6, atol=1e-10)
    assert acc.all()


def test_reference_loop_synthetic(cuda_device):
    """Golden run of the verbatim compare_faces loop on the scaled-down cfg2 recipe."""
    import b200face
    d = golden("gallery_synth.npz")
    q, g = torch.tensor(d["q"], device=cuda_device), torch.tensor(d["g"], device=cuda_device)
    idx, score, acc = b200face.gallery_topk(q, g, 1, float(d["thresh"]), "l2eps")
    ref_idx = d["idx"]
    assert np.array_equal(acc.cpu().numpy(), ref_idx >= 0)
    assert np.array_equal(idx.cpu().numpy()[ref_idx >= 0, 0], ref_idx[ref_idx >= 0])
    np.testing.assert_allclose(score.cpu().numpy()[:, 0], d["dist"], rtol=2e-6, atol=1e-10)
    assert int(idx[1, 0]) == 17                              # duplicate gallery rows 17/150: first wins
    gi = b200face.GalleryIndex.from_refs([{"name": f"id{j}", "embedding": torch.tensor(d["g"][j:j + 1])}
                                          for j in range(d["g"].shape[0])], device=cuda_device)
    for i in range(0, q.shape[0], 3):
        name, dist, k = gi.compare_faces(q[i:i + 1], 1.0)
        assert (-1 if k is None else k) == int(ref_idx[i])
        assert name == str(d["name"][i])
    assert gi.compare_faces(None, 1.0) == ("Unknown", float("inf"), None)


@pytest.mark.parametrize("metric", ["l2eps", "cos"])
@pytest.mark.parametrize("Q,N,D,k", [(300, 5000, 512, 5), (1, 1, 512, 1), (129, 257, 64, 16), (17, 3, 40, 8),
                                      (64, 1000, 512, 1), (1000, 10000, 512, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_vs_oracle(cuda_device, metric, Q, N, D, k, dtype):
    import b200face
    if Q * N > 2_000_000 and dtype == torch.bfloat16:
        pytest.skip("cfg2 full size is checked once, in fp32")
    g_ = torch.Generator().manual_seed(Q * 7 + N)
    G = torch.nn.functional.normalize(torch.randn(N, D, generator=g_), dim=1)
    Qm = torch.nn.functional.normalize(torch.randn(Q, D, generator=g_), dim=1)
    half = min(Q // 2, N)
    if half:
        tau = 0.5 + 2.0 * torch.rand(half, 1, generator=g_)
        src = torch.randint(0, N, (half,), generator=g_)
        Qm[:half] = torch.nn.functional.normalize(G[src] + tau / D ** 0.5 * torch.randn(half, D, generator=g_), dim=1)
    if N > 20:
        G[N - 1] = G[5]; Qm[0] = G[5]                       # planted tie, lowest index must win
    G, Qm = G.to(dtype), Qm.to(dtype)
    thresh = 1.0 if metric == "l2eps" else 0.5
    idx, score, acc = b200face.gallery_topk(Qm.to(cuda_device), G.to(cuda_device), k, thresh, metric)
    ridx, rscore, racc = oracle.gallery_topk(Qm.float().numpy(), G.float().numpy(), k, thresh, metric)
    idx, score, acc = idx.cpu().numpy(), score.cpu().numpy(), acc.cpu().numpy()
    kk = min(k, N)
    assert _ids_equal_up_to_ties(idx[:, :kk], score[:, :kk], ridx[:, :kk], rscore[:, :kk])
    assert (idx[:, :kk] == ridx[:, :kk]).mean() > 0.999
    np.testing.assert_allclose(score[:, :kk], rscore[:, :kk], rtol=3e-6, atol=2e-7)
    assert np.all(idx[:, kk:] == -1)
    near = np.abs(rscore[:, 0].astype(np.float64) - thresh) <= 1e-6
    assert np.array_equal(acc[~near], racc[~near])
    if N > 20:
        assert idx[0, 0] == 5


def test_empty_gallery_and_errors(cuda_device):
    import b200face
    q = torch.randn(4, 64, device=cuda_device)
    idx, score, acc = b200face.gallery_topk(q, torch.empty(0, 64, device=cuda_device), 3, 1.0)
    assert torch.all(idx == -1) and torch.all(torch.isinf(score)) and not acc.any()
    with pytest.raises(ValueError):
        b200face.gallery_topk(q, q, 17, 1.0)
    with pytest.raises(TypeError):
        b200face.gallery_topk(q, q.bfloat16(), 1, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        b200face.gallery_topk(q.cpu(), q.cpu(), 1, 1.0)


def test_cosine_class_match_golden(cuda_device):
    """hyperparameter_tuning.py:1039-1046,1076 recorded outputs."""
    import b200face
    d = golden("cosine_match.npz")
    pred, best = b200face.cosine_class_match(torch.tensor(d["emb"], device=cuda_device),
                                             torch.tensor(d["w"], device=cuda_device), float(d["s"]))
    assert np.array_equal(pred.cpu().numpy(), d["pred"])
    np.testing.assert_allclose(best.cpu().numpy(), d["best"], rtol=1e-5)


def test_merge_kernel_equals_global(cuda_device):
    import b200face
    from b200face.gallery import merge_topk
    g_ = torch.Generator().manual_seed(3)
    G = torch.nn.functional.normalize(torch.randn(3000, 128, generator=g_), dim=1).to(cuda_device)
    Q = torch.nn.functional.normalize(torch.randn(200, 128, generator=g_), dim=1).to(cuda_device)
    G[2500] = G[10]; Q[0] = G[10]
    for metric, thr in (("l2eps", 1.2), ("cos", 0.2)):
        gi, gs, ga = b200face.gallery_topk(Q, G, 5, thr, metric)
        parts = [(0, 700), (700, 701), (701, 2400), (2400, 3000)]
        li, ls = zip(*[b200face.gallery_topk(Q, G[a:b], 5, thr, metric, index_offset=a)[:2] for a, b in parts])
        mi, ms, ma = merge_topk(torch.stack(li), torch.stack(ls), thr, metric)
        assert torch.equal(mi, gi) and torch.equal(ma, ga)
        torch.testing.assert_close(ms, gs, rtol=0, atol=0)


def test_gallery_index_edit_operations(cuda_device):
    """src/app.py:428-433,477-513: add / rename / delete references between matches."""
    import b200face
    gi = b200face.GalleryIndex(dim=32, device=cuda_device, capacity=2)
    vecs = torch.nn.functional.normalize(torch.randn(5, 32), dim=1)
    for i in range(5):
        gi.add(f"p{i}", vecs[i:i + 1])
    assert len(gi) == 5
    assert gi.compare_faces(vecs[3:4], 1.0)[::2] == ("p3", 3)
    gi.rename(3, "renamed")
    gi.delete(1)
    assert gi.compare_faces(vecs[3:4], 1.0)[::2] == ("renamed", 2)
    assert gi.compare_faces(vecs[1:2], 0.01)[::2] == ("Unknown", None)
    saved = gi.to_saved()
    again = b200face.GalleryIndex.from_saved(saved, device=cuda_device)
    assert again.names == gi.names and torch.equal(again.embeddings, gi.embeddings)


def test_gallery_index_pickle_is_the_reference_format(cuda_device, tmp_path):
    """save() writes what the reference's save_refs writes (src/app.py:82-91: list of {'name','embedding_numpy'
    (1,D) float32,'image_path'}), and load() reads the reference's own face_references.pkl layout back."""
    import pickle
    import b200face
    d = golden("gallery_fixture.npz")
    emb = torch.tensor(d["emb"])
    refs = [{"name": str(n), "embedding": emb[i:i + 1]} for i, n in enumerate(d["names"])]
    gi = b200face.GalleryIndex.from_refs(refs, device=cuda_device)
    path = tmp_path / "face_references.pkl"
    gi.save(str(path), image_paths=[f"face_references/{n}_{i}.jpg" for i, n in enumerate(gi.names)])
    saved = pickle.load(open(path, "rb"))
    assert isinstance(saved, list) and set(saved[0]) == {"name", "embedding_numpy", "image_path"}
    assert saved[0]["embedding_numpy"].shape == (1, emb.shape[1]) and saved[0]["embedding_numpy"].dtype == np.float32
    again = b200face.GalleryIndex.load(str(path), device=cuda_device)
    assert again.names == [str(n) for n in d["names"]]
    assert torch.equal(again.embeddings.cpu(), emb)
    for i in range(len(refs)):
        assert again.compare_faces(emb[i:i + 1], 1.0)[::2] == (str(d["self_name"][i]), int(d["self_idx"][i]))


def test_cfg5_scale_properties(cuda_device):
    """A 1M x 512 gallery shard layout at reduced Q: every gallery row queried against the gallery
    finds itself first at distance sqrt(D)*1e-6 (the eps term), the result is independent of how the
    gallery is sharded, and appending rows never worsens a best match."""
    import b200face
    from b200face.gallery import merge_topk
    N, D, Q = 200_000, 512, 256
    g_ = torch.Generator(device=cuda_device).manual_seed(5)
    G = torch.nn.functional.normalize(torch.randn(N, D, generator=g_, device=cuda_device), dim=1)
    pick = torch.randint(0, N, (Q,), generator=g_, device=cuda_device)
    Qm = G[pick].clone()
    idx, score, acc = b200face.gallery_topk(Qm, G, 5, 1.0, "l2eps")
    assert torch.equal(idx[:, 0], pick) and acc.all()
    torch.testing.assert_close(score[:, 0], torch.full((Q,), D ** 0.5 * 1e-6, device=cuda_device), rtol=1e-3, atol=0)
    assert torch.all(score[:, 1:] > 1.2)                     # random unit vectors: d ~ 1.41
    shards = [(N * r // 8, N * (r + 1) // 8) for r in range(8)]
    li, ls = zip(*[b200face.gallery_topk(Qm, G[a:b], 5, 1.0, "l2eps", index_offset=a)[:2] for a, b in shards])
    mi, ms, ma = merge_topk(torch.stack(li), torch.stack(ls), 1.0, "l2eps")
    assert torch.equal(mi, idx) and torch.equal(ms, score) and torch.equal(ma, acc)
    idx_h, score_h, _ = b200face.gallery_topk(Qm, G[: N // 2], 1, 1.0, "l2eps")
    assert torch.all(score[:, 0] <= score_h[:, 0])
