"""Pin the CPU oracle against vectors produced by the reference itself
(tests/golden/make_golden.py ran /root/reference/src/face_models.py and src/app.py).
CPU only."""
import numpy as np
import pytest

from conftest import cfg_from_golden, golden, head_golden_names, rel_err
import oracle
from oracle import arcface_oracle as ao


@pytest.mark.parametrize("name", head_golden_names())
def test_head_forward_matches_reference(name):
    d = golden(f"head_{name}.npz")
    cfg = cfg_from_golden(d)
    mf, sf = ao.warmup_schedule(cfg)
    if cfg.training:
        assert mf == pytest.approx(float(d["cfg"][7]), abs=0) and sf == pytest.approx(float(d["cfg"][8]), abs=0)
    z, cmax, cmin, nan_seen = oracle.arc_logits(d["x"], d["w"], d["y"], cfg, dtype=np.float64)
    assert not nan_seen
    # reference is fp32: 1e-5 relative is the fp32 bar of the north star.  The "extreme" cases
    # plant cos == +-1: acos is infinitely ill-conditioned there (theta = sqrt(2(1-c)), one fp32
    # ulp of the dot product moves theta by ~1e-4), so those rows only agree to ~1e-4.
    assert rel_err(z, d["logits"]) < (2e-4 if name.startswith("extreme") else 1e-5)
    assert cmax == pytest.approx(float(d["cos_max"]), abs=2e-6)
    assert cmin == pytest.approx(float(d["cos_min"]), abs=2e-6)
    loss, _ = oracle.smoothed_cross_entropy(z, d["y"], cfg.label_smoothing)
    assert float(loss) == pytest.approx(float(d["loss"]), rel=1e-5)


@pytest.mark.parametrize("name", head_golden_names())
def test_head_backward_matches_reference(name):
    d = golden(f"head_{name}.npz")
    cfg = cfg_from_golden(d)
    r = oracle.head_forward_backward(d["x"], d["w"], d["y"], cfg, dtype=np.float64)
    # the extreme cases sit on the fp32 clamp where sin(theta) ~ 4.9e-4: the reference's own
    # fp32 autograd is only good to ~1e-3 there, everything else holds 1e-5
    tol = 5e-3 if name.startswith("extreme") else 2e-5
    assert rel_err(r["dx"], d["dx"]) < tol
    assert rel_err(r["dw"], d["dw"]) < tol
    assert float(r["loss"]) == pytest.approx(float(d["loss"]), rel=1e-5)


@pytest.mark.parametrize("name", ["epoch5", "postwarm", "easy", "ls15_m03"])
def test_head_fp32_oracle_close_to_fp64(name):
    d = golden(f"head_{name}.npz")
    cfg = cfg_from_golden(d)
    r32 = oracle.head_forward_backward(d["x"], d["w"], d["y"], cfg, dtype=np.float32)
    r64 = oracle.head_forward_backward(d["x"], d["w"], d["y"], cfg, dtype=np.float64)
    assert rel_err(r32["dx"], r64["dx"]) < 1e-5
    assert rel_err(r32["dw"], r64["dw"]) < 1e-5


def test_hook_matches_arcfacenet():
    """face_models.py:538-570: no renormalisation on the first step, Frobenius-norm clip on
    the second (n ~ 1.9 > thr 0.5)."""
    d = golden("hook_arcfacenet.npz")
    max_gn, phase, epoch = [float(v) for v in d["meta"]]
    cfg = oracle.HeadConfig(current_epoch=int(epoch), training=True, label_smoothing=0.05)
    # step 0: hook not registered yet
    r0 = oracle.head_forward_backward(d["emb0"], d["w"], d["y0"], cfg, hook=None)
    assert rel_err(r0["dx"], d["demb0"]) < 2e-5
    assert rel_err(r0["dw"], d["dw0"]) < 2e-5
    assert float(r0["loss"]) == pytest.approx(float(d["loss0"]), rel=1e-5)
    # step 1: hook active
    hook = dict(max_grad_norm=max_gn, phase=int(phase), current_epoch=int(epoch))
    r1 = oracle.head_forward_backward(d["emb1"], d["w"], d["y1"], cfg, hook=hook)
    assert r1["gnorm"] == pytest.approx(float(d["last_grad_norm1"]), rel=1e-5)
    assert r1["kappa"] < 1.0
    assert rel_err(r1["dx"], d["demb1"]) < 2e-5
    assert rel_err(r1["dw"], d["dw1"]) < 2e-5


@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_partial_fc_algebra(n_shards):
    d = golden("head_postwarm.npz")
    cfg = cfg_from_golden(d)
    full = oracle.head_forward_backward(d["x"], d["w"], d["y"], cfg)
    sh = oracle.sharded_head_forward_backward(d["x"], d["w"], d["y"], cfg, n_shards)
    assert float(sh["loss"]) == pytest.approx(float(full["loss"]), rel=1e-12)
    np.testing.assert_allclose(sh["lse"], full["lse"], rtol=1e-12)
    np.testing.assert_allclose(np.concatenate(sh["shard_logits"], axis=1), full["logits"], rtol=1e-12)


def test_gallery_fixture_distance_matrix():
    """The reference's only real data: face_references.pkl, 7 unit-norm FaceNet embeddings."""
    d = golden("gallery_fixture.npz")
    dist = oracle.pairwise_distance_eps(d["emb"], d["emb"])
    np.testing.assert_allclose(dist, d["dist"], rtol=2e-6, atol=1e-9)
    assert np.allclose(np.diag(dist), np.sqrt(512) * 1e-6, rtol=1e-3)
    refs = [{"name": str(n), "embedding": e[None, :]} for n, e in zip(d["names"], d["emb"])]
    for i in range(7):
        name, dmin, idx = oracle.compare_faces(d["emb"][i:i + 1], refs, 1.0)
        assert (name, idx) == (str(d["self_name"][i]), int(d["self_idx"][i]))
        assert dmin == pytest.approx(float(d["self_dist"][i]), rel=2e-6)
    t = 0
    for thr in d["loo_thr"]:
        for i in range(7):
            sub = refs[:i] + refs[i + 1:]
            name, dmin, idx = oracle.compare_faces(d["emb"][i:i + 1], sub, float(thr))
            assert name == str(d["loo_name"][t])
            assert (-1 if idx is None else idx) == int(d["loo_idx"][t])
            assert dmin == pytest.approx(float(d["loo_dist"][t]), rel=2e-6)
            t += 1


def test_gallery_synth_matches_reference_loop():
    d = golden("gallery_synth.npz")
    idx, sc, acc = oracle.gallery_topk(d["q"], d["g"], 1, float(d["thresh"]), "l2eps")
    ref_idx = d["idx"]
    assert np.array_equal(acc, ref_idx >= 0)
    assert np.array_equal(idx[acc, 0], ref_idx[acc])
    np.testing.assert_allclose(sc[:, 0], d["dist"], rtol=2e-6, atol=1e-9)
    assert idx[1, 0] == 17                      # planted duplicate rows 17/150: first wins
    # batched form == the verbatim loop
    refs = [{"name": f"id{j}", "embedding": d["g"][j:j + 1]} for j in range(d["g"].shape[0])]
    for i in range(0, d["q"].shape[0], 5):
        name, dmin, k = oracle.compare_faces(d["q"][i:i + 1], refs, 1.0)
        assert (-1 if k is None else k) == int(ref_idx[i])
    assert oracle.compare_faces(None, refs, 1.0) == ("Unknown", float("inf"), None)
    assert oracle.compare_faces(d["q"][:1], [], 1.0) == ("Unknown", float("inf"), None)


def test_cosine_class_match():
    d = golden("cosine_match.npz")
    pred, best = oracle.cosine_class_match(d["emb"], d["w"], float(d["s"]))
    assert np.array_equal(pred, d["pred"])
    np.testing.assert_allclose(best, d["best"], rtol=1e-5)


def test_topk_merge_equals_global():
    rng = np.random.default_rng(0)
    g = rng.standard_normal((300, 64)).astype(np.float32)
    q = rng.standard_normal((9, 64)).astype(np.float32)
    g[200] = g[3]
    q[0] = g[3]
    for metric, largest in (("l2eps", False), ("cos", True)):
        gi, gs, _ = oracle.gallery_topk(q, g, 5, 1.0, metric)
        parts_i, parts_s = [], []
        for lo, hi in ((0, 100), (100, 250), (250, 300)):
            i, s, _ = oracle.gallery_topk(q, g[lo:hi], 5, 1.0, metric)
            parts_i.append(np.where(i >= 0, i + lo, -1)); parts_s.append(s)
        mi, ms = oracle.merge_topk_shards(parts_i, parts_s, 5, largest)
        assert np.array_equal(mi, gi)
        np.testing.assert_allclose(ms, gs, rtol=1e-6)   # BLAS blocking differs per shard shape


@pytest.mark.parametrize("easy", [False, True])
@pytest.mark.parametrize("epoch,ls", [(0, 0.05), (12, 0.0), (12, 0.15)])
def test_closed_form_gradients_match_finite_differences(easy, epoch, ls):
    """A pin of the oracle that does not go through autograd at all: the closed-form dL/dx and dL/dW of
    head_forward_backward against central differences of its own loss in fp64 (size-independent property; the golden
    vectors above pin the same quantities to the reference's autograd)."""
    rng = np.random.default_rng(7 + epoch + int(easy))
    B, C, D = 5, 7, 6
    x = rng.standard_normal((B, D))
    w = rng.standard_normal((C, D)) * 0.5
    y = rng.integers(0, C, size=B)
    x[0] = 2.0 * w[y[0]] + 0.1 * rng.standard_normal(D)       # a row near its class centre: the margin branch matters
    cfg = oracle.HeadConfig(current_epoch=epoch, training=True, label_smoothing=ls, easy_margin=easy)
    ref = oracle.head_forward_backward(x, w, y, cfg, dtype=np.float64)
    loss = lambda xx, ww: float(oracle.head_forward_backward(xx, ww, y, cfg, dtype=np.float64)["loss"])
    h = 1e-6
    fd_x = np.zeros_like(x)
    for i in range(B):
        for j in range(D):
            xp, xm = x.copy(), x.copy()
            xp[i, j] += h; xm[i, j] -= h
            fd_x[i, j] = (loss(xp, w) - loss(xm, w)) / (2 * h)
    fd_w = np.zeros_like(w)
    for i in range(C):
        for j in range(D):
            wp, wm = w.copy(), w.copy()
            wp[i, j] += h; wm[i, j] -= h
            fd_w[i, j] = (loss(x, wp) - loss(x, wm)) / (2 * h)
    assert rel_err(ref["dx"], fd_x) < 1e-6
    assert rel_err(ref["dw"], fd_w) < 1e-6
