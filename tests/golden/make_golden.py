#!/usr/bin/env python
"""Freeze golden vectors from the REFERENCE ITSELF (run in the build container only).

The reference ships no tests, so there are no reference-owned golden vectors; this script
executes the reference's own code -- loaded BY FILE PATH, never ``import src`` (that pulls
facenet_pytorch and creates directories at import time) -- on seeded inputs and stores
inputs + outputs as small .npz fixtures next to this file.  /root/reference does not exist
on the GPU box, so tests only ever read the committed .npz files.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Reference code exercised:
  src/face_models.py:297-445  ArcMarginProduct   (head_*.npz)
  src/face_models.py:447-613  ArcFaceNet + its backward hook  (hook_arcfacenet.npz)
  src/training.py:341         nn.CrossEntropyLoss(label_smoothing=...)
  src/app.py:50-64            compare_faces      (gallery_*.npz)
  src/hyperparameter_tuning.py:1039-1046,1076    cosine class-centre match (cosine_match.npz)
  face_references/face_references.pkl            the only real-data fixture
"""
import importlib.util
import os
import pickle
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_face_models():
    import torchvision.models as tvm
    orig = tvm.resnet18
    tvm.resnet18 = lambda *a, **k: orig(weights=None)      # offline: no IMAGENET download
    spec = importlib.util.spec_from_file_location("ref_face_models", f"{REF}/src/face_models.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_app():
    sys.modules.setdefault("streamlit", MagicMock())
    stub = types.ModuleType("facenet_pytorch")
    stub.MTCNN = object
    stub.InceptionResnetV1 = object
    sys.modules.setdefault("facenet_pytorch", stub)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())                            # app.py:26 makedirs in CWD
    try:
        spec = importlib.util.spec_from_file_location("ref_app", f"{REF}/src/app.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def head_case(fm, name, B, C, D, seed, *, epoch=0, easy=False, training=True, ls=0.05,
              s=32.0, m=0.5, planted=0, bf16_inputs=False, warm_up_epochs=10,
              extreme=False):
    g = torch.Generator().manual_seed(seed)
    head = fm.ArcMarginProduct(D, C, s=s, m=m, use_warm_up=True, easy_margin=easy)
    with torch.no_grad():
        head.weight.copy_(torch.randn(C, D, generator=g) * 0.05)
    head.warm_up_epochs = warm_up_epochs
    head.update_epoch(epoch)
    head.train(training)
    x = torch.randn(B, D, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    for i in range(planted):                                # rows close to their class centre
        x[i] = 3.0 * head.weight[y[i]].detach() + 0.3 * torch.randn(D, generator=g) * 0.05
    if extreme:                                             # hit the clamp and the pi-1e-4 branch
        x[0] = head.weight[y[0]].detach() * 7.0             # cos == 1 exactly-ish
        x[1] = -head.weight[y[1]].detach() * 2.0            # cos == -1: theta+m >= pi-1e-4
    if bf16_inputs:
        x = x.bfloat16().float()
        with torch.no_grad():
            head.weight.copy_(head.weight.bfloat16().float())
    x.requires_grad_(True)
    crit = nn.CrossEntropyLoss(label_smoothing=ls)
    out = head(x, y)
    loss = crit(out, y)
    loss.backward()
    np.savez(os.path.join(OUT, f"head_{name}.npz"),
             x=x.detach().numpy(), w=head.weight.detach().numpy(), y=y.numpy(),
             logits=out.detach().numpy(), loss=np.float32(loss.item()),
             dx=x.grad.numpy(), dw=head.weight.grad.numpy(),
             cos_max=np.float32(head.max_cos_theta), cos_min=np.float32(head.min_cos_theta),
             cfg=np.array([s, m, float(easy), float(warm_up_epochs), float(epoch),
                           float(training), ls, head.margin_factor, head.scale_factor],
                          dtype=np.float64))
    print(f"head_{name}: loss={loss.item():.6f} mf={head.margin_factor} sf={head.scale_factor}")


def hook_case(fm):
    """Two training steps of the real ArcFaceNet; the hook is registered after the first
    forward (face_models.py:569-570) so only step 2 is renormalised."""
    torch.manual_seed(7)
    C, B = 36, 8
    net = fm.ArcFaceNet(num_classes=C, dropout_rate=0.2, s=32.0, m=0.5)
    net.train()
    cap = {}

    def fwd_hook(mod, inp, out):
        inp[0].retain_grad()
        cap["emb"], cap["y"] = inp[0], inp[1]
    net.arcface.register_forward_hook(fwd_hook)
    crit = nn.CrossEntropyLoss(label_smoothing=0.05)
    rec = {}
    for step in range(2):
        net.zero_grad()
        img = torch.randn(B, 3, 64, 64)
        y = torch.randint(0, C, (B,))
        out = net(img, y)
        loss = crit(out, y)
        loss.backward()
        rec[f"emb{step}"] = cap["emb"].detach().numpy().copy()
        rec[f"y{step}"] = y.numpy().copy()
        rec[f"demb{step}"] = cap["emb"].grad.numpy().copy()
        rec[f"dw{step}"] = net.arcface.weight.grad.numpy().copy()
        rec[f"loss{step}"] = np.float32(loss.item())
        rec[f"last_grad_norm{step}"] = np.float64(net.last_grad_norm)
    rec["w"] = net.arcface.weight.detach().numpy().copy()
    rec["meta"] = np.array([net.max_grad_norm, net.phase, net.current_epoch], dtype=np.float64)
    np.savez(os.path.join(OUT, "hook_arcfacenet.npz"), **rec)
    print("hook: grad norms", rec["last_grad_norm0"], rec["last_grad_norm1"])


def gallery_cases(app):
    with open(f"{REF}/face_references/face_references.pkl", "rb") as f:
        saved = pickle.load(f)
    names = [r["name"] for r in saved]
    embs = np.stack([r["embedding_numpy"].reshape(-1) for r in saved]).astype(np.float32)
    refs = [{"name": n, "embedding": torch.tensor(e).reshape(1, -1), "image": None}
            for n, e in zip(names, embs)]
    dist = np.zeros((7, 7), dtype=np.float64)
    for i in range(7):
        for j in range(7):
            dist[i, j] = F.pairwise_distance(refs[i]["embedding"], refs[j]["embedding"]).item()
    res_idx, res_dist, res_name = [], [], []
    for thr in (1.0, 1.3, 2.0):
        for i in range(7):
            # leave-one-out: drop the query itself so thresholds matter
            sub = refs[:i] + refs[i + 1:]
            n, d, k = app.compare_faces(refs[i]["embedding"], sub, thr)
            res_name.append(n); res_dist.append(d); res_idx.append(-1 if k is None else k)
    self_match = [app.compare_faces(refs[i]["embedding"], refs, 1.0) for i in range(7)]
    np.savez(os.path.join(OUT, "gallery_fixture.npz"),
             names=np.array(names), emb=embs, dist=dist,
             loo_name=np.array(res_name), loo_dist=np.array(res_dist, dtype=np.float64),
             loo_idx=np.array(res_idx, dtype=np.int64), loo_thr=np.array([1.0, 1.3, 2.0]),
             self_name=np.array([s[0] for s in self_match]),
             self_dist=np.array([s[1] for s in self_match], dtype=np.float64),
             self_idx=np.array([s[2] for s in self_match], dtype=np.int64))
    print("gallery fixture: diag", dist[0, 0], "offdiag min", dist[dist > 1e-3].min())

    # synthetic: cfg2 recipe scaled down, run through the reference loop
    g = torch.Generator().manual_seed(1234)
    N, Q, D = 200, 24, 512
    G = F.normalize(torch.randn(N, D, generator=g), dim=1)
    Qm = torch.empty(Q, D)
    for i in range(Q):
        if i % 2 == 0:
            tau = 0.5 + 2.0 * torch.rand(1, generator=g).item()
            j = int(torch.randint(0, N, (1,), generator=g))
            Qm[i] = F.normalize(G[j] + (tau / D ** 0.5) * torch.randn(D, generator=g), dim=0)
        else:
            Qm[i] = F.normalize(torch.randn(D, generator=g), dim=0)
    G[150] = G[17]                                           # planted tie: first index must win
    Qm[1] = G[17]
    refs = [{"name": f"id{j}", "embedding": G[j:j + 1].clone(), "image": None} for j in range(N)]
    out = [app.compare_faces(Qm[i:i + 1], refs, 1.0) for i in range(Q)]
    np.savez(os.path.join(OUT, "gallery_synth.npz"), q=Qm.numpy(), g=G.numpy(),
             name=np.array([o[0] for o in out]),
             dist=np.array([o[1] for o in out], dtype=np.float64),
             idx=np.array([-1 if o[2] is None else o[2] for o in out], dtype=np.int64),
             thresh=np.float64(1.0))
    print("gallery synth: accepted", sum(o[2] is not None for o in out), "of", Q)
    assert app.compare_faces(None, refs, 1.0) == ("Unknown", float("inf"), None)
    assert app.compare_faces(Qm[:1], [], 1.0) == ("Unknown", float("inf"), None)


def cosine_case():
    """hyperparameter_tuning.py:1039-1046,1076 executed verbatim on seeded tensors."""
    g = torch.Generator().manual_seed(99)
    emb = torch.randn(40, 512, generator=g)
    w = torch.randn(36, 512, generator=g) * 0.05
    s = 32.0
    normalized_embeddings = F.normalize(emb, p=2, dim=1)
    class_centers = F.normalize(w, p=2, dim=1)
    logits = torch.matmul(normalized_embeddings, class_centers.t()) * s
    best, pred = logits.max(1)
    np.savez(os.path.join(OUT, "cosine_match.npz"), emb=emb.numpy(), w=w.numpy(),
             s=np.float64(s), pred=pred.numpy(), best=best.numpy(), logits=logits.numpy())


def main():
    fm = load_face_models()
    head_case(fm, "epoch0", 8, 36, 512, 1, epoch=0)
    head_case(fm, "epoch5", 16, 36, 512, 2, epoch=5, planted=4)
    head_case(fm, "postwarm", 16, 50, 128, 3, epoch=12, planted=4)
    head_case(fm, "easy", 16, 50, 128, 4, epoch=12, easy=True, planted=4)
    head_case(fm, "eval", 8, 36, 64, 5, epoch=3, training=False)
    head_case(fm, "ls0", 8, 36, 64, 6, epoch=10, ls=0.0, planted=2)
    head_case(fm, "ls15_m03", 8, 36, 64, 7, epoch=10, ls=0.15, m=0.3, s=16.0, planted=2)
    head_case(fm, "bf16in", 32, 100, 128, 8, epoch=10, planted=8, bf16_inputs=True)
    head_case(fm, "extreme", 8, 20, 64, 9, epoch=10, extreme=True)
    head_case(fm, "extreme_easy", 8, 20, 64, 10, epoch=10, extreme=True, easy=True)
    head_case(fm, "warm20", 8, 36, 64, 11, epoch=7, warm_up_epochs=20)
    hook_case(fm)
    gallery_cases(load_app())
    cosine_case()


if __name__ == "__main__":
    main()
