"""CPU PyTorch restatement of the reference hot path, for TIMING the reference's CPU arm
(TEST / BENCH INFRASTRUCTURE ONLY -- bench.py's cpu_baseline and --impl reference legs and the CPU
tests may import this; the product never does).

/root/reference is pure Python and does not exist on the GPU box, and it cannot be pip-installed
(no setup.py / pyproject; importing ``src`` needs facenet_pytorch and creates directories), so the
"reference arm" is this port: the same torch ops in the same order as
  ArcMarginProduct.forward        /root/reference/src/face_models.py:334-429
  nn.CrossEntropyLoss(label_smoothing) + loss.backward()   src/training.py:341,515-521
  compare_faces                   /root/reference/src/app.py:50-64
with autograd doing the backward exactly as it does for the reference.  Pinned against the
reference-generated golden vectors by tests/test_torch_port.py.  kind = "port".
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class HeadPort(torch.nn.Module):
    """State and forward of the reference head (ctor defaults face_models.py:307-332)."""

    def __init__(self, in_feats, out_feats, s=32.0, m=0.5, use_warm_up=True, easy_margin=False):
        super().__init__()
        self.s, self.m, self.easy_margin, self.use_warm_up = s, m, easy_margin, use_warm_up
        self.warm_up_epochs, self.margin_factor, self.scale_factor, self.current_epoch = 10, 0.0, 0.3, 0
        self.weight = torch.nn.Parameter(torch.empty(out_feats, in_feats))
        torch.nn.init.xavier_normal_(self.weight, gain=math.sqrt(2))
        self.max_cos_theta = self.min_cos_theta = 0.0

    def forward(self, inp, label):
        if self.training and self.use_warm_up:                               # :336-348
            if self.current_epoch < self.warm_up_epochs:
                p = self.current_epoch / self.warm_up_epochs
                self.margin_factor = min(0.9, p * p)
                self.scale_factor = min(0.8, 0.3 + 0.5 * p)
            else:
                self.margin_factor, self.scale_factor = 0.9, 0.8
        xn = F.normalize(inp, p=2, dim=1, eps=1e-12)                          # :351-352
        wn = F.normalize(self.weight, p=2, dim=1, eps=1e-12)
        cos = F.linear(xn, wn)                                                # :355
        with torch.no_grad():                                                 # :358-360 (host syncs)
            self.max_cos_theta = cos.max().item()
            self.min_cos_theta = cos.min().item()
        c = torch.clamp(cos, min=-1.0 + 1e-7, max=1.0 - 1e-7)                 # :363
        theta = torch.acos(c)                                                 # :366
        m_eff = self.m * self.margin_factor if self.training else self.m      # :369
        hot = torch.zeros_like(c)
        hot.scatter_(1, label.view(-1, 1), 1)
        if self.easy_margin:                                                  # :372-384
            phi = torch.where(c > 0, torch.cos(theta + m_eff), c)
        else:                                                                 # :385-397
            phi = torch.cos(torch.minimum(torch.tensor(math.pi - 1e-4), theta + m_eff))
        out = torch.where(hot.bool(), phi, c)
        s_eff = min(self.s, 24.0)                                             # :401-409
        s_eff = s_eff * min(0.8, self.scale_factor) if self.training else s_eff
        if self.m > 0.4 and self.training:
            s_eff = s_eff * (0.8 - 0.5 * self.margin_factor)
        out = out * s_eff                                                     # :412
        if torch.isnan(out).any() or torch.isinf(out).any():                  # :423-427
            out = torch.where(torch.isnan(out) | torch.isinf(out), torch.zeros_like(out), out)
        return out


def head_step(head: HeadPort, x: torch.Tensor, y: torch.Tensor, label_smoothing: float = 0.05):
    """One fwd+bwd of the head exactly as the training loop drives it (training.py:511-521)."""
    head.zero_grad(set_to_none=True)
    x = x.detach().requires_grad_(True)
    loss = torch.nn.CrossEntropyLoss(label_smoothing=label_smoothing)(head(x, y), y)
    loss.backward()
    return loss.detach(), x.grad, head.weight.grad


def compare_faces_loop(emb, refs, thresh):
    """The Python loop of app.py:50-64 (one F.pairwise_distance + .item() per reference)."""
    if emb is None or not refs:
        return "Unknown", float("inf"), None
    min_dist, best, best_i = float("inf"), "Unknown", None
    e = emb.cpu()
    for i, ref in enumerate(refs):
        d = F.pairwise_distance(e, ref["embedding"].cpu()).item()
        if d < min_dist:
            min_dist, best, best_i = d, ref["name"], i
    return (best, min_dist, best_i) if min_dist <= thresh else ("Unknown", min_dist, None)


def gallery_vectorised(q: torch.Tensor, g: torch.Tensor, k: int, thresh: float):
    """All-cores CPU stand-in named by the north star: torch.cdist + topk (no eps term; BASELINE.md
    notes max |cdist - exact| = 4e-6 with the same argmin on the cfg2 recipe)."""
    d = torch.cdist(q, g)
    score, idx = d.topk(k, dim=1, largest=False)
    return idx, score, score[:, 0] <= thresh


class ArcFaceNetPort(torch.nn.Module):
    """cfg1's model for the CPU timing arm: the reference ArcFaceNet (face_models.py:447-535) restated --
    torchvision ResNet18 trunk (random init: the ImageNet weights cannot be downloaded offline; same FLOPs),
    Linear(512,512,no bias) -> BatchNorm1d -> dropout(0.2, train) -> F.normalize -> ArcMarginProduct (HeadPort).
    The backward hook (:538-570) rescales a [B,36] gradient: no measurable cost, left out of the timing port."""

    def __init__(self, num_classes=36, dropout_rate=0.2, s=32.0, m=0.5):
        super().__init__()
        import torchvision.models as models
        backbone = models.resnet18(weights=None)
        self.features = torch.nn.Sequential(*list(backbone.children())[:-1])
        self.embedding = torch.nn.Linear(512, 512, bias=False)
        self.bn = torch.nn.BatchNorm1d(512, eps=1e-5)
        self.dropout = torch.nn.Dropout(p=dropout_rate)
        self.arcface = HeadPort(512, num_classes, s=s, m=m)
        self.current_epoch = 0

    def forward(self, x, labels):
        x = self.features(x).view(x.size(0), -1)                              # :511-512
        x = self.bn(self.embedding(x))                                        # :515-516
        if self.training:
            x = self.dropout(x)                                               # :519-520
        emb = F.normalize(x, p=2, dim=1, eps=1e-12)                           # :524
        self.arcface.current_epoch = self.current_epoch                       # :531
        return self.arcface(emb, labels)                                      # :534


def arcfacenet_train_step(model: ArcFaceNetPort, opt, x, y, label_smoothing=0.05):
    """One full training step as src/training.py:505-546 drives it (zero_grad, forward, CE, backward, AdamW step)."""
    opt.zero_grad()
    loss = torch.nn.CrossEntropyLoss(label_smoothing=label_smoothing)(model(x, y), y)
    loss.backward()
    opt.step()
    return loss.detach()
