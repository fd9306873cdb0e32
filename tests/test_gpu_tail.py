"""The fused embedding tail: BatchNorm1d -> dropout -> row L2 norm (+ the head's K1 operands) as one
kernel each way (b200f_bn_stats / b200f_tail_fwd / b200f_tail_bwd) against the reference's own op sequence
(src/face_models.py:516-525: self.bn(x), self.dropout(x), F.normalize(x, p=2, dim=1, eps=1e-12)) evaluated by torch on the CPU
in fp64, train and eval mode, forward, backward and the running-statistics update."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _reference(z, bn, mask, p, training, R):
    """torch, CPU, fp64: bn -> dropout with the GIVEN keep mask (x * mask / (1 - p)) -> normalize; loss = <emb, R>."""
    bn = copy.deepcopy(bn).double().cpu().train(training)
    zz = z.detach().double().cpu().requires_grad_(True)
    y = bn(zz)
    if training and p > 0:
        y = y * mask.double().cpu() / (1.0 - p)
    emb = F.normalize(y, p=2, dim=1, eps=1e-12)
    (emb * R.double().cpu()).sum().backward()
    return dict(y=y.detach(), emb=emb.detach(), dz=zz.grad, dgamma=bn.weight.grad, dbeta=bn.bias.grad,
                rm=bn.running_mean.clone(), rv=bn.running_var.clone(), nbt=int(bn.num_batches_tracked))


@pytest.mark.parametrize("B,D", [(32, 512), (257, 512), (64, 136), (2, 8)])
@pytest.mark.parametrize("training,p", [(True, 0.2), (True, 0.0), (False, 0.2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_tail_vs_reference_chain(cuda_device, B, D, training, p, dtype):
    from b200face.head import fused_tail
    g = torch.Generator().manual_seed(B * 7 + D)
    z = (torch.randn(B, D, generator=g) * 2.0 + 0.5).to(dtype)
    bn = torch.nn.BatchNorm1d(D, eps=1e-5)
    with torch.no_grad():
        bn.weight.copy_(1.0 + 0.3 * torch.randn(D, generator=g)); bn.bias.copy_(0.2 * torch.randn(D, generator=g))
        bn.running_mean.copy_(0.3 * torch.randn(D, generator=g)); bn.running_var.copy_(0.5 + torch.rand(D, generator=g))
    mask = (torch.rand(B, D, generator=g) >= p).to(torch.uint8)
    R = torch.randn(B, D, generator=g)
    ref = _reference(z.float(), bn, mask, p, training, R)
    bn_d = copy.deepcopy(bn).to(cuda_device).train(training)
    zd = z.to(cuda_device).requires_grad_(True)
    y, side = fused_tail(zd, bn_d, p, training, mask=mask.to(cuda_device), want_operands=(D % 8 == 0), want_emb=True)
    (F.normalize(y, p=2, dim=1, eps=1e-12) * R.to(cuda_device)).sum().backward()
    tol = 2e-6
    assert rel_err(y.detach().cpu().numpy(), ref["y"].numpy()) < tol
    assert rel_err(side["emb"].cpu().numpy(), ref["emb"].numpy()) < tol
    np.testing.assert_allclose(side["inv_norm"].cpu().double().numpy(), 1.0 / ref["y"].norm(dim=1).clamp_min(1e-12).numpy(), rtol=2e-6)
    if D % 8 == 0:
        xo, inv = side["x_operands"]
        assert xo.dtype == torch.float16 and inv is side["inv_norm"]
        np.testing.assert_allclose(xo.float().cpu().numpy() / 256.0, ref["emb"].numpy(), atol=2 ** -11, rtol=0)
    gt = 5e-6 if dtype == torch.float32 else 2 ** -8          # z.grad comes back in z's dtype
    if B < 8:
        gt = max(gt, 1e-2)        # two rows: xbn = +-1 and dz is the residue of a near-total cancellation (|dz| ~ 1e-5 of its terms);
                                  # fp32 against the fp64 oracle measured 1.9e-3 on B200
    assert rel_err(zd.grad.float().cpu().numpy(), ref["dz"].numpy()) < gt
    assert rel_err(bn_d.weight.grad.cpu().numpy(), ref["dgamma"].numpy()) < 5e-6
    assert rel_err(bn_d.bias.grad.cpu().numpy(), ref["dbeta"].numpy()) < 5e-6
    np.testing.assert_allclose(bn_d.running_mean.cpu().double().numpy(), ref["rm"].numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(bn_d.running_var.cpu().double().numpy(), ref["rv"].numpy(), rtol=2e-6, atol=1e-7)
    assert int(bn_d.num_batches_tracked) == ref["nbt"]


def test_fused_tail_needs_two_rows_in_training(cuda_device):
    from b200face.head import fused_tail
    bn = torch.nn.BatchNorm1d(64).to(cuda_device).train()
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        fused_tail(torch.randn(1, 64, device=cuda_device), bn, 0.0, True)


@pytest.mark.parametrize("compute_dtype", [None, torch.bfloat16])
def test_arcfacenet_fused_tail_equals_module_chain(cuda_device, compute_dtype):
    """ArcFaceNet.forward_loss through the fused tail == through the torch modules (bn, dropout p = 0, normalise), loss and
    every gradient; compute_dtype = bf16 runs the head on the tcgen05 engine from the operands the tail emitted."""
    import b200face
    torch.manual_seed(3)
    net = b200face.ArcFaceNet(num_classes=40, dropout_rate=0.0).to(cuda_device).train()
    net.arcface.compute_dtype = compute_dtype
    img = torch.randn(16, 3, 64, 64, device=cuda_device)
    y = torch.randint(0, 40, (16,), device=cuda_device)
    res = {}
    for fused in (True, False):
        n2 = copy.deepcopy(net); n2.fused_tail = fused
        n2.zero_grad()
        loss = n2.forward_loss(img, y)
        loss.backward()
        res[fused] = (float(loss), n2.embedding.weight.grad.clone(), n2.bn.weight.grad.clone(), n2.bn.bias.grad.clone(),
                      n2.arcface.weight.grad.clone(), n2.bn.running_mean.clone(), n2.bn.running_var.clone())
    tol = 2e-5 if compute_dtype is None else 2e-3       # bf16 path: the unfused arm rounds y to bf16 before K1, the fused one does not
    assert res[True][0] == pytest.approx(res[False][0], rel=tol)
    for a, b in zip(res[True][1:5], res[False][1:5]):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < (tol if compute_dtype is None else 5e-3)
    for a, b in zip(res[True][5:], res[False][5:]):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    net.eval()
    with torch.no_grad():
        e_f = net.get_embedding(img)
        net.fused_tail = False
        e_u = net.get_embedding(img)
    assert torch.allclose(e_f, e_u, atol=2e-6)
