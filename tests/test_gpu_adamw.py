"""K5 (b200f_head_adamw through b200face.HeadAdamW): the fused AdamW / AMSGrad step of the class weights against the
oracle and against torch.optim.AdamW on CPU, plus the fused K1 outputs and the graphed training step."""
import numpy as np
import pytest
import torch

from oracle import adamw_oracle

pytestmark = pytest.mark.gpu
TOL = 2e-6          # fp32 arithmetic, same operation order: a few ulps


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("C,D", [(1000, 512), (257, 64), (300, 136), (64, 1024)])
@pytest.mark.parametrize("amsgrad", [True, False])
def test_adamw_vs_oracle_and_torch(cuda_device, C, D, amsgrad):
    import b200face
    g = torch.Generator().manual_seed(C + D)
    w0 = torch.randn(C, D, generator=g) * 0.05
    wd, lr = 1e-4, 2e-3
    wg = w0.clone().to(cuda_device)
    opt = b200face.HeadAdamW(wg, lr=lr, weight_decay=wd, amsgrad=amsgrad)
    p = torch.nn.Parameter(w0.clone())
    ref = torch.optim.AdamW([p], lr=lr, weight_decay=wd, amsgrad=amsgrad)
    w = w0.numpy().copy(); m = np.zeros_like(w); v = np.zeros_like(w); vmax = np.zeros_like(w) if amsgrad else None
    for step in range(1, 6):
        grad = torch.randn(C, D, generator=g) * (0.02 if step % 2 else 1.5)
        opt.step(grad.to(cuda_device))
        p.grad = grad.clone(); ref.step()
        w, m, v, vmax = adamw_oracle.adamw_step(w, grad.numpy(), m, v, vmax, step, lr=lr, weight_decay=wd)
    torch.cuda.synchronize()
    assert _rel(wg.cpu().numpy(), w) < TOL and _rel(wg.cpu().numpy(), p.detach().numpy()) < TOL
    assert _rel(opt.exp_avg.cpu().numpy(), m) < TOL
    assert _rel(opt.exp_avg_sq.cpu().numpy(), v) < TOL
    if amsgrad:
        assert _rel(opt.max_exp_avg_sq.cpu().numpy(), vmax) < TOL
    # the fused K1 outputs == K1 over the updated weights (row-sum order differs: an fp16 ulp here and there)
    w_hat, inv = b200face.head._k1(wg, True)
    assert _rel(opt.inv_norm.cpu().numpy(), inv.cpu().numpy()) < 1e-6
    d = (opt.w_hat.float() - w_hat.float()).abs().cpu().numpy()
    assert d.max() <= 0.25 and (d > 0).mean() < 1e-2          # values are w_hat * 256: one fp16 ulp is <= 0.125


def test_adamw_grad_scale_and_state_dict(cuda_device):
    import b200face
    g = torch.Generator().manual_seed(5)
    w0 = torch.randn(200, 512, generator=g) * 0.05
    grad = torch.randn(200, 512, generator=g)
    a = b200face.HeadAdamW(w0.clone().to(cuda_device), lr=1e-3)
    b = b200face.HeadAdamW(w0.clone().to(cuda_device), lr=1e-3)
    coef = torch.tensor([0.25], device=cuda_device)
    a.step(grad.to(cuda_device), grad_scale=coef)
    b.step((grad * 0.25).to(cuda_device))
    assert torch.equal(a.weight, b.weight) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
    # torch.optim.AdamW accepts our state and continues identically
    p = torch.nn.Parameter(a.weight.detach().cpu().clone())
    ref = torch.optim.AdamW([p], lr=1e-3, amsgrad=True)
    sd = a.state_dict()
    sd["state"][0] = {k: (t.cpu() if torch.is_tensor(t) else t) for k, t in sd["state"][0].items()}
    ref.load_state_dict(sd)
    g2 = torch.randn(200, 512, generator=g)
    p.grad = g2.clone(); ref.step()
    a.step(g2.to(cuda_device))
    assert _rel(a.weight.cpu().numpy(), p.detach().numpy()) < TOL


def test_graphed_training_step_with_fused_optimizer(cuda_device):
    """graphed_step(optimizer=...) (no K1 over W inside the graph) + HeadAdamW.step() == eager forward_loss/backward +
    K1 + the same optimizer, over three steps."""
    import b200face
    B, C, D = 128, 3000, 512
    g = torch.Generator().manual_seed(9)
    w0 = (torch.randn(C, D, generator=g) * 0.03)
    xs = [torch.randn(B, D, generator=g).bfloat16().to(cuda_device) for _ in range(3)]
    ys = [torch.randint(0, C, (B,), generator=g).to(cuda_device) for _ in range(3)]

    def make():
        h = b200face.ArcMarginProduct(D, C).to(cuda_device)
        with torch.no_grad():
            h.weight.copy_(w0)
        h.train(); h.update_epoch(12)
        return h
    h1, h2 = make(), make()
    o1 = b200face.HeadAdamW(h1, lr=1e-2, weight_decay=1e-4)
    o2 = b200face.HeadAdamW(h2, lr=1e-2, weight_decay=1e-4)
    step = h1.graphed_step(B, 0.05, torch.bfloat16, optimizer=o1)
    h2.cache_weight_prep = False                              # eager side: a real K1 over W every step
    for x, y in zip(xs, ys):
        l1 = step(x, y)
        o1.step()
        h2.zero_grad(set_to_none=True)
        l2 = h2.forward_loss(x.clone().requires_grad_(True), y, 0.05)
        l2.backward()
        o2.step()
        torch.cuda.synchronize()
        assert float(l1) == pytest.approx(float(l2), rel=2e-5)
    assert _rel(h1.weight.detach().cpu().numpy(), h2.weight.detach().cpu().numpy()) < 1e-4
