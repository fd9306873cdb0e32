"""One rank of the multi-GPU parity check (launched by tests/test_gpu_multirank.py through torch.distributed.run, or by
hand: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_rank_main.py).

Every rank builds the SAME full problem from a global seed, takes its class shard, and runs the class-parallel head
(parallel.ShardedArcMarginProduct: one NCCL all-reduce of [B,4] forward, one of [B,D] backward) eagerly and through the
captured CUDA graph; rank 0 also evaluates the unsharded head on its own GPU and the CPU oracle, and every rank checks

  sharded loss / dx == unsharded loss / dx, its dW rows == the unsharded dW[lo:hi]   (same kernels, other partition)
  all of them == oracle.head_forward_backward to the bf16 bar
  gather_weight() / load_full_weight() round-trip the reference's [C, D] layout, strict load into an unsharded module
  the sharded gallery top-k (all-gather + merge) == the unsharded top-k
and then tears the process group down normally (the graph is closed first).  Prints 'RANK r OK'.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def main():
    import b200face
    from b200face import parallel
    import oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, C, D = 384, 20_000, 512
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(B, D, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    head = parallel.ShardedArcMarginProduct(D, C, seed=7).to(dev)
    head.update_epoch(12); head.train()
    lo, hi = head.lo, head.hi
    # the init is the shard of ONE full-matrix draw, whatever the world size
    w_full = parallel.full_matrix_init_rows(0, C, C, D, seed=7)
    assert torch.equal(head.weight.detach().cpu(), w_full[lo:hi])
    wb = w_full.bfloat16().float()                              # bf16-representable weights: the oracle's inputs
    x[:48] = 3.0 * wb[y[:48]] + wb.std() * torch.randn(48, D, generator=g)
    xb = x.bfloat16()
    head.load_full_weight(wb)
    assert torch.equal(head.gather_weight().cpu(), wb)          # reference layout back, bit for bit
    xg = xb.to(dev).requires_grad_(True)
    loss, pred = head.forward_loss(xg, y.to(dev), 0.05, return_pred=True)
    loss.backward()
    torch.cuda.synchronize()
    l_sh, dx_sh, dw_sh = float(loss), head.local.last_stats.dx_f32.clone(), head.weight.grad.clone()
    # graphed sharded step == eager sharded step, bit for bit
    step = head.graphed_step(B, 0.05, torch.bfloat16)
    lg = step(xb.to(dev), y.to(dev))
    torch.cuda.synchronize()
    assert float(lg) == l_sh, (float(lg), l_sh)
    assert torch.equal(head.weight.grad, dw_sh)
    # unsharded evaluation of the same problem on this rank's own GPU
    full = b200face.ArcMarginProduct(D, C).to(dev)
    full.update_epoch(12); full.train()
    full.load_state_dict(head.full_state_dict(), strict=True)   # the reference's checkpoint layout, strict
    x1 = xb.to(dev).requires_grad_(True)
    l1, pred1 = full.forward_loss(x1, y.to(dev), 0.05, return_pred=True)
    l1.backward()
    torch.cuda.synchronize()
    e = dict(loss=abs(l_sh - float(l1)) / abs(float(l1)), dx=rel(dx_sh.cpu(), full.last_stats.dx_f32.cpu()),
             dw=rel(dw_sh.cpu(), full.weight.grad[lo:hi].cpu()))
    assert e["loss"] < 2e-6 and e["dx"] < 1e-5 and e["dw"] < 1e-5, e
    assert float((pred == pred1).float().mean()) > 0.995
    ref = oracle.head_forward_backward(xb.float().numpy(), wb.numpy(), y.numpy(),
                                       oracle.HeadConfig(current_epoch=12, training=True, label_smoothing=0.05))
    eo = dict(loss=abs(l_sh - float(ref["loss"])) / abs(float(ref["loss"])), dx=rel(dx_sh.cpu(), ref["dx"]),
              dw=rel(dw_sh.cpu(), ref["dw"][lo:hi]))
    assert eo["loss"] < 1e-3 and eo["dx"] < 1e-3 and eo["dw"] < 1e-3, eo
    # gallery: rows sharded, per-shard top-k, all-gather, merge == unsharded top-k
    N, Q, k = 30_000, 200, 5
    G = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    Qm = torch.nn.functional.normalize(G[:Q] + 0.05 * torch.randn(Q, D, generator=g), dim=1)
    G[150] = G[17]
    glo, ghi = parallel.shard_bounds(N, world, rank)
    idx, score, acc = parallel.sharded_gallery_topk(Qm.to(dev), G[glo:ghi].to(dev).contiguous(), k, 1.0, "l2eps", index_offset=glo)
    i1, s1, a1 = b200face.gallery_topk(Qm.to(dev), G.to(dev), k, 1.0, "l2eps")
    assert torch.equal(idx, i1) and torch.equal(acc, a1) and torch.allclose(score, s1, rtol=1e-6)
    head.close()                                                # graphs hold NCCL kernels: release before teardown
    dist.barrier()
    dist.destroy_process_group()
    print(f"RANK {rank} OK sharded-vs-unsharded {e} sharded-vs-oracle {eo}", flush=True)


if __name__ == "__main__":
    main()
