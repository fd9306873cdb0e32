"""Pin oracle/torch_port.py (the CPU arm bench.py times) against the reference-generated golden
vectors.  CPU only."""
import numpy as np
import pytest
import torch

from conftest import cfg_from_golden, golden, head_golden_names, rel_err
from oracle import torch_port


@pytest.mark.parametrize("name", head_golden_names())
def test_head_port_matches_reference(name):
    d = golden(f"head_{name}.npz")
    cfg = cfg_from_golden(d)
    head = torch_port.HeadPort(d["w"].shape[1], d["w"].shape[0], s=cfg.s, m=cfg.m, easy_margin=cfg.easy_margin)
    head.warm_up_epochs, head.current_epoch = cfg.warm_up_epochs, cfg.current_epoch
    head.train(cfg.training)
    with torch.no_grad():
        head.weight.copy_(torch.tensor(d["w"]))
    loss, dx, dw = torch_port.head_step(head, torch.tensor(d["x"]), torch.tensor(d["y"]), cfg.label_smoothing)
    # same torch build, same ops: agreement is at rounding level
    assert float(loss) == pytest.approx(float(d["loss"]), rel=1e-6)
    assert rel_err(dx.numpy(), d["dx"]) < 1e-5
    assert rel_err(dw.numpy(), d["dw"]) < 1e-5
    assert head.max_cos_theta == pytest.approx(float(d["cos_max"]), abs=1e-7)


def test_gallery_port_matches_reference():
    d = golden("gallery_synth.npz")
    refs = [{"name": f"id{j}", "embedding": torch.tensor(d["g"][j:j + 1])} for j in range(d["g"].shape[0])]
    for i in range(d["q"].shape[0]):
        name, dist, idx = torch_port.compare_faces_loop(torch.tensor(d["q"][i:i + 1]), refs, float(d["thresh"]))
        assert name == str(d["name"][i]) and (-1 if idx is None else idx) == int(d["idx"][i])
        assert dist == pytest.approx(float(d["dist"][i]), rel=1e-6)
    idx, score, acc = torch_port.gallery_vectorised(torch.tensor(d["q"]), torch.tensor(d["g"]), 1, 1.0)
    ok = d["idx"] >= 0
    assert np.array_equal(acc.numpy(), ok)
    assert np.array_equal(idx.numpy()[ok, 0], d["idx"][ok])
