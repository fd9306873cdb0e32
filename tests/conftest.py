import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def head_golden_names():
    return sorted(os.path.basename(p)[len("head_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN, "head_*.npz")))


def cfg_from_golden(d):
    """HeadConfig as the reference module was configured when the fixture was made
    (tests/golden/make_golden.py:head_case)."""
    from oracle import HeadConfig
    s, m, easy, wue, epoch, training, ls, _mf, _sf = [float(v) for v in d["cfg"]]
    return HeadConfig(s=s, m=m, easy_margin=bool(easy), use_warm_up=True,
                      warm_up_epochs=int(wue), current_epoch=int(epoch),
                      training=bool(training), label_smoothing=ls)


def rel_err(a, b):
    """Norm-wise relative error ||a-b||_F / ||b||_F (the metric every tolerance in this
    suite is stated in)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
