import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a, b):
    """max |a-b| / max(|b|) -- norm-wise relative error used for reporting."""
    import torch
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, rtol=1e-3, what=""):
    """|a - b| <= rtol * max(|b|, rms(b)) elementwise: 1e-3 relative, with the tensor's own RMS as the floor for
    elements that are (near) zero -- the tolerance stated in BASELINE.json for features, losses and scores."""
    import torch
    a = torch.as_tensor(np.asarray(a.detach().cpu() if hasattr(a, "detach") else a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b.detach().cpu() if hasattr(b, "detach") else b), dtype=torch.float64)
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if b.numel() == 0:
        return
    assert torch.isfinite(a).all(), f"{what}: non-finite values"
    rms = b.pow(2).mean().sqrt()
    tol = rtol * torch.maximum(b.abs(), rms.expand_as(b))
    bad = (a - b).abs() > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{b.numel()} elements outside {rtol} relative; "
                           f"max abs err {float((a - b).abs().max()):.3e}, rms(ref) {float(rms):.3e}")
