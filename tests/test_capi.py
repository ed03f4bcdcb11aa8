"""CPU: the C-ABI library builds, loads, and exports exactly what include/ammc_b200.h declares."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "ammc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ammc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ammcnet_aaai2021_b200 import _capi, build
    build.build(verbose=False)
    lib = _capi.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "library does not export " + n
    assert sorted(_capi.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert lib.ammc_version() == 100


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: CPU tensors are rejected before any native call."""
    import torch
    import ammcnet_aaai2021_b200 as A
    m = A.enc_quan_dec_res_topk(64, 16, 8, k=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 64, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        A.psnr_error(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8))
    q = A.Quantize_topk(16, 8, k=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        q(torch.zeros(1, 4, 4, 16))
    # the widened rows (generator engine, preprocessing) follow the same rule
    with pytest.raises(RuntimeError, match="CUDA"):
        A.GeneratorEngine(A.get_twostream().eval())(torch.zeros(1, 12, 64, 64), torch.zeros(1, 6, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA uint8"):
        A.preprocess_frames(torch.zeros(1, 8, 8, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="CUDA float32"):
        A.preprocess_flow(torch.zeros(1, 8, 8, 2))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from ammcnet_aaai2021_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.load()
