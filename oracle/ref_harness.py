"""Import the UNMODIFIED reference (/root/reference) in-process, for fixture generation only.

TEST INFRASTRUCTURE -- not product code.  Only `oracle/gen_golden.py` and the container-only
the host-logic test that patches the live reference (`tests/test_host_logic.py`) use this file.  `/root/reference` does not exist on the
GPU box, so nothing under `-m gpu`, `smoke()` or `bench.py` may import it.

Recipe follows SURVEY.md Appendix B: the reference imports a handful of packages that are not
installed here (`torchsummaryX`, `tensorboardX`, `png`, `lmdb`, `turbojpeg`, `matplotlib`); none of
them is used by the hot path, so empty stub modules are enough.  The reference files themselves are
imported from where they lie; nothing is copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("AMMC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Code", "models"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    """Stub the third-party modules the reference imports but the hot path never calls."""
    _stub("torchsummaryX", summary=lambda *a, **k: None)

    class SummaryWriter:  # tensorboardX.SummaryWriter, ctor only
        def __init__(self, *a, **k):
            pass

    _stub("tensorboardX", SummaryWriter=SummaryWriter)
    _stub("png")
    _stub("lmdb")
    _stub("turbojpeg", TurboJPEG=object, TJPF_GRAY=0, TJSAMP_GRAY=0, TJFLAG_PROGRESSIVE=0)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.colors = _stub("matplotlib.colors")


def import_reference():
    """Returns (unet_module, utils_module, eval_metric_module) of the live reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torch
    state = torch.get_rng_state()          # Code/models/unet.py:4 reseeds the global RNG on import
    import Code.models.unet as ref_unet
    import Code.utils.utils as ref_utils
    import Code.main.eval_metric as ref_eval
    torch.set_rng_state(state)
    return ref_unet, ref_utils, ref_eval
