"""ctypes binding of libammc_b200_debug.so: the product library plus the hardware probes of include/ammc_b200_debug.h.

Built on demand with `python -m ammcnet_aaai2021_b200.build --debug`; only tools/ use it, the package never does."""
from __future__ import annotations

import ctypes
import os

from . import _capi
from ._capi import I, P

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libammc_b200_debug.so")
SIGNATURES = {
    "ammc_debug_mma_rate": (I, [P, I, I, I, P]),
    "ammc_debug_fp8_probe": (I, [P, P, P, P, P, I, P]),
    "ammc_debug_desc_probe": (I, [P, P, P, I, I, I, P]),
    "ammc_debug_tma_probe": (I, [P, P, P, P, P, P, I, P]),
}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: build it with `python -m ammcnet_aaai2021_b200.build --debug`" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(SIGNATURES.items()) + [("ammc_last_error", (ctypes.c_char_p, []))]:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def call(name: str, *args):
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError("ammc_b200 debug %s failed (code %d): %s" % (name, rc, (load().ammc_last_error() or b"").decode()))
