"""TEST-ONLY loader of the host simulation (see cuda_shim.h)."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from relagn_b200 import _capi  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        sys.path.insert(0, HERE)
        import build as hs_build
        _lib = _capi.bind(C.CDLL(hs_build.build()))
    return _lib


def context():
    return _capi.Context(0, lib=lib())
