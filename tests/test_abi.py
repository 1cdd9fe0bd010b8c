"""The C-ABI library loads and exports every symbol include/relagn_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

from tests.util import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "relagn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from relagn_b200 import build as B, _capi
    so = B.build()
    lib = ctypes.CDLL(so)
    decl = declared_symbols()
    assert len(decl) >= 15
    for sym in decl:
        assert hasattr(lib, sym), sym
    assert sorted(_capi.SYMBOLS) == decl
    for sym in _capi.XSPEC_SYMBOLS:  # the gfortran names XSPEC resolves (lmod_relagn.dat, lmod_relqso.dat)
        assert hasattr(lib, sym), sym
    assert lib.rg_version() == 100


def test_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from relagn_b200 import _capi
    with pytest.raises(_capi.RelagnError):
        _capi.Context(0)
    from relagn_b200.batch import BatchEvaluator
    import numpy as np
    with pytest.raises(_capi.RelagnError):
        BatchEvaluator(np.geomspace(1e-4, 1e3, 1001))
