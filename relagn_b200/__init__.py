"""relagn_b200 -- B200-native (sm_100a) evaluation of the RELAGN / RELQSO AGN SED models.

Drop-in for the spectrum path of scotthgn/RELAGN's Python classes (src/python_version/relagn.py): same class
names, constructor parameters, getters and physical attributes; the arithmetic runs in hand-written CUDA kernels
behind the C ABI of include/relagn_b200.h.  There is no CPU fallback.
"""
from ._capi import RelagnError  # noqa: F401

__all__ = ["relagn", "relqso", "BatchEvaluator", "RelagnError"]


def __getattr__(name):
    if name in ("relagn", "relqso"):
        from . import model
        return getattr(model, name)
    if name in ("BatchEvaluator", "evaluate_sharded", "shard_bounds"):
        from . import batch
        return getattr(batch, name)
    raise AttributeError(name)
