"""Drop-in for the one PyXspec entry point the reference uses: `xspec.callModelFunction('kyconv', energies, params,
flux)` (relagn.py:1000-1003, 1097-1099, 1355-1357), backed by rg_kyconv (seam B1 of include/relagn_b200.h).

    import relagn_b200.xspec_shim as xspec      # instead of `import xspec`
    xspec.callModelFunction('kyconv', Ebins, kyparams, phs)        # mutates `phs` in place, like PyXspec

`bind(ctx)` selects the library context (one per device) whose resident Kerr tables are used.
"""
import numpy as np

_ctx = None


def bind(ctx):
    global _ctx
    _ctx = ctx


def _context():
    global _ctx
    if _ctx is None:
        from . import _capi, tables
        _ctx = _capi.Context(-1)
        tables.load_default(_ctx)
    return _ctx


def callModelFunction(name, energies, params, flux, *rest):
    if name != 'kyconv':
        raise NotImplementedError("relagn_b200.xspec_shim only provides 'kyconv' (the model the reference calls)")
    out = _context().kyconv(np.asarray(energies, dtype=np.float64), np.asarray(params, dtype=np.float64),
                            np.array(flux, dtype=np.float64))
    flux[:] = out.tolist() if isinstance(flux, list) else out
