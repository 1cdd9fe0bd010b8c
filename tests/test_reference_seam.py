"""The UNMODIFIED reference (relagn.py) driving the product's kyconv seam: its `import xspec` is satisfied by
relagn_b200.xspec_shim bound to the test-only host simulation of the kernels, so do_rel*Spec runs end to end with
the product's rg_kyconv at relagn.py:1001,1098,1356.  Runs only where /root/reference exists (the build container)."""
import contextlib
import io
import os
import sys
import types

import numpy as np
import pytest

from tests.util import GOLDEN, peak_relerr, golden_tables, RELAGN_DEF

REF = "/root/reference/src/python_version"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "relagn.py")), reason="reference not mounted")


def test_unmodified_reference_through_rg_kyconv():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))
    sys.path.insert(0, GOLDEN)
    import sim
    import ref_stubs
    from relagn_b200 import xspec_shim
    ctx = sim.context()
    t = golden_tables()
    ctx.tables_upload(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], t["data"])
    xspec_shim.bind(ctx)
    saved = {k: sys.modules.get(k) for k in ("xspec", "astropy", "astropy.units", "astropy.constants", "relagn")}
    try:
        ref_stubs.install(lambda e, p, f: f)          # astropy stand-ins (and a placeholder xspec)
        sys.modules["xspec"] = xspec_shim             # ... replaced by the product's seam
        sys.path.insert(0, REF)
        sys.modules.pop("relagn", None)
        import relagn as ref
        d = np.load(os.path.join(GOLDEN, "golden_rel.npz"))
        keys = ["M", "dist", "log_mdot", "a", "cos_inc", "kTe_hot", "kTe_warm", "gamma_hot", "gamma_warm", "r_hot",
                "r_warm", "log_rout", "fcol", "h_max", "z"]
        with contextlib.redirect_stdout(io.StringIO()):
            o = ref.relagn(**dict(zip(keys, RELAGN_DEF)))
            o.new_ear(np.array(d["grid1000"]))
            o.set_units("SI")
            tot = o.get_totSED(rel=True)
        # 105 annuli, each through the product's kyconv: equals the golden made with the oracle's kyconv
        assert peak_relerr(tot, d["relagn_spec1000"][0][3]) < 1e-10
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
