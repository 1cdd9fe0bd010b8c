"""Kernel LOGIC checks without a GPU: the product's .cu sources compiled with g++ against the test-only CUDA shim
(tests/hostsim/cuda_shim.h; fibers emulate threads, barriers and warp shuffles) and driven through the same C ABI.
This is development infrastructure -- the shipped library is nvcc/sm_100a only and `-m gpu` holds the parity tests
proper -- but it pins the arithmetic of every kernel against the reference goldens on the CPU-only build box."""
import os
import sys

import numpy as np
import pytest

from tests.util import GOLDEN, peak_relerr, golden_tables, RELAGN_DEF

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "hostsim"))


@pytest.fixture(scope="module")
def ctx():
    import sim
    c = sim.context()
    t = golden_tables()
    c.tables_upload(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["dl"], t["data"])
    return c


def test_nonrel_and_rel_vs_reference_goldens(ctx):
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    dr = np.load(os.path.join(GOLDEN, "golden_rel.npz"))
    ctx.set_grid(dn["grid1000"])
    sel = [0, 2, 9]  # default, fcol<0, r_w > r_out clamp
    out, at = ctx.eval_batch_host(dn["relagn_params"][sel], model=0, rel=False, components=True)
    for k, i in enumerate(sel):
        for c in range(4):
            assert peak_relerr(out[k, c], dn["relagn_spec1000"][i][c]) < 1e-10
    out, at = ctx.eval_batch_host(dr["relagn_params"][[0, 1]], model=0, rel=True, components=True)
    for k in range(2):
        for c in range(4):
            assert peak_relerr(out[k, c], dr["relagn_spec1000"][k][c]) < 1e-10
    out, at = ctx.eval_batch_host(dr["relqso_params"][[0]], model=1, rel=True, components=True)
    for c in range(4):
        assert peak_relerr(out[0, c], dr["relqso_spec1000"][0][c]) < 1e-10
    from relagn_b200._capi import A
    assert at[0, A["r_h"]] == 14.261351436104999 and at[0, A["r_w"]] == 28.522702872209997


def test_python_default_grid(ctx):
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    ctx.set_grid(dn["grid_def"])   # 1999 bins -> the M = 8 instantiation
    out, _ = ctx.eval_batch_host([RELAGN_DEF], model=0, rel=False, components=True)
    for c in range(4):
        assert peak_relerr(out[0, c], dn["relagn_specdef"][0][c]) < 1e-10


def test_kyconv_seam(ctx):
    from oracle import oracle as O
    t = golden_tables()
    otab = O.Tables(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["dl"], data=t["data"])
    eb = np.geomspace(0.1, 2.0, 501)
    ec = 0.5 * (eb[1:] + eb[:-1])
    line = np.exp(-0.5 * ((ec - 1.0) / 0.01) ** 2) * np.diff(eb)
    par = [0.998, 30.0, 1.236971, 0, 100.0, 3.0, 3.0, 10.0, 0, 0, 500, -1]
    assert peak_relerr(ctx.kyconv(eb, par, line.copy()), O.kyconv(otab, eb, par, line)) < 1e-12
    with pytest.raises(Exception):
        ctx.kyconv(eb, [0.998, 30.0, 1.3, 1, 100.0, 3, 3, 10, 0, 0, 500, -1], line.copy())  # ms = 1 unsupported


def test_class_api_units(ctx, monkeypatch):
    from relagn_b200 import model
    cached = {"ctx": ctx, "tables": True}
    monkeypatch.setattr(model, "_context", lambda device=None: cached)
    u = np.load(os.path.join(GOLDEN, "golden_units.npz"))
    o = model.relagn()
    assert o.Ebins.size == 2000 and o.Egrid.size == 1999
    o.new_ear(np.geomspace(1e-4, 1e3, 1001))
    for unit in ["cgs", "cgs_wave", "SI", "SI_wave", "counts"]:
        for fl in (False, True):
            o.set_units(unit)
            o.set_flux() if fl else o.set_lum()
            assert peak_relerr(o.get_totSED(rel=False), u["%s_%d" % (unit, fl)]) < 1e-10
    assert o.get_Ledd() == pytest.approx(float(u["Ledd_cgs"]))
    with pytest.raises(ValueError, match="not physical"):
        model.relagn(a=0.9999)
    with pytest.raises(ValueError, match="Inclination out of bounds"):
        model.relagn(cos_inc=0.05)
    with pytest.raises(ValueError, match="mdot is out of bounds"):
        model.relqso(log_mdot=-2.0)
    q = model.relqso()   # examples/relqso_modSpec.ipynb:81-83
    assert q.r_h == 14.261351436104999 and q.r_w == 28.522702872209997
    assert q.gamma_h == pytest.approx(1.9803689563797076, rel=1e-12)
