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
    c.tables_upload(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], t["data"])
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


def test_host_pipeline_pieces(ctx, monkeypatch):
    """rg_eval_batch_host: the spectra do not depend on how the spectrum kernel is cut into pieces (and every row and
    attribute reaches the host): 12 sets as one piece, as pieces of 5 (5 + 7: the short remainder joins) and of 3."""
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    ctx.set_grid(dn["grid1000"])
    p = np.concatenate([dn["relagn_params"][:6], dn["relagn_params"][:6][::-1]])  # 12 sets >= 2 x 4 simulated SMs
    ref, at_ref = ctx.eval_batch_host(p, model=0, rel=True)
    assert np.isfinite(ref).all() and (ref.max(axis=1) > 0).all()
    for piece in ("5", "3"):
        monkeypatch.setenv("RG_PIECE_SETS", piece)
        out, at = ctx.eval_batch_host(p, model=0, rel=True)
        np.testing.assert_array_equal(out, ref)
        np.testing.assert_array_equal(at, at_ref)


def test_python_default_grid(ctx):
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    ctx.set_grid(dn["grid_def"])   # 1999 bins -> the M = 8 instantiation
    out, _ = ctx.eval_batch_host([RELAGN_DEF], model=0, rel=False, components=True)
    for c in range(4):
        assert peak_relerr(out[0, c], dn["relagn_specdef"][0][c]) < 1e-10


def test_kyconv_seam(ctx):
    from oracle import oracle as O
    t = golden_tables()
    otab = O.Tables(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], data=t["data"])
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


def test_observation_epilogue_vs_oracle(ctx):
    """SURVEY 8 f1/f2 logic on the CPU-only box: unit conversion (relagn.py:487-564), the XSPEC-contract rebin
    (relagn.f:98-112) and the folded fit statistic against oracle/observe_oracle.c; the conversion also against
    the unmodified reference's getter outputs."""
    from oracle import oracle as O
    from tests.util import toy_response
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    u = np.load(os.path.join(GOLDEN, "golden_units.npz"))
    eb = dn["grid1000"]
    nE = eb.size - 1
    ctx.set_grid(eb)
    sel = [0, 6, 8, 15]  # default, z = 0.5, face-on, a C4 draw
    P = len(sel)
    par = np.ascontiguousarray(dn["relagn_params"][sel])
    L = np.ascontiguousarray(dn["relagn_spec1000"][sel][:, 3, :])
    # ---- f1: units x flux, bit-identical to the oracle; the default set against the reference's own getters ----
    for unit in ["SI", "cgs", "cgs_wave", "SI_wave", "counts"]:
        for fl in (False, True):
            out = np.empty_like(L)
            ctx.convert_device(L.ctypes.data, par.ctypes.data, P, 1, unit, fl, out.ctypes.data)
            for k in range(P):
                assert np.array_equal(out[k], O.convert(L[k], eb, unit, fl, par[k, 1], par[k, 14])), (unit, fl, k)
            assert peak_relerr(out[0], u["%s_%d" % (unit, fl)]) < 1e-14
    # components layout [P][4][nE], in place
    L4 = np.ascontiguousarray(dn["relagn_spec1000"][sel])
    buf = L4.copy()
    ctx.convert_device(buf.ctypes.data, par.ctypes.data, P, 4, "counts", True, buf.ctypes.data)
    assert np.array_equal(buf[1, 0], O.convert(L4[1, 0], eb, "counts", True, par[1, 1], par[1, 14]))
    assert peak_relerr(buf[1, 3], u["z05_counts_1"]) < 1e-14

    # ---- f1: rebin onto instrument grids ----
    grids = [np.linspace(0.3, 10.0, 257),                 # linear channels inside the model grid
             np.geomspace(5e-5, 3e3, 65),                 # wider than the model grid on both sides: edge bins clipped
             np.array([2e3, 3e3, 4e3]),                   # entirely outside: zeros
             eb[::7].copy()]                              # every 7th model edge: whole-bin sums (z = 0 sets)
    for ear in grids:
        ctx.set_instrument(ear)
        ph = np.empty((P, ear.size - 1))
        ctx.photar_device(L.ctypes.data, par.ctypes.data, P, ph.ctypes.data)
        for k in range(P):
            assert np.array_equal(ph[k], O.photar(L[k], eb, par[k, 1], par[k, 14], ear)), k
    assert np.all(ph[[0, 2, 3]] > 0)
    cnt = np.stack([O.convert(L[k], eb, "counts", True, par[k, 1], par[k, 14]) for k in range(P)]) * np.diff(eb)
    k0 = 0  # z = 0: grid of every 7th edge -> sums of 7 model bins
    ref = np.add.reduceat(cnt[k0][:(nE // 7) * 7], np.arange(0, (nE // 7) * 7, 7))
    assert np.allclose(ph[k0][:nE // 7], ref, rtol=1e-13, atol=0)
    ctx.set_instrument(np.array([2e3, 3e3, 4e3]))
    z = np.empty((P, 2))
    ctx.photar_device(L.ctypes.data, par.ctypes.data, P, z.ctypes.data)
    assert np.all(z == 0)
    # identity: the observer-frame image of the model grid returns the per-bin photons (z = 0.5 set; the overlap
    # fractions are exactly 1, the photons are grouped L * [dE / (h E)] / [...] instead of left to right: <= 2 ulp)
    k = 1
    ear = eb / (1 + par[k, 14])
    ctx.set_instrument(ear)
    idn = np.empty((P, nE))
    ctx.photar_device(L.ctypes.data, par.ctypes.data, P, idn.ctypes.data)
    np.testing.assert_allclose(idn[k], cnt[k], rtol=1e-15, atol=0)
    # photon conservation on a coarse grid spanning the whole (redshifted) model grid
    ear = np.geomspace(eb[0] / 1.5, eb[-1], 40)
    ear[0], ear[-1] = eb[0] / 1.5, eb[-1]
    ctx.set_instrument(ear)
    co = np.empty((P, 39))
    ctx.photar_device(L.ctypes.data, par.ctypes.data, P, co.ctypes.data)
    assert co[k].sum() == pytest.approx(cnt[k].sum(), rel=1e-13)

    # ---- f2: response fold + statistic ----
    ear = np.geomspace(0.3, 10.0, 201)
    rp, cl, vl = toy_response(ear)
    expo = 5e4
    ctx.set_instrument(ear)
    ctx.set_response(rp, cl, vl, expo)
    pho = np.stack([O.photar(L[k], eb, par[k, 1], par[k, 14], ear) for k in range(P)])
    _, folded0 = O.fold_stat(pho[0], rp, cl, vl, expo, np.zeros(rp.size - 1), None, 1)
    rng = np.random.Generator(np.random.PCG64(3))
    data = rng.poisson(np.maximum(folded0, 0)).astype(float)
    sig = np.sqrt(np.maximum(data, 1.0))
    for stat, kind in (("chi2", 0), ("cstat", 1)):
        ctx.set_data(data, sig if kind == 0 else None, stat)
        st = np.empty(P)
        fo = np.empty((P, rp.size - 1))
        ctx.fitstat_device(L.ctypes.data, par.ctypes.data, P, st.ctypes.data, fo.ctypes.data)
        for k in range(P):
            s_ref, f_ref = O.fold_stat(pho[k], rp, cl, vl, expo, data, sig if kind == 0 else None, kind)
            assert np.array_equal(fo[k], f_ref)
            assert st[k] == pytest.approx(s_ref, rel=1e-12)
        assert st[0] < st[2]  # the data were drawn from set 0
    # parameters -> statistic in one call equals evaluate + epilogue
    st2 = ctx.eval_fitstat_host(par, model=0, rel=False)
    spec, _ = ctx.eval_batch_host(par, model=0, rel=False)
    for k in range(P):
        s_ref, _ = O.fold_stat(O.photar(spec[k], eb, par[k, 1], par[k, 14], ear), rp, cl, vl, expo, data, None, 1)
        assert st2[k] == pytest.approx(s_ref, rel=1e-12)
    # error behaviour
    with pytest.raises(Exception):
        ctx.set_data(data[:-1], None, "cstat")
    with pytest.raises(Exception):
        ctx.set_instrument(np.array([1.0, 0.5, 2.0]))
    with pytest.raises(Exception):
        ctx.set_response(rp, cl + 1000, vl, expo)


def test_xspec_local_model_abi(ctx):
    """Seam B3 on the CPU-only box: relagn_ / relqso_ called the way XSPEC calls a Fortran local model (real*4 by
    reference) against the oracle's statement of relagn.f:60-112."""
    from oracle import oracle as O
    from tests.util import xspec_reference
    t = golden_tables()
    otab = O.Tables(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], data=t["data"])
    ctx.xspec_bind()
    try:
        ear = np.linspace(0.3, 10.0, 101).astype(np.float32)
        par = np.array([1e8, 100, -1, 0.7, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0.05], dtype=np.float32)
        got = ctx.xspec_model("relagn", ear, par)
        ref = xspec_reference(O, otab, 0, True, ear.astype(np.float64), par.astype(np.float64))
        assert got.dtype == np.float32 and np.all(got > 0)
        np.testing.assert_allclose(got, ref.astype(np.float32), rtol=3e-7)
        assert np.array_equal(ctx.xspec_model("relagn", ear, par), got)  # memoised (relagn.f:37,62-66)
        # relqso: M, Dist, log_mdot, astar, cos_inc, redshift, $rel
        q = np.array([1e7, 10, -0.5, 0.0, 0.5, 0.0, 0], dtype=np.float32)
        got = ctx.xspec_model("relqso", ear[:40], q)
        p7 = [float(q[0]), float(q[1]), float(q[2]), float(q[3]), float(q[4]), 1.0, float(q[5])]
        ref = xspec_reference(O, otab, 1, False, ear[:40].astype(np.float64), np.array(p7))
        np.testing.assert_allclose(got, ref.astype(np.float32), rtol=3e-7)
    finally:
        ctx.lib.rg_xspec_bind(None)


def test_hot_spectral_shape_attribute(ctx):
    """rg_last_hot_shape: Fnu_seed_hot (relagn.py:1200-1223) of a set of the last batch, against the reference."""
    d = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    ctx.set_grid(d["grid1000"])
    sel = [0, 1, 4]
    ctx.eval_batch_host(d["relagn_params"][sel], model=0, rel=False)
    for k, i in enumerate(sel):
        ref = d["relagn_fnu_seed_hot1000"][i]
        got = ctx.last_hot_shape(k)
        if ref.max() == 0:
            assert got.max() == 0
        else:
            assert peak_relerr(got, ref) < 1e-10, i


@pytest.mark.parametrize("grid_case", ["wide_10_decades", "hard_band_above_1keV", "soft_band_below_1keV",
                                       "above_the_warm_x_grid", "below_every_x_grid"])
def test_kompaneets_large_path_unusual_grids_vs_oracle(ctx, monkeypatch, grid_case):
    """The large-batch Kompaneets kernels (forward sweep per thread, back substitution as a warp scan, interpolation
    lines in shared memory) on grids that move their row bookkeeping: a 10-decade grid (wider than a warp's table: the
    variant with the tables in global scratch; starts below 1e-4 keV: low-frequency extension of the disc), a band that
    starts above 1 keV (the normalisation rows at 1 keV lie below the grid's first cell) and one that ends below it
    (they lie above its last).  Non-relativistic, against the oracle's restatement of pyNTHCOMP.py."""
    from oracle import oracle as O
    from tests.util import c4_params
    grid = {"wide_10_decades": np.geomspace(1e-6, 1e4, 1201), "hard_band_above_1keV": np.geomspace(2.0, 150.0, 401),
            "soft_band_below_1keV": np.geomspace(1e-3, 0.5, 301),
            # every bin above the top of the warm Comptonisation grids (40 kTe_w <= 40 keV) / below the first row of every
            # Kompaneets grid (511 xmin = 1e-4 kT_seed ~ 1e-7 keV): the cursor's clamps, empty and constant spectra
            "above_the_warm_x_grid": np.geomspace(45.0, 400.0, 101),
            "below_every_x_grid": np.geomspace(1e-11, 3e-10, 65)}[grid_case]
    monkeypatch.setenv("RG_NO_SMALL_PATH", "1")
    p = c4_params(6, seed=77)
    ctx.set_grid(grid)
    st, ref, _ = O.evaluate_batch(p, grid, model=0, rel=False)
    outs = []
    for force_global in (False, True):
        if force_global:
            monkeypatch.setenv("RG_INTERP_GLOBAL", "1")
        out, at = ctx.eval_batch_host(p, model=0, rel=False, components=True)
        outs.append(out)
        for k in range(p.shape[0]):
            if st[k] != 0:
                continue
            for c in range(4):
                assert peak_relerr(out[k, c], ref[k, c]) < 1e-9, (grid_case, force_global, k, c)
    np.testing.assert_array_equal(outs[0], outs[1])  # both table placements: same bits
    if grid_case == "hard_band_above_1keV":
        # relativistic, same band: the cold annuli show only their far Wien tail, which the reference scales up to the
        # annulus' full luminosity -- norm / radiance leaves the double range, norm (B / radiance) does not
        t = golden_tables()
        otab = O.Tables(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], t["data"])
        pr = p[:3].copy()
        pr[:, 3] = np.clip(pr[:, 3], t["spins"].min(), t["spins"].max())
        st, ref, _ = O.evaluate_batch(pr, grid, model=0, rel=True, tab=otab)
        out, at = ctx.eval_batch_host(pr, model=0, rel=True, components=True)
        assert np.isfinite(out).all()
        for k in range(pr.shape[0]):
            if st[k] != 0:
                continue
            for c in range(4):
                assert peak_relerr(out[k, c], ref[k, c]) < 1e-9, ("rel", k, c)


def test_large_path_random_sets_vs_oracle(ctx, monkeypatch):
    """The large-batch kernels over a spread of corona parameters (24 C4 draws: Gamma 1.3-5, kTe 0.1-300 keV, seed
    temperatures from 1e5 to 1e10 solar masses' discs): Kompaneets forward sweep + scan + interpolation lines, the
    coalesced warm-row loads and the two-CTA item loop of the spectrum kernel, against the oracle; non-relativistic
    for all, relativistic for the first six."""
    from oracle import oracle as O
    from tests.util import c4_params
    monkeypatch.setenv("RG_NO_SMALL_PATH", "1")
    grid = np.geomspace(1e-4, 1e3, 1001)
    ctx.set_grid(grid)
    t = golden_tables()
    p = c4_params(24, seed=2024)
    p[:, 3] = np.clip(p[:, 3], t["spins"].min(), t["spins"].max())
    st, ref, rat = O.evaluate_batch(p, grid, model=0, rel=False)
    out, at = ctx.eval_batch_host(p, model=0, rel=False, components=True)
    np.testing.assert_array_equal(at[:, 21].astype(int), st)
    for k in np.nonzero(st == 0)[0]:
        for c in range(4):
            assert peak_relerr(out[k, c], ref[k, c]) < 1e-9, (k, c)
    otab = O.Tables(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], t["data"])
    st, ref, _ = O.evaluate_batch(p[:6], grid, model=0, rel=True, tab=otab)
    out, at = ctx.eval_batch_host(p[:6], model=0, rel=True, components=True)
    for k in np.nonzero(st == 0)[0]:
        for c in range(4):
            assert peak_relerr(out[k, c], ref[k, c]) < 1e-9, ("rel", k, c)


def test_large_path_relqso_vs_oracle(ctx, monkeypatch):
    """relqso through the large-batch kernels (r_hot search, gamma_hot from lumin_kernel's Ldiss / Lseed, Kompaneets
    kernels) against the oracle: a 3 x 3 mass x mdot grid with spin and inclination varied."""
    from oracle import oracle as O
    from tests.util import c3_params
    monkeypatch.setenv("RG_NO_SMALL_PATH", "1")
    grid = np.geomspace(1e-4, 1e3, 1001)
    ctx.set_grid(grid)
    t = golden_tables()
    otab = O.Tables(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], t["data"])
    p = c3_params(3)
    p[:, 3] = np.linspace(0.0, 0.9, p.shape[0])
    p[:, 4] = np.linspace(0.3, 0.95, p.shape[0])
    for rel in (False, True):
        st, ref, rat = O.evaluate_batch(p, grid, model=1, rel=rel, tab=otab)
        out, at = ctx.eval_batch_host(p, model=1, rel=rel, components=True)
        np.testing.assert_array_equal(at[:, 21].astype(int), st)
        ok = np.nonzero(st == 0)[0]
        assert ok.size >= 6
        np.testing.assert_allclose(at[ok, 10], rat[ok, 10], rtol=1e-9)  # gamma_h
        for k in ok:
            for c in range(4):
                assert peak_relerr(out[k, c], ref[k, c]) < 1e-9, (rel, k, c)
