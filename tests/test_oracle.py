"""The CPU oracle (oracle/) against the golden fixtures generated from the
UNMODIFIED reference (tests/golden/make_golden.py) and the reference's only
published known answer (examples/relqso_modSpec.ipynb:81-83)."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def relerr(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    scale = np.maximum(np.abs(b), 1e-300)
    m = (b != 0) | (a != 0)
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a - b)[m] / scale[m]))


def peak_relerr(a, b, floor=1e-12):
    """max |a-b| / max(|b|, floor*peak): the far Wien tail (1e-200 of the peak) is excluded."""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    pk = np.max(np.abs(b))
    if pk == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * pk)))


def test_notebook_known_answer():
    # examples/relqso_modSpec.ipynb:81-83
    st, at = O.setup([1e5, 1, -1, 0, 0.5, 1, 0], model=1)
    assert st == 0
    assert at["gamma_h"] == pytest.approx(1.9803689563797076, rel=1e-12)
    assert at["r_h"] == 14.261351436104999
    assert at["r_w"] == 28.522702872209997


def test_analytic_kats():
    L = O.lib()
    assert L.ro_risco(0.0) == 6.0
    assert L.ro_risco(0.998) == pytest.approx(1.236971, abs=1e-6)
    st, at = O.setup([1e8, 100, -1, 0, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0])
    assert at["eta"] == pytest.approx(1 - np.sqrt(8.0 / 9.0), rel=1e-14)
    assert at["r_sg"] == pytest.approx(772.6699377179948, rel=1e-13)
    assert (at["n_disc"], at["n_warm"], at["n_hot"]) == (79, 15, 11)


@pytest.mark.parametrize("tag,model", [("relagn", 0), ("relqso", 1)])
def test_attrs_and_nonrel_spectra(tag, model):
    d = np.load(os.path.join(G, "golden_nonrel.npz"))
    names = list(d["attr_names"])
    for i, name in enumerate(d[tag + "_names"]):
        pars = d[tag + "_params"][i]
        for grid, key in ((d["grid1000"], "_spec1000"), (d["grid_def"], "_specdef")):
            st, out, at = O.evaluate(pars, grid, model=model, rel=False)
            assert st == 0
            gold = d[tag + key][i]
            for c in range(4):
                assert peak_relerr(out[c], gold[c]) < 1e-9, (tag, name, key, c)
        for k, nm in enumerate(names):
            assert at[nm] == pytest.approx(d[tag + "_attrs"][i][k], rel=1e-11, abs=0), (name, nm)
        ex = d[tag + "_extra"][i]
        assert at["Ldiss"] == pytest.approx(ex[0], rel=1e-7)
        assert at["Lseed"] == pytest.approx(ex[1], rel=1e-11)
        nb = ex[2:].astype(int)  # edge counts of the reference's three bin arrays
        has = [at["n_disc"], at["n_warm"], at["n_hot"]]
        for n_edges, n_ann in zip(nb, has):
            assert n_ann in (n_edges - 1, 0)


def test_nthcomp_golden():
    d = np.load(os.path.join(G, "golden_nthcomp.npz"))
    for i, (g, kte, kbb) in enumerate(d["par"]):
        for eg, key in ((d["egrid_def"], "out_def"), (d["egrid_1000"], "out_1000")):
            got = O.donthcomp(eg, g, kte, kbb)
            assert peak_relerr(got, d[key][i]) < 1e-9, (i, key)


def test_rel_golden_reference_driver():
    """reference driver (unmodified relagn.py do_rel*Spec) + oracle kyconv at the seam vs the
    oracle's own whole-path evaluation: pins pack/unpack, the 10^3 Rg switch and the region sums."""
    d = np.load(os.path.join(G, "golden_rel.npz"))
    tab = O.Tables(d["tab_spins"], d["tab_thetas"], int(d["tab_NR"]), int(d["tab_NPHI"]), float(d["tab_r_in"]),
                   data=d["tab_data"])
    for tag, model in (("relagn", 0), ("relqso", 1)):
        for i, name in enumerate(d[tag + "_names"]):
            st, out, at = O.evaluate(d[tag + "_params"][i], d["grid1000"], model=model, rel=True, tab=tab)
            assert st == 0
            gold = d[tag + "_spec1000"][i]
            for c in range(4):
                assert peak_relerr(out[c], gold[c]) < 1e-9, (tag, name, c)


def test_kyconv_flat_limit_and_additivity():
    d = np.load(os.path.join(G, "golden_rel.npz"))
    tab = O.Tables(d["tab_spins"], d["tab_thetas"], int(d["tab_NR"]), int(d["tab_NPHI"]), float(d["tab_r_in"]),
                   data=d["tab_data"])
    dlnE = np.log(1e7) / 1000
    th = np.deg2rad(60.0)
    # flat-space limit (relagn.py:968-969, 1010-1013): sum H -> pi (r2^2 - r1^2) cos i at large r
    k0, H = O.ky_kernel(tab, 0.0, th, 900.0, 950.0, dlnE)
    assert H.sum() == pytest.approx(np.pi * (950.0 ** 2 - 900.0 ** 2) * np.cos(th), rel=2e-2)
    # additivity over annuli (tests/blurr_test.py:77-139): sum of annulus kernels == whole-disc kernel
    edges = np.geomspace(6.0, 100.0, 41)
    kw, Hw = O.ky_kernel(tab, 0.0, th, 6.0, 100.0, dlnE, nsub=40 * 4)
    acc = np.zeros(4096)
    for r1, r2 in zip(edges[:-1], edges[1:]):
        k, Hh = O.ky_kernel(tab, 0.0, th, r1, r2, dlnE, nsub=4)
        acc[k + 2048:k + 2048 + Hh.size] += Hh
    whole = np.zeros(4096)
    whole[kw + 2048:kw + 2048 + Hw.size] = Hw
    assert np.max(np.abs(acc - whole)) < 1e-9 * whole.max()


def test_observation_epilogue_oracle_pins():
    """observe_oracle.c against the unmodified reference's getters (golden_units.npz) and its own invariants."""
    import os
    from tests.util import GOLDEN, peak_relerr
    u = np.load(os.path.join(GOLDEN, "golden_units.npz"))
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    eb = dn["grid1000"]
    L = dn["relagn_spec1000"][0][3]
    for unit in ["cgs", "cgs_wave", "SI", "SI_wave", "counts"]:
        for fl in (0, 1):
            assert peak_relerr(O.convert(L, eb, unit, fl, 100.0, 0.0), u["%s_%d" % (unit, fl)]) < 1e-14
    Lz = dn["relagn_spec1000"][6][3]  # z = 0.5
    cz = O.convert(Lz, eb, "counts", 1, 100.0, 0.5)
    assert peak_relerr(cz, u["z05_counts_1"]) < 1e-14
    # XSPEC contract (relagn.f:98-112): on the observer-frame image of the model grid photar = counts flux x dE
    ph = O.photar(Lz, eb, 100.0, 0.5, eb / 1.5)
    np.testing.assert_allclose(ph, cz * np.diff(eb), rtol=1e-15, atol=0)
    # photon conservation on any grid that covers the model grid; zeros outside it; clipped edge bins
    ear = np.geomspace(eb[0] / 3, eb[-1] * 2, 57)
    assert O.photar(Lz, eb, 100.0, 0.5, ear).sum() == pytest.approx(ph.sum(), rel=1e-13)
    assert np.all(O.photar(Lz, eb, 100.0, 0.5, np.array([2e3, 3e3, 5e3])) == 0)
    half = O.photar(Lz, eb, 100.0, 0.5, np.array([eb[10] / 1.5, 0.5 * (eb[10] + eb[11]) / 1.5]))
    assert half[0] == pytest.approx(0.5 * ph[10], rel=1e-12)
    # statistic: chi^2 of the model against itself is 0; Cash of the model against itself is 0 and positive elsewhere
    from tests.util import toy_response
    ear = np.geomspace(0.3, 10.0, 201)
    rp, cl, vl = toy_response(ear)
    pa = O.photar(L, eb, 100.0, 0.0, ear)
    _, folded = O.fold_stat(pa, rp, cl, vl, 1e4, np.zeros(rp.size - 1), None, 1)
    assert folded[5] == 0 and np.all(folded[6:] > 0)
    s0, _ = O.fold_stat(pa, rp, cl, vl, 1e4, folded, np.ones_like(folded), 0)
    assert s0 == 0
    c0, _ = O.fold_stat(pa, rp, cl, vl, 1e4, folded, None, 1)
    assert abs(c0) < 1e-6 * folded.sum() * 1e-6 + 1e-290 * 2 * folded.size + 1e-9
    c1, _ = O.fold_stat(pa, rp, cl, vl, 1e4, 1.2 * folded, None, 1)
    assert c1 > 0


def test_hot_spectral_shape_vs_reference():
    """Fnu_seed_hot (relagn.py:1200-1223), an attribute of the reference's class, on the 1000-bin grid."""
    d = np.load(os.path.join(G, "golden_nonrel.npz"))
    for tag, model in (("relagn", 0), ("relqso", 1)):
        for i in range(len(d[tag + "_names"])):
            ref = d[tag + "_fnu_seed_hot1000"][i]
            got = O.hot_shape(d[tag + "_params"][i], d["grid1000"], model=model)
            if ref.max() == 0:
                assert got.max() == 0, (tag, i)
            else:
                assert peak_relerr(got, ref) < 1e-10, (tag, d[tag + "_names"][i])
