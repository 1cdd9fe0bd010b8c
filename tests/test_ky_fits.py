"""importer for Kerr transfer tables in the FITS layout of the KY package (relagn_b200/ky_fits.py) and the
limb-darkening laws of kyconv's parameter 10 (relagn.py:979-991).  UNVERIFIED AGAINST A REAL HEASOFT FILE (none is
available offline): the fixture is exported from the oracle's own ray tracer in that layout -- rows r_h + r_vector,
columns at absolute azimuth, planes alpha / beta / lensing / g-factor / cosine / delay -- and imported back."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from relagn_b200 import ky_fits
from tests.util import ring_metrics

SPINS = [0.9, 0.95]
INCS = [30.0, 60.0]
N_R, N_PHI = 100, 128


@pytest.fixture(scope="module")
def ky_file(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("ky") / "KBHtables_fixture.fits")
    r_vector = np.geomspace(0.3, 1000.0, N_R)
    phi = 2 * np.pi * np.arange(N_PHI) / N_PHI
    rh = np.array([1 + np.sqrt(1 - a * a) for a in SPINS])
    nodes = {}
    for ih, a in enumerate(SPINS):
        for ii, inc in enumerate(INCS):
            n = O.make_node_full(a, np.deg2rad(inc), rh[ih] + r_vector, N_PHI, absolute_phi=True)
            assert n["unresolved"] == 0
            g, jn, mue = n["g"].astype(np.float64), n["jn"].astype(np.float64), n["mue"]
            nodes[(ih, ii)] = dict(alpha=n["alpha"], beta=n["beta"], g=g, cosine=mue, lensing=jn * g / mue)
    ky_fits.write_ky_fits(path, rh, INCS, r_vector, phi, nodes)
    return path, nodes, rh, r_vector


def test_fits_io_round_trip(ky_file):
    path, nodes, rh, rv = ky_file
    hdus = ky_fits.read_fits_tables(path)
    assert [n for n, _ in hdus[:4]] == ["r_horizon", "inclination", "r_vector", "phi_vector"]
    assert len(hdus) == 4 + len(SPINS) * len(INCS)
    np.testing.assert_allclose(hdus[0][1]["r_horizon"][:, 0], rh, rtol=1e-7)
    np.testing.assert_allclose(hdus[2][1]["r_vector"][:, 0], rv, rtol=1e-7)
    cols = hdus[4 + 1 * len(INCS) + 0][1]   # (spin 1, inclination 0)
    assert set(cols) == {"alpha", "beta", "lensing", "g-factor", "cosine", "delay"}
    np.testing.assert_array_equal(cols["g-factor"], nodes[(1, 0)]["g"].astype(np.float32).astype(np.float64))
    assert cols["alpha"].shape == (N_R, N_PHI)


def test_import_matches_directly_traced_nodes(ky_file):
    """KY layout -> table format 2 (resampled rows, front-ray azimuth origin, Jn = l mu_e / g) against nodes traced
    directly in format 2: ring flux and shift centroid per row."""
    path, _, _, _ = ky_file
    NR, NPHI, r_in = 32, 64, 2.0
    t = ky_fits.import_ky_tables(path, NR=NR, NPHI=NPHI, r_in=r_in)
    np.testing.assert_allclose(t["spins"], SPINS, rtol=1e-6)
    np.testing.assert_allclose(np.rad2deg(t["thetas"]), INCS, rtol=1e-6)
    assert t["data"].shape == (2, 2, 2, NR, NPHI)
    for m, a in enumerate(SPINS):
        for n, inc in enumerate(INCS):
            g, jn, bad = O.make_node(a, np.deg2rad(inc), NR, NPHI, r_in, threads=4)
            gi, ji = t["data"][m, n, 0], t["data"][m, n, 1]
            assert (gi > 0).all() and (g > 0).all()
            F0, c0 = ring_metrics(np.log(g.astype(np.float64)), jn.astype(np.float64) * g.astype(np.float64) ** 3)
            F1, c1 = ring_metrics(np.log(gi.astype(np.float64)), ji.astype(np.float64) * gi.astype(np.float64) ** 3)
            assert np.abs(F1 / F0 - 1).max() < 5e-3, (a, inc, np.abs(F1 / F0 - 1).max())
            assert np.abs(c1 - c0).max() < 0.1, (a, inc, np.abs(c1 - c0).max())
            # column alignment: the front ray sits at column 0 in both (alpha = 0: the shift there is purely
            # gravitational + transverse Doppler, the same number from either route)
            np.testing.assert_allclose(gi[:, 0], g[:, 0], rtol=2e-3)
    # the imported set is a valid table set for the seam
    tab = O.Tables(t["spins"], t["thetas"], NR, NPHI, r_in, data=t["data"])
    eb = np.geomspace(0.1, 10.0, 201)
    line = np.zeros(200)
    line[100] = 1.0
    out = O.kyconv(tab, eb, [0.92, 45.0, 6.0, 0, 20.0, 0, 0, 10.0, 0, 0, 200, -1], line)
    assert out.sum() > 0 and np.isfinite(out).all()


def test_limb_darkening_laws(ky_file):
    """kyconv parameter 10 (relagn.py:988; the reference passes 0): the law multiplies the weight of every sample."""
    path, nodes, _, _ = ky_file
    iso = ky_fits.import_ky_tables(path, NR=24, NPHI=64, r_in=2.0, limb=0)["data"]
    laor = ky_fits.import_ky_tables(path, NR=24, NPHI=64, r_in=2.0, limb=-1)["data"]
    haardt = ky_fits.import_ky_tables(path, NR=24, NPHI=64, r_in=2.0, limb=-2)["data"]
    np.testing.assert_array_equal(iso[:, :, 0], laor[:, :, 0])          # g does not depend on the law
    ratio = laor[:, :, 1] / iso[:, :, 1]
    assert ratio.min() >= 1.0 and ratio.max() <= 3.06 + 1e-6             # 1 + 2.06 mu_e, 0 <= mu_e <= 1
    assert (haardt[:, :, 1] > 0).all()
    # far from the hole mu_e -> cos(i): the laws reduce to constants there
    for n, inc in enumerate(INCS):
        mu = np.cos(np.deg2rad(inc))
        assert abs(ratio[0, n, -1].mean() / (1 + 2.06 * mu) - 1) < 2e-2
        assert abs((haardt[0, n, 1, -1] / iso[0, n, 1, -1]).mean() / np.log(1 + 1 / mu) - 1) < 2e-2
    # the oracle's generator with the law baked in agrees with Jn (1 + 2.06 mu_e) of the isotropic one
    rows = np.array([O.lib().ro_row_radius(2.0, 12, s) for s in range(12)])
    a_iso = O.make_node_full(0.9, np.deg2rad(60.0), rows, 32)
    a_laor = O.make_node_full(0.9, np.deg2rad(60.0), rows, 32, limb=-1)
    np.testing.assert_allclose(a_laor["jn"], a_iso["jn"] * (1 + 2.06 * a_iso["mue"]), rtol=3e-7)
    with pytest.raises(ValueError):
        ky_fits.limb_factor(3, 0.5)
