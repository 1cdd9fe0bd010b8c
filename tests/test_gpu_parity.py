"""Parity of the CUDA path (through the C ABI, librelagn_b200.so) against the CPU oracle and the golden fixtures
generated from the unmodified reference.  Tolerances: FP64 arithmetic on both sides; the two differ only by
summation order, FMA contraction and CUDA-libm vs glibc (<= 2 ulp) -> 1e-9 of the bin value for every bin within
12 decades of the spectral peak (north_star asks <= 1e-5)."""
import os

import numpy as np
import pytest

from tests.util import GOLDEN, peak_relerr, golden_tables, c4_params, c3_params, offnode_errors

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from relagn_b200 import _capi, build
    build.build()  # no-op when relagn_b200/_build/librelagn_b200.so is current (it travels with the snapshot)
    c = _capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as O
    return O


@pytest.fixture(scope="module")
def otab(oracle):
    t = golden_tables()
    return oracle.Tables(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], data=t["data"])


def upload_golden_tables(ctx):
    t = golden_tables()
    ctx.tables_upload(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], t["data"])


@pytest.mark.parametrize("tag,model", [("relagn", 0), ("relqso", 1)])
@pytest.mark.parametrize("gridkey,speckey", [("grid1000", "_spec1000"), ("grid_def", "_specdef")])
def test_nonrel_vs_reference_golden(ctx, tag, model, gridkey, speckey):
    d = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    ctx.set_grid(d[gridkey])
    out, at = ctx.eval_batch_host(d[tag + "_params"], model=model, rel=False, components=True)
    for i, name in enumerate(d[tag + "_names"]):
        for c in range(4):
            assert peak_relerr(out[i, c], d[tag + speckey][i][c]) < TOL, (tag, name, c)
    from relagn_b200._capi import A
    for k, nm in enumerate(d["attr_names"]):
        np.testing.assert_allclose(at[:, A[str(nm)]], d[tag + "_attrs"][:, k], rtol=1e-11, err_msg=str(nm))
    np.testing.assert_allclose(at[:, A["Ldiss"]], d[tag + "_extra"][:, 0], rtol=1e-7)
    np.testing.assert_allclose(at[:, A["Lseed"]], d[tag + "_extra"][:, 1], rtol=1e-11)


def test_notebook_known_answer(ctx):
    # examples/relqso_modSpec.ipynb:81-83
    d = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    ctx.set_grid(d["grid1000"])
    from relagn_b200._capi import A
    out, at = ctx.eval_batch_host([[1e5, 1, -1, 0, 0.5, 1, 0]], model=1, rel=False)
    assert at[0, A["r_h"]] == 14.261351436104999
    assert at[0, A["r_w"]] == 28.522702872209997
    assert at[0, A["gamma_h"]] == pytest.approx(1.9803689563797076, rel=1e-12)


@pytest.mark.parametrize("tag,model", [("relagn", 0), ("relqso", 1)])
def test_rel_vs_reference_driver_golden(ctx, tag, model):
    """unmodified reference do_rel*Spec + oracle kyconv at the seam (tests/golden/make_golden.py)"""
    d = np.load(os.path.join(GOLDEN, "golden_rel.npz"))
    ctx.set_grid(d["grid1000"])
    upload_golden_tables(ctx)
    out, at = ctx.eval_batch_host(d[tag + "_params"], model=model, rel=True, components=True)
    for i, name in enumerate(d[tag + "_names"]):
        for c in range(4):
            assert peak_relerr(out[i, c], d[tag + "_spec1000"][i][c]) < TOL, (tag, name, c)


@pytest.mark.parametrize("rel", [False, True])
def test_random_batch_vs_oracle(ctx, oracle, otab, rel):
    grid = np.geomspace(1e-4, 1e3, 1001)
    ctx.set_grid(grid)
    upload_golden_tables(ctx)
    p = c4_params(96, seed=7)
    # edge cases the reference handles: absent regions, clamps, fcol < 0, outer disc beyond 10^3 Rg
    p[0, 9] = -1            # no hot region
    p[1, 10] = p[1, 9]      # no warm region
    p[2, 12] = -1           # colour correction on
    p[3, 11] = 3.6; p[3, 0] = 1e6   # flat-space annuli beyond 10^3 Rg
    p[4, 10] = 1e6          # r_w > r_out -> clamp, no disc
    p[5, 9] = 2.0           # r_h < risco -> clamp
    p[6, 3] = -0.9          # retrograde
    p[7, 4] = 1.0           # face on
    st, ref, rat = oracle.evaluate_batch(p, grid, model=0, rel=rel, tab=otab)
    out, at = ctx.eval_batch_host(p, model=0, rel=rel, components=True)
    assert (st == 0).all()
    np.testing.assert_array_equal(at[:, 17:20], rat[:, 17:20])  # annulus counts
    np.testing.assert_array_equal(at[:, 20], rat[:, 20])        # warning flags
    for i in range(p.shape[0]):
        for c in range(4):
            assert peak_relerr(out[i, c], ref[i, c]) < TOL, (i, c)
    tot, _ = ctx.eval_batch_host(p, model=0, rel=rel, components=False)
    # total-only mode accumulates all regions in one register set; component mode sums three stored components
    np.testing.assert_allclose(tot, out[:, 3], rtol=1e-13)


def test_relqso_grid_large_batch_vs_oracle(ctx, oracle, otab):
    """relqso through the large-batch path (a 10 x 10 mass x mdot grid = 100 sets, config 3 in small): r_hot search,
    gamma_hot from Ldiss / Lseed (lumin_kernel), Kompaneets kernels, against the oracle; spin and inclination varied."""
    grid = np.geomspace(1e-4, 1e3, 1001)
    ctx.set_grid(grid)
    upload_golden_tables(ctx)
    p = c3_params(10)
    p[:, 3] = np.linspace(0.0, 0.95, p.shape[0])
    p[:, 4] = np.linspace(0.2, 0.95, p.shape[0])[::-1]
    for rel in (False, True):
        st, ref, rat = oracle.evaluate_batch(p, grid, model=1, rel=rel, tab=otab)
        out, at = ctx.eval_batch_host(p, model=1, rel=rel, components=True)
        np.testing.assert_array_equal(at[:, 21].astype(int), st)
        ok = st == 0
        assert ok.sum() >= 80
        np.testing.assert_allclose(at[ok, 10], rat[ok, 10], rtol=1e-9)  # gamma_h
        for i in np.nonzero(ok)[0]:
            for c in range(4):
                assert peak_relerr(out[i, c], ref[i, c]) < TOL, (rel, i, c)


def test_raytracer_vs_oracle_nodes(ctx, oracle):
    """The device ray tracer (rg_tables_generate, csrc/rg_raytrace.cu) against the oracle's node generator
    (kerr_oracle.c: ro_make_node) on 9 nodes incl. the corners of the default grid: retrograde / maximal spin,
    face-on / 85 degrees.  Same algorithm on both sides; the differences are CUDA libm vs glibc in the geodesic
    integration: g agrees to the float32 rounding (<= 1 ulp, most points bit-equal), Jn to 2e-7 in the median and 2e-5
    at worst (the Jacobian is a finite-difference quotient); the sets of points either solver lost differ by <= 4."""
    NR, NPHI, r_in = 64, 64, 1.2369
    nodes = [(-0.998, 0.0), (-0.998, 85.0), (0.0, 40.0), (0.5, 0.0), (0.9, 75.0), (0.98, 82.0), (0.998, 85.0),
             (0.998, 30.0), (0.9875, 60.0)]
    for a, deg in nodes:
        th = np.deg2rad(deg)
        bad = ctx.tables_generate(np.array([a]), np.array([th]), NR, NPHI, r_in)
        d = ctx.tables_download()[0, 0]
        g, jn, obad = oracle.make_node(a, th, NR, NPHI, r_in)
        rows = np.array([oracle.lib().ro_row_radius(r_in, NR, s) for s in range(NR)])
        inside = rows <= oracle.lib().ro_r_valid(a)
        assert (d[0][inside] == 0).all() and (d[1][inside] == 0).all()
        lost_dev, lost_orc = (d[1] == 0) & ~inside[:, None], (jn == 0) & ~inside[:, None]
        assert (lost_dev != lost_orc).sum() <= 4, (a, deg, lost_dev.sum(), lost_orc.sum())
        assert abs(bad - obad) <= 4 and bad <= 8, (a, deg, bad, obad)
        # unresolved points sit on the innermost rows only (inside the ISCO of every model that uses this node's rows
        # there, or at the ISCO ring of a = 0.998 seen at >= 80 degrees)
        if lost_dev.any():
            assert rows[np.where(lost_dev)[0]].max() < 1.3, (a, deg)
        ok = ~lost_dev & ~lost_orc & ~inside[:, None]
        np.testing.assert_allclose(d[0][ok], g[ok], rtol=2.4e-7)
        assert (d[0][ok] != g[ok]).mean() < 0.02, (a, deg)
        np.testing.assert_allclose(d[1][ok], jn[ok], rtol=2e-5)
        assert np.median(np.abs(d[1][ok] / jn[ok] - 1)) < 2e-7, (a, deg)


def test_default_tables_parity_c4(ctx, oracle):
    """The configuration bench.py runs: the DEFAULT 14 x 14 node grid (relagn_b200/tables.py; the cached array is
    ray-traced on the device), 2000 C4 sets plus the corners of the parameter box -- retrograde and maximal spin,
    cos i = 1.0 and 0.09 -- GPU against the oracle on the same table array."""
    from relagn_b200 import tables as T
    grid = np.geomspace(1e-4, 1e3, 1001)
    ctx.set_grid(grid)
    T.load_default(ctx)
    data = ctx.tables_download()
    cfg = T.DEFAULT
    otab = oracle.Tables(cfg["spins"], cfg["thetas"], cfg["NR"], cfg["NPHI"], cfg["r_in"], data=data)
    p = c4_params(2000, seed=23)
    p[0, 3], p[0, 4] = -0.998, 1.0
    p[1, 3], p[1, 4] = 0.998, 0.09
    p[2, 3], p[2, 4] = -0.5, 0.09
    p[3, 3], p[3, 4] = 0.998, 1.0
    p[4, 3], p[4, 4] = 0.0, 0.5
    p[5, 3], p[5, 4] = 0.9901, 0.2
    for k in range(6):  # keep the geometry legal for the changed spin (r_h >= risco)
        p[k, 9] = 30.0; p[k, 10] = 60.0; p[k, 13] = 10.0
    st, ref, rat = oracle.evaluate_batch(p, grid, model=0, rel=True, tab=otab)
    out, at = ctx.eval_batch_host(p, model=0, rel=True, components=True)
    assert (st == 0).all() and (at[:, 21] == 0).all()
    worst = max(peak_relerr(out[i, c], ref[i, c]) for i in range(p.shape[0]) for c in range(4))
    assert worst < TOL, worst
    upload_golden_tables(ctx)


def test_offnode_self_consistency_default_tables(ctx, oracle):
    """the default table set blended at 32 random (a, i) -- where every benchmarked spectrum
    lives -- against a node ray-traced AT (a, i): ring flux and centroid of the shift histogram per table row.
    Asked: flux <= 1e-3, centroid <= 0.05 bin (1000-bin grid) for r >= 1.5 r_isco."""
    from relagn_b200 import tables as T
    cfg = T.DEFAULT
    T.load_default(ctx)
    otab = oracle.Tables(cfg["spins"], cfg["thetas"], cfg["NR"], cfg["NPHI"], cfg["r_in"], data=ctx.tables_download())
    rows = otab.rows()
    rng = np.random.default_rng(20241020)
    a_t = np.concatenate([rng.uniform(0.0, 0.998, 20), rng.uniform(-0.998, 0.0, 4), rng.uniform(0.95, 0.998, 8)])
    th_t = np.arccos(rng.uniform(0.09, 1.0, a_t.size))
    worst = np.zeros(4)
    for a, th in zip(a_t, th_t):
        bad = ctx.tables_generate(np.array([a]), np.array([th]), cfg["NR"], cfg["NPHI"], cfg["r_in"])
        d = ctx.tables_download()[0, 0]
        lg_b, jn_b, s_lo = otab.slice(a, th)
        e = offnode_errors(rows, a, lg_b, jn_b, np.where(d[0] > 0, d[0], 1.0), d[1])
        worst = np.maximum(worst, e)
        assert e[0] <= 1e-3 and e[1] <= 0.05, (a, np.rad2deg(th), e, bad)
        assert e[2] <= 1e-2 and e[3] <= 0.15, (a, np.rad2deg(th), e, bad)
    print("off-node worst over 32 (a, i): r >= 1.5 risco flux %.2e dc %.4f | from the ISCO row flux %.2e dc %.4f"
          % tuple(worst))
    upload_golden_tables(ctx)


def test_hot_spectral_shape_attribute(ctx):
    """Fnu_seed_hot (relagn.py:1200-1223), a public attribute of the reference's class: through
    rg_last_hot_shape for every golden relagn / relqso set, and as the attribute of the class mirror."""
    d = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    ctx.set_grid(d["grid1000"])
    for tag, model in (("relagn", 0), ("relqso", 1)):
        ctx.eval_batch_host(d[tag + "_params"], model=model, rel=False)
        for i in range(len(d[tag + "_names"])):
            ref = d[tag + "_fnu_seed_hot1000"][i]
            got = ctx.last_hot_shape(i)
            if ref.max() == 0:
                assert got.max() == 0, (tag, i)
            else:
                assert peak_relerr(got, ref) < TOL, (tag, str(d[tag + "_names"][i]))
    import relagn_b200
    m = relagn_b200.relagn()
    m.new_ear(d["grid1000"])
    assert peak_relerr(m.Fnu_seed_hot, d["relagn_fnu_seed_hot1000"][0]) < TOL


def test_invalid_sets_are_flagged(ctx):
    ctx.set_grid(np.geomspace(1e-4, 1e3, 1001))
    p = c4_params(4, seed=3)
    p[1, 3] = 0.9999   # |a| > 0.998 -> ValueError in the reference (relagn.py:372-377)
    p[2, 4] = 0.05     # cos_inc < 0.09 (relagn.py:380-386)
    out, at = ctx.eval_batch_host(p, model=0, rel=False)
    assert list(at[:, 21]) == [0, 1, 2, 0]
    assert (out[1] == 0).all() and (out[2] == 0).all() and out[0].max() > 0
    q = np.zeros((2, 15))
    q[:, :7] = [[1e8, 100, -2.0, 0, 0.5, 1, 0], [1e8, 100, -1.0, 0, 0.5, 1, 0]]
    out, at = ctx.eval_batch_host(q, model=1, rel=False)   # relqso mdot bounds (relagn.py:1861-1874)
    assert list(at[:, 21]) == [3, 0]


def test_kyconv_seam_blurr_config(ctx, oracle, otab):
    """tests/blurr_test.py: narrow Gaussian line through kyconv, a=0.998 (config 2)."""
    upload_golden_tables(ctx)
    eb = np.geomspace(0.1, 2.0, 501)
    ec = 0.5 * (eb[1:] + eb[:-1])
    line = np.exp(-0.5 * ((ec - 1.0) / 0.01) ** 2) / (0.01 * np.sqrt(2 * np.pi)) * np.diff(eb)
    for inc in (30.0, 60.0, 75.0):
        par = [0.998, inc, 1.236971, 0, 100.0, 3.0, 3.0, 10.0, 0, 0, 500, -1]
        ref = oracle.kyconv(otab, eb, par, line)
        got = ctx.kyconv(eb, par, line.copy())
        assert peak_relerr(got, ref) < TOL
        # annulus additivity (blurr_test.py:77-139): sum over annuli with r^-3 weights ~ whole disc
        edges = np.geomspace(1.236971, 100.0, 200)
        acc = np.zeros_like(line)
        for r1, r2 in zip(edges[:-1], edges[1:]):
            rm = np.sqrt(r1 * r2)
            acc += rm ** -3 * ctx.kyconv(eb, [0.998, inc, r1, 0, r2, 0, 0, rm, 0, 0, 500, -1], line.copy())
        assert abs(acc.sum() / got.sum() - 1) < 2e-2
    k0, H = ctx.ky_kernel(0.0, np.deg2rad(60.0), 900.0, 950.0, np.log(1e7) / 1000)
    k1, H1 = oracle.ky_kernel(otab, 0.0, np.deg2rad(60.0), 900.0, 950.0, np.log(1e7) / 1000)
    assert k0 == k1 and np.allclose(H, H1, rtol=1e-10, atol=1e-12 * H1.max())


def test_kyconv_config2_inclination_sweep(ctx, oracle):
    """BASELINE config 2 as SURVEY 8(d) C2 specifies it (tests/blurr_test.py:23-54, 77-139): Gaussian line at 1 keV
    (sigma 0.01 keV) on 500 log bins 0.1-2 keV, a = 0.998, r_in = 1.236, r_out = 100, over the inclination sweep
    cos i = 0.09 ... 0.998, on the DEFAULT tables.  (i) one call with emissivity alpha = beta = 3, r_break = 10 against
    the oracle; (ii) 199 log-spaced annuli with alpha = beta = 0 and explicit r_mid^-3 weights summed: the additivity
    the RELAGN design rests on.  (ii) is a coarser quadrature of the same r^-3 integral (midpoint weights over annuli of
    0.022 dex), so it agrees with (i) to the quadrature error, not to rounding: 1 % of the line flux, 2 % in the L1
    norm of the profile."""
    from relagn_b200 import tables as T
    T.load_default(ctx)
    cfg = T.DEFAULT
    otab = oracle.Tables(cfg["spins"], cfg["thetas"], cfg["NR"], cfg["NPHI"], cfg["r_in"], data=ctx.tables_download())
    eb = np.geomspace(0.1, 2.0, 501)
    ec = 0.5 * (eb[1:] + eb[:-1])
    line = np.exp(-0.5 * ((ec - 1.0) / 0.01) ** 2) / (0.01 * np.sqrt(2 * np.pi)) * np.diff(eb)
    edges = np.geomspace(1.236, 100.0, 200)
    for ci in (0.09, 0.2, 0.3, 0.5, 0.7, 0.9, 0.998):
        inc = float(np.rad2deg(np.arccos(ci)))
        par = [0.998, inc, 1.236, 0, 100.0, 3.0, 3.0, 10.0, 0, 0, 500, -1]
        ref = oracle.kyconv(otab, eb, par, line)
        got = ctx.kyconv(eb, par, line.copy())
        assert peak_relerr(got, ref) < TOL, ci
        acc = np.zeros_like(line)
        for r1, r2 in zip(edges[:-1], edges[1:]):
            rm = 0.5 * (r1 + r2)  # blurr_test.py:89
            acc += rm ** -3 * ctx.kyconv(eb, [0.998, inc, r1, 0, r2, 0, 0, rm, 0, 0, 500, -1], line.copy())
        assert abs(acc.sum() / got.sum() - 1) < 1e-2, (ci, acc.sum() / got.sum())
        assert np.abs(acc - got).sum() / got.sum() < 2e-2, (ci, np.abs(acc - got).sum() / got.sum())
        # unit bookkeeping (blurr_test.py:148-150): the kernel of the whole disc integrates the r^-3 emissivity times
        # the projected area: positive, finite, and the redshifted wing carries flux below the line energy
        assert np.isfinite(got).all() and got.min() >= 0 and got[ec < 0.95].sum() > 0
    upload_golden_tables(ctx)


def test_limb_darkening_tables(ctx, oracle):
    """kyconv parameter 10 (relagn.py:988): the device ray tracer with the law baked in (rg_tables_set_limb before
    rg_tables_generate) against the oracle's, and the seam's check of the parameter against the resident set."""
    NR, NPHI, r_in = 32, 32, 2.0
    a, th = np.array([0.9]), np.array([np.deg2rad(60.0)])
    rows = np.array([oracle.lib().ro_row_radius(r_in, NR, s) for s in range(NR)])
    try:
        for law in (-1, -2):
            ctx.tables_set_limb(law)
            bad = ctx.tables_generate(a, th, NR, NPHI, r_in)
            d = ctx.tables_download()[0, 0]
            ref = oracle.make_node_full(0.9, th[0], rows, NPHI, limb=law)
            assert bad == 0 and ref["unresolved"] == 0
            np.testing.assert_allclose(d[0], ref["g"], rtol=2.4e-7)
            np.testing.assert_allclose(d[1], ref["jn"], rtol=2e-5)
            eb = np.geomspace(0.1, 10.0, 201)
            line = np.zeros(200)
            line[100] = 1.0
            par = [0.9, 60.0, 6.0, 0, 20.0, 0, 0, 10.0, 0, float(law), 200, -1]
            out = ctx.kyconv(eb, par, line.copy())
            assert out.sum() > 0
            with pytest.raises(Exception):
                ctx.kyconv(eb, par[:9] + [0.0] + par[10:], line.copy())   # isotropic asked, Laor / Haardt resident
    finally:
        ctx.tables_set_limb(0)
        upload_golden_tables(ctx)


def test_full_size_properties_relqso_grid(ctx):
    """Config 3 at full size (1e4 sets): properties that need no oracle."""
    grid = np.geomspace(1e-4, 1e3, 1001)
    ctx.set_grid(grid)
    upload_golden_tables(ctx)
    p = c3_params(100)
    out, at = ctx.eval_batch_host(p, model=1, rel=True, components=True)
    assert np.isfinite(out).all()
    assert (at[:, 21] == 0).all()
    assert (out >= 0).all()
    # total = disc + warm + hot
    np.testing.assert_allclose(out[:, 3], out[:, 0] + out[:, 1] + out[:, 2], rtol=1e-14)
    # batch-order invariance, bit for bit (no atomics anywhere in the path: the summation order is fixed)
    perm = np.random.default_rng(0).permutation(p.shape[0])[:512]
    out2, _ = ctx.eval_batch_host(p[perm], model=1, rel=True, components=True)
    np.testing.assert_array_equal(out2, out[perm])
    # relativistic and flat-space bolometric luminosities agree within a factor of order unity
    outn, atn = ctx.eval_batch_host(p[::97], model=1, rel=False, components=True)
    nu = (grid[:-1] + 0.5 * np.diff(grid)) * 1.602176634e-16 / 6.62607015e-34
    Lflat = np.trapezoid(outn[:, 3], nu, axis=1)
    Lrel = np.trapezoid(out[::97, 3], nu, axis=1)
    assert (Lflat > 0).all()
    assert np.all((Lrel / Lflat > 0.5) & (Lrel / Lflat < 1.5))


def test_full_size_c4_batch_chunking_invariance():
    """Config 4 at the per-GPU size of the metric (125 000 sets): the spectra do not depend on how the batch is cut
    into chunks / host slabs, bit for bit -- a small batch, then the full batch through the host-slab pipeline (the
    context's arenas grow from the 4096-set to the 65 536-set chunk in between), then the device-resident path."""
    import torch
    from relagn_b200 import _capi
    from relagn_b200.batch import BatchEvaluator
    grid = np.geomspace(1e-4, 1e3, 1001)
    t = golden_tables()
    ev = BatchEvaluator(grid, model="relagn", device=0, tables=t)
    p = c4_params(125000, seed=11)
    small, at_small = ev.evaluate(p[:300])
    small, at_small = small.copy(), at_small.copy()
    full, at_full = ev.evaluate(p)                      # chunks of 65 536 + 59 464 sets, copied out in pieces of 16 384
    full, at_full = full.copy(), at_full.copy()
    assert full.shape == (125000, 1000)
    ok = at_full[:, _capi.A["status"]] == 0
    assert ok.mean() > 0.9 and np.isfinite(full[ok]).all() and (full[ok] >= 0).all()
    np.testing.assert_array_equal(full[:300], small)
    np.testing.assert_array_equal(at_full[:300], at_small)
    # device-resident path: chunks of 65 536 + 59 464 sets
    dev, _ = ev.evaluate_device(torch.from_numpy(p).to("cuda:0"))
    np.testing.assert_array_equal(dev.cpu().numpy(), full)
    # the same with the completion hook (rg_eval_batch_pieces): contiguous pieces that cover the batch, same bits
    pieces = []
    dev2, _ = ev.evaluate_device_pieces(torch.from_numpy(p).to("cuda:0"), lambda lo, n: pieces.append((lo, n)),
                                        piece_sets=16384)
    torch.cuda.synchronize()
    assert pieces[0][0] == 0 and sum(n for _, n in pieces) == 125000
    assert all(a[0] + a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    assert max(n for _, n in pieces) <= 16384 + 8192 and len(pieces) >= 7
    np.testing.assert_array_equal(dev2.cpu().numpy(), full)
    # a slice from the middle of the batch, evaluated on its own
    mid, _ = ev.evaluate(p[70000:70300])   # (>= 2 x #SM sets: below that a set is split over several CTAs)
    np.testing.assert_array_equal(mid, full[70000:70300])


def test_python_default_grid_rel(ctx, oracle, otab):
    """The Python class's own default grid (1999 bins, 1e-4..1e4 keV): the NC = 8 instantiation, relativistic."""
    grid = np.geomspace(1e-4, 1e4, 2000)
    ctx.set_grid(grid)
    upload_golden_tables(ctx)
    p = np.array([[1e8, 100, -1, 0, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0],
                  [1e7, 100, -0.5, 0.9, 0.8, 150, 0.3, 1.9, 2.5, 8, 30, -1, 1, 8, 0]])
    st, ref, rat = oracle.evaluate_batch(p, grid, model=0, rel=True, tab=otab)
    out, at = ctx.eval_batch_host(p, model=0, rel=True, components=True)
    for i in range(2):
        for c in range(4):
            assert peak_relerr(out[i, c], ref[i, c]) < TOL, (i, c)


def test_hires_config5(ctx, oracle, otab):
    """BASELINE config 5: 5000 energy bins x ~2000 annuli (dr_dex = 700, log_rout = 3), a = 0.998 -- the NC = 20
    instantiation with the annuli of one set split over several CTAs."""
    grid = np.geomspace(1e-4, 1e3, 5001)
    ctx.set_grid(grid)
    upload_golden_tables(ctx)
    p = np.array([[1e8, 100, -1, 0.998, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, 3, 1, 10, 0],
                  [1e8, 100, -1, 0.998, 0.9, 100, 0.2, 1.7, 2.7, 10, 20, 3, 1, 10, 0]])
    st, ref, rat = oracle.evaluate_batch(p, grid, model=0, rel=True, tab=otab, dr_dex=700.0)
    out, at = ctx.eval_batch_host(p, model=0, rel=True, components=True, dr_dex=700.0)
    assert list(at[0, 17:20]) == [1189, 210, 635]
    for i in range(2):
        for c in range(4):
            assert peak_relerr(out[i, c], ref[i, c]) < TOL, (i, c)
    tot, _ = ctx.eval_batch_host(p, model=0, rel=True, components=False, dr_dex=700.0)
    np.testing.assert_allclose(tot, out[:, 3], rtol=1e-12)


def test_edge_grids(ctx, oracle, otab):
    """Grids the reference accepts through new_ear (relagn.py:780-805): a grid reaching below 1e-4 keV switches on the
    low-frequency extension of disc_annuli (:868-873); a non-geometric grid is fine without the transfer and is
    refused with it (kyconv restatement needs log-spaced bins); a tiny grid; an empty batch."""
    from relagn_b200 import _capi
    upload_golden_tables(ctx)
    p = np.array([[1e8, 100, -1, 0, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0],
                  [1e6, 100, -0.3, 0.5, 0.7, 80, 0.3, 2.0, 2.4, 12, 40, -1, -1, 10, 0]])
    grid = np.geomspace(1e-5, 1e2, 701)
    ctx.set_grid(grid)
    for rel in (False, True):
        st, ref, rat = oracle.evaluate_batch(p, grid, model=0, rel=rel, tab=otab)
        out, at = ctx.eval_batch_host(p, model=0, rel=rel, components=True)
        for i in range(2):
            for c in range(4):
                assert peak_relerr(out[i, c], ref[i, c]) < TOL, ("ext", rel, i, c)
    g2 = np.concatenate([np.linspace(1e-3, 1e-1, 200, endpoint=False), np.geomspace(1e-1, 1e2, 300)])
    ctx.set_grid(g2)
    st, ref, rat = oracle.evaluate_batch(p, g2, model=0, rel=False)
    out, at = ctx.eval_batch_host(p, model=0, rel=False, components=True)
    for i in range(2):
        for c in range(4):
            assert peak_relerr(out[i, c], ref[i, c]) < TOL, ("nongeo", i, c)
    with pytest.raises(_capi.RelagnError) as ei:
        ctx.eval_batch_host(p, model=0, rel=True)
    assert ei.value.code == -4  # RG_E_GRID
    g3 = np.geomspace(0.1, 10, 9)
    ctx.set_grid(g3)
    st, ref, rat = oracle.evaluate_batch(p, g3, model=0, rel=True, tab=otab)
    out, at = ctx.eval_batch_host(p, model=0, rel=True, components=True)
    for i in range(2):
        for c in range(4):
            assert peak_relerr(out[i, c], ref[i, c]) < TOL, ("tiny", i, c)
    out0, at0 = ctx.eval_batch_host(np.zeros((0, 15)), model=0, rel=True)
    assert out0.shape == (0, 8) and at0.shape == (0, 24)


@pytest.mark.parametrize("grid_case", ["wide_10_decades", "hard_band_above_1keV", "soft_band_below_1keV",
                                       "above_the_warm_x_grid", "below_every_x_grid"])
def test_band_grids_large_batch_vs_oracle(ctx, oracle, otab, monkeypatch, grid_case):
    """The large-batch path (200 sets: forward sweep per thread, back substitution as a warp scan, interpolation lines in
    shared memory) on grids that move the Kompaneets row bookkeeping and the scale of the local spectra: a 10-decade
    grid (wider than a warp's table -> tables in global scratch; below 1e-4 keV -> low-frequency extension), a band
    that starts above 1 keV (normalisation rows below the grid's first cell; cold annuli show only their far Wien
    tail and norm / radiance leaves the double range) and one that ends below 1 keV."""
    grid = {"wide_10_decades": np.geomspace(1e-6, 1e4, 1201), "hard_band_above_1keV": np.geomspace(2.0, 150.0, 401),
            "soft_band_below_1keV": np.geomspace(1e-3, 0.5, 301),
            # every bin above the top of the warm Comptonisation grids / below the first row of every Kompaneets grid
            "above_the_warm_x_grid": np.geomspace(45.0, 400.0, 101),
            "below_every_x_grid": np.geomspace(1e-11, 3e-10, 65)}[grid_case]
    upload_golden_tables(ctx)
    t = golden_tables()
    p = c4_params(200, seed=77)
    p[:, 3] = np.clip(p[:, 3], t["spins"][0], t["spins"][-1])
    ctx.set_grid(grid)
    n_ref = 24
    for rel in (False, True):
        st, ref, _ = oracle.evaluate_batch(p[:n_ref], grid, model=0, rel=rel, tab=otab)
        out, at = ctx.eval_batch_host(p, model=0, rel=rel, components=True)
        ok = at[:, 21] == 0
        assert np.isfinite(out[ok]).all(), (grid_case, rel)
        for k in range(n_ref):
            if st[k] != 0:
                continue
            for c in range(4):
                assert peak_relerr(out[k, c], ref[k, c]) < TOL, (grid_case, rel, k, c)
        monkeypatch.setenv("RG_INTERP_GLOBAL", "1")  # the wide-grid variant of the interpolation kernel: same bits
        out2, _ = ctx.eval_batch_host(p, model=0, rel=rel, components=True)
        monkeypatch.delenv("RG_INTERP_GLOBAL")
        np.testing.assert_array_equal(out2, out)


def test_class_api_on_device():
    """The class API end to end on the GPU: default relagn / relqso objects against the reference goldens."""
    from relagn_b200 import relagn, relqso
    from relagn_b200 import model
    t = golden_tables()
    model.set_tables(t)
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    dr = np.load(os.path.join(GOLDEN, "golden_rel.npz"))
    u = np.load(os.path.join(GOLDEN, "golden_units.npz"))
    o = relagn()
    assert o.Egrid.size == 1999 and (o.risco, len(o.logr_ad_bins), len(o.logr_wc_bins), len(o.logr_hc_bins)) == (6.0, 80, 16, 12)
    o.set_units("SI")
    assert peak_relerr(o.get_totSED(rel=False), dn["relagn_specdef"][0][3]) < TOL
    o.new_ear(dn["grid1000"])
    assert peak_relerr(o.get_totSED(rel=True), dr["relagn_spec1000"][0][3]) < TOL
    assert peak_relerr(o.get_DiscComponent(rel=True), dr["relagn_spec1000"][0][0]) < TOL
    o.set_units("counts")
    o.set_flux()
    assert peak_relerr(o.get_totSED(rel=False), u["counts_1"]) < TOL
    q = relqso()
    assert q.r_h == 14.261351436104999 and q.r_w == 28.522702872209997
    assert q.gamma_h == pytest.approx(1.9803689563797076, rel=1e-12)
    with pytest.raises(ValueError, match="not physical"):
        relagn(a=1.2)


def test_fp32_mode_stated_bound(ctx, oracle, otab):
    """RG_FLAG_FP32: single-precision local spectra / convolution / accumulation.  Stated bound: every bin within
    20 decades of the spectral peak agrees with the FP64 oracle to 1e-5 relative (measured ~6e-6; 3e-6 within 12
    decades); fainter Wien-tail bins lose precision and flush to zero below ~1e-35 of the peak."""
    grid = np.geomspace(1e-4, 1e3, 1001)
    ctx.set_grid(grid)
    upload_golden_tables(ctx)
    p = c4_params(64, seed=5)
    p[3, 11] = 3.6; p[3, 0] = 1e6
    p[5, 9] = -1
    p[6, 10] = p[6, 9]
    p[7, 12] = -1
    p[8, 0] = 1e10
    p[9, 0] = 1e5
    for rel in (True, False):
        st, ref, rat = oracle.evaluate_batch(p, grid, model=0, rel=rel, tab=otab)
        out, at = ctx.eval_batch_host(p, model=0, rel=rel, components=True, fp32=True)
        assert np.isfinite(out).all()
        for i in range(p.shape[0]):
            for c in range(4):
                assert peak_relerr(out[i, c], ref[i, c], floor=1e-20) < 1e-5, (rel, i, c)


def test_observation_epilogue(ctx, oracle, otab):
    """SURVEY 8 f1/f2 on the device, through the C ABI with torch-owned buffers: unit conversion bit-identical to the
    oracle (and <= 1e-14 from the unmodified reference's getters), XSPEC-contract photar bit-identical, folded counts
    bit-identical, statistic 1e-12; parameters -> statistic in one call on a C4 batch, relativistic."""
    import torch
    from tests.util import toy_response
    O = oracle
    dn = np.load(os.path.join(GOLDEN, "golden_nonrel.npz"))
    u = np.load(os.path.join(GOLDEN, "golden_units.npz"))
    eb = dn["grid1000"]
    ctx.set_grid(eb)
    sel = [0, 6, 8, 15]
    P = len(sel)
    par = np.ascontiguousarray(dn["relagn_params"][sel])
    L = np.ascontiguousarray(dn["relagn_spec1000"][sel][:, 3, :])
    d_par = torch.from_numpy(par).cuda()
    d_L = torch.from_numpy(L).cuda()
    for unit in ["SI", "cgs", "cgs_wave", "SI_wave", "counts"]:
        for fl in (False, True):
            d_out = torch.empty_like(d_L)
            ctx.convert_device(d_L.data_ptr(), d_par.data_ptr(), P, 1, unit, fl, d_out.data_ptr())
            out = d_out.cpu().numpy()
            for k in range(P):
                assert np.array_equal(out[k], O.convert(L[k], eb, unit, fl, par[k, 1], par[k, 14])), (unit, fl, k)
            assert peak_relerr(out[0], u["%s_%d" % (unit, fl)]) < 1e-14
    for ear in (np.linspace(0.3, 10.0, 257), np.geomspace(5e-5, 3e3, 65), np.array([2e3, 3e3, 4e3]), eb[::7].copy(),
                eb / 1.5):
        ctx.set_instrument(ear)
        d_ph = torch.empty((P, ear.size - 1), dtype=torch.float64, device="cuda")
        ctx.photar_device(d_L.data_ptr(), d_par.data_ptr(), P, d_ph.data_ptr())
        ph = d_ph.cpu().numpy()
        for k in range(P):
            assert np.array_equal(ph[k], O.photar(L[k], eb, par[k, 1], par[k, 14], ear)), k
    ear = np.geomspace(0.3, 10.0, 201)
    rp, cl, vl = toy_response(ear)
    expo = 5e4
    ctx.set_instrument(ear)
    ctx.set_response(rp, cl, vl, expo)
    pho = np.stack([O.photar(L[k], eb, par[k, 1], par[k, 14], ear) for k in range(P)])
    _, folded0 = O.fold_stat(pho[0], rp, cl, vl, expo, np.zeros(rp.size - 1), None, 1)
    data = np.random.Generator(np.random.PCG64(3)).poisson(np.maximum(folded0, 0)).astype(float)
    sig = np.sqrt(np.maximum(data, 1.0))
    for stat, kind in (("chi2", 0), ("cstat", 1)):
        ctx.set_data(data, sig if kind == 0 else None, stat)
        d_st = torch.empty(P, dtype=torch.float64, device="cuda")
        d_fo = torch.empty((P, rp.size - 1), dtype=torch.float64, device="cuda")
        ctx.fitstat_device(d_L.data_ptr(), d_par.data_ptr(), P, d_st.data_ptr(), d_fo.data_ptr())
        st, fo = d_st.cpu().numpy(), d_fo.cpu().numpy()
        for k in range(P):
            s_ref, f_ref = O.fold_stat(pho[k], rp, cl, vl, expo, data, sig if kind == 0 else None, kind)
            assert np.array_equal(fo[k], f_ref)
            assert st[k] == pytest.approx(s_ref, rel=1e-12)
    # parameters -> Cash statistic in one call (relativistic, C4 draws, two slabs' worth of chunking not needed here)
    upload_golden_tables(ctx)
    t = golden_tables()
    p = c4_params(300)
    p[:, 3] = np.clip(p[:, 3], t["spins"][0], t["spins"][-1])
    p[:, 14] = np.linspace(0, 0.8, 300)  # redshifts
    st = ctx.eval_fitstat_host(p, model=0, rel=True)
    spec, at = ctx.eval_batch_host(p, model=0, rel=True)
    for k in range(300):
        s_ref, _ = O.fold_stat(O.photar(spec[k], eb, p[k, 1], p[k, 14], ear), rp, cl, vl, expo, data, None, 1)
        assert st[k] == pytest.approx(s_ref, rel=1e-12), k
    # ... and against the oracle's own spectra (tolerance of the spectrum path)
    _, osp, _ = O.evaluate_batch(p[:40], eb, model=0, rel=True, tab=otab)
    for k in range(40):
        s_ref, _ = O.fold_stat(O.photar(osp[k, 3], eb, p[k, 1], p[k, 14], ear), rp, cl, vl, expo, data, None, 1)
        assert st[k] == pytest.approx(s_ref, rel=1e-8), k


def test_batch_evaluator_fitstat_api():
    """The host-facing call of the fit / MCMC inner loop: BatchEvaluator.fitstat == evaluate + host epilogue."""
    import torch
    from relagn_b200.batch import BatchEvaluator
    from oracle import oracle as O
    from tests.util import toy_response
    eb = np.geomspace(1e-4, 1e3, 1001)
    t = golden_tables()
    ev = BatchEvaluator(eb, model="relagn", tables=t)
    ear = np.geomspace(0.3, 10.0, 201)
    rp, cl, vl = toy_response(ear)
    ev.set_instrument(ear, rp, cl, vl, exposure=2e4)
    p = c4_params(20000, seed=5)  # more than one 16384-set slab
    p[:, 3] = np.clip(p[:, 3], t["spins"][0], t["spins"][-1])
    p[:, 14] = 0.1
    spec, _ = ev.evaluate(p[:64], rel=True)
    _, f0 = O.fold_stat(O.photar(spec[0], eb, p[0, 1], p[0, 14], ear), rp, cl, vl, 2e4, np.zeros(rp.size - 1), None, 1)
    data = np.round(f0)
    ev.set_data(data, np.sqrt(np.maximum(data, 1.0)), "chi2")
    st = ev.fitstat(p, rel=True).copy()
    assert st.shape == (20000,) and np.all(np.isfinite(st))
    for k in range(64):
        s_ref, _ = O.fold_stat(O.photar(spec[k], eb, p[k, 1], p[k, 14], ear), rp, cl, vl, 2e4, data,
                               np.sqrt(np.maximum(data, 1.0)), 0)
        assert st[k] == pytest.approx(s_ref, rel=1e-12)
    assert np.argmin(st[:64]) == 0
    # the same sets in another order and in other slabs give the same bits (batch-order independence)
    st_q = ev.fitstat(p[::-1].copy(), rel=True)
    assert np.array_equal(st_q[::-1], st)
    # device-resident variant + in-place unit conversion of a spectra batch
    d_p = torch.from_numpy(p[:64]).cuda()
    d_st = ev.fitstat_device(d_p, rel=True)
    # (a 64-set batch splits every set's annuli over several CTAs: same numbers to rounding, not the same bits)
    np.testing.assert_allclose(d_st.cpu().numpy(), st[:64], rtol=1e-11)
    d_spec, _ = ev.evaluate_device(d_p, rel=True)
    ref = O.convert(d_spec[3].cpu().numpy(), eb, "counts", True, p[3, 1], p[3, 14])
    ev.convert_device(d_spec, d_p, unit="counts", as_flux=True)
    assert np.array_equal(d_spec[3].cpu().numpy(), ref)


def test_xspec_local_model_abi(ctx, oracle, otab):
    """Seam B3: relagn_ / relqso_ with XSPEC's calling convention (real*4 by reference) against the oracle's statement
    of relagn.f:60-112; rg_tables_default from the package's .npy cache."""
    from tests.util import xspec_reference
    upload_golden_tables(ctx)
    ctx.xspec_bind()
    try:
        for ear in (np.linspace(0.3, 10.0, 1025), np.geomspace(0.1, 200.0, 301),
                    np.concatenate([[0.0], np.linspace(0.05, 12.0, 500)])):
            ear = ear.astype(np.float32)
            par = np.array([1e8, 100, -1, 0.7, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0.05], dtype=np.float32)
            got = ctx.xspec_model("relagn", ear, par)
            ref = xspec_reference(oracle, otab, 0, True, ear.astype(np.float64), par.astype(np.float64))
            np.testing.assert_allclose(got, ref.astype(np.float32), rtol=3e-7)
            assert np.array_equal(ctx.xspec_model("relagn", ear, par), got)
        for relsw in (0, 1):
            q = np.array([1e7, 10, -0.5, 0.0, 0.5, 0.02, relsw], dtype=np.float32)
            got = ctx.xspec_model("relqso", ear, q)
            p7 = np.array([q[0], q[1], q[2], q[3], q[4], 1.0, q[5]], dtype=np.float64)
            ref = xspec_reference(oracle, otab, 1, bool(relsw), ear.astype(np.float64), p7)
            np.testing.assert_allclose(got, ref.astype(np.float32), rtol=3e-7)
        # out-of-range spin: the reference raises; XSPEC has no error channel -> zeros
        bad = par.copy()
        bad[3] = 0.9999
        assert np.all(ctx.xspec_model("relagn", ear, bad) == 0)
    finally:
        ctx.lib.rg_xspec_bind(None)
    # default tables through the C entry point (the .npy cache of relagn_b200/_tables travels with the package)
    from relagn_b200 import tables as T
    import glob
    files = glob.glob(os.path.join(T.CACHE_DIR, "kerr_*.npy"))
    if files:
        ctx.tables_default(files[0])
        info = ctx.tables_info()
        assert (info["n_spin"], info["n_inc"], info["NR"], info["NPHI"]) == (14, 14, 64, 64)
        assert np.array_equal(ctx.tables_download(), np.load(files[0]))
    with pytest.raises(Exception):
        ctx.tables_default("/nonexistent/kerr.npy")
