"""Shared helpers of the test-suite."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

RELAGN_DEF = [1e8, 100, -1, 0, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0]
RELQSO_DEF = [1e5, 1, -1, 0, 0.5, 1, 0]


def peak_relerr(a, b, floor=1e-12):
    """max |a-b| / max(|b|, floor*peak(b)).  The reference's spectra span 300 decades (Wien tails down to
    denormals); bins below floor*peak carry no flux and are compared on the absolute scale of the peak."""
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    pk = np.max(np.abs(b))
    if pk == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * pk)))


def golden_tables():
    d = np.load(os.path.join(GOLDEN, "golden_rel.npz"))
    return dict(spins=d["tab_spins"], thetas=d["tab_thetas"], NR=int(d["tab_NR"]), NPHI=int(d["tab_NPHI"]),
                r_in=float(d["tab_r_in"]), data=d["tab_data"])


def c4_params(P, seed=20240229, tab_spins=None, tab_thetas=None):
    """RELAGN full model, independent uniforms inside lmod_relagn.dat hard limits."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    from relagn_b200.host import risco as _risco
    M = 10 ** rng.uniform(5, 10, P)
    lm = rng.uniform(-1.5, 0.3, P)
    a = rng.uniform(0, 0.998, P)
    ci = rng.uniform(0.09, 0.998, P)
    kth = rng.uniform(10, 300, P)
    ktw = rng.uniform(0.1, 1, P)
    gh = rng.uniform(1.3, 3, P)
    gw = rng.uniform(2, 5, P)
    ri = _risco(a)
    rh = ri + rng.uniform(0.5, 100, P)
    rw = rh * (1 + rng.uniform(0, 4, P))
    hm = np.minimum(10.0, rh)
    p = np.stack([M, np.full(P, 100.0), lm, a, ci, kth, ktw, gh, gw, rh, rw, np.full(P, -1.0), np.ones(P), hm,
                  np.zeros(P)], axis=1)
    return np.ascontiguousarray(p)


def c3_params(n=100):
    """relqso mass x mdot grid (n x n sets)."""
    M = np.logspace(5, 10, n)
    lm = np.linspace(-1.6, 0.5, n)
    MM, LL = np.meshgrid(M, lm, indexing="ij")
    P = MM.size
    p = np.zeros((P, 15))
    p[:, 0] = MM.ravel()
    p[:, 1] = 100.0
    p[:, 2] = LL.ravel()
    p[:, 3] = 0.0
    p[:, 4] = 0.5
    p[:, 5] = 1.0
    p[:, 6] = 0.0
    return p


def toy_response(ear, n_chan=96, seed=7):
    """A small instrument for the observation-epilogue tests: Gaussian redistribution (sigma ~ 2.5 instrument bins,
    cut at 4 sigma) times a smooth effective area, CSR over channels; ragged rows incl. one empty channel."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ne = len(ear) - 1
    ec = 0.5 * (np.asarray(ear)[1:] + np.asarray(ear)[:-1])
    centre = np.linspace(0, ne - 1, n_chan)
    rowptr, col, val = [0], [], []
    for c in range(n_chan):
        if c == 5:
            rowptr.append(len(col))  # a dead channel
            continue
        sig = 1.5 + 2.0 * rng.random()
        lo, hi = int(max(0, np.floor(centre[c] - 4 * sig))), int(min(ne - 1, np.ceil(centre[c] + 4 * sig)))
        for i in range(lo, hi + 1):
            w = np.exp(-0.5 * ((i - centre[c]) / sig) ** 2) / (sig * np.sqrt(2 * np.pi))
            area = 400.0 * np.exp(-0.5 * (np.log(ec[i] / 1.5) / 1.2) ** 2)
            col.append(i)
            val.append(w * area)
        rowptr.append(len(col))
    return np.array(rowptr, dtype=np.int32), np.array(col, dtype=np.int32), np.array(val)


def xspec_reference(O, otab, model, rel, ear, par):
    """What relagn_ / relqso_ must return (relagn.f:60-112 with the Python class's physics), from the oracle:
    2000-bin model grid over [min(1e-4, ear0 (1+z)), max(1e3, earN (1+z))], evaluate, redshift + inibin/erebin."""
    ear = np.asarray(ear, dtype=np.float64)
    z = par[6] if model == 1 else par[14]
    lo = ear[1] - (ear[2] - ear[1]) / 10.0 if ear[0] == 0 else min(1e-4, ear[0] * (1 + z))
    hi = max(1e3, ear[-1] * (1 + z))
    enew = np.exp(np.log(lo) + np.log(hi / lo) / 2000 * np.arange(2001))
    enew[0], enew[-1] = lo, hi
    st, out, _ = O.evaluate(par, enew, model=model, rel=rel, tab=otab)
    assert st == 0
    return O.photar(out[3], enew, par[1], z, ear)


# ---- off-node accuracy of a Kerr table set (tests/test_tables.py, tests/test_gpu_parity.py, tools/table_error.py) ----
DLNE_1000 = np.log(1e7) / 1000.0  # bin width of the 1000-bin metric grid


def risco(a):
    a = np.asarray(a, dtype=np.float64)
    z1 = 1 + (1 - a * a) ** (1 / 3) * ((1 + a) ** (1 / 3) + (1 - a) ** (1 / 3))
    z2 = np.sqrt(3 * a * a + z1 * z1)
    return 3 + z2 - np.sign(a) * np.sqrt((3 - z1) * (3 + z1 + 2 * z2))


def ring_metrics(lg, w, dlnE=DLNE_1000):
    """Per table row of a slice (ln g [NR][NPHI], weight density w = Jn g^3): the ring's flux sum_t W_t dphi and the
    centroid of its shift histogram in energy bins, sum_t W_t (s_t + s_{t+1})/2 / sum W_t with W_t = (w_t + w_{t+1})/2,
    s = ln g / dlnE -- the two moments the hat deposit of the transfer kernel preserves exactly."""
    lg = np.asarray(lg, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    W = 0.5 * (w + np.roll(w, -1, axis=1))
    F = W.sum(axis=1) * (2 * np.pi / lg.shape[1])
    c = (W * 0.5 * (lg + np.roll(lg, -1, axis=1))).sum(axis=1) / W.sum(axis=1) / dlnE
    return F, c


def offnode_errors(rows, a, lg_blend, jn_blend, g_true, jn_true):
    """(worst |flux ratio - 1|, worst |centroid shift| in bins) over the rows r >= 1.5 r_isco and over all rows the
    model uses (from the row bracketing r_isco from below): blended slice against a node traced at the same (a, i)."""
    ri = float(risco(a))
    s_first = max(int(np.searchsorted(rows, ri, side="right")) - 1, 0)
    sl = slice(s_first, len(rows))
    gt = np.asarray(g_true, dtype=np.float64)[sl]
    F0, c0 = ring_metrics(np.log(gt), np.asarray(jn_true, dtype=np.float64)[sl] * gt ** 3)
    jb = np.maximum(np.asarray(jn_blend)[sl], 0.0)
    F1, c1 = ring_metrics(np.asarray(lg_blend)[sl], jb * np.exp(3 * np.asarray(lg_blend)[sl]))
    fe, ce = np.abs(F1 / F0 - 1), np.abs(c1 - c0)
    far = rows[sl] >= 1.5 * ri
    return fe[far].max(), ce[far].max(), fe.max(), ce.max()
