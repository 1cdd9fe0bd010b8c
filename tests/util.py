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
                dl=float(d["tab_dl"]), data=d["tab_data"])


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
