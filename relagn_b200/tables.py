"""Kerr transfer tables: default node grid, on-device generation and the on-disk cache.

The reference gets these numbers from HEASOFT's KYCONV FITS tables (not in the reference repository, not
installable offline: relagn.py:979-1001 calls xspec 'kyconv').  Here they are ray-traced on the GPU by
rg_tables_generate (csrc/rg_raytrace.cu) and cached as .npy next to the package.
"""
import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CACHE_DIR = os.path.join(HERE, "_tables")

# default node grid (table format 2: rows are physical radii shared by all nodes, cubic Lagrange lookup in a and
# theta).  Spins denser towards 0.998: neighbouring nodes must both hold circular orbits at the ISCO of any model
# between them; inclinations denser towards edge-on, where lensing of the far side sharpens.  Measured off-node
# error of this grid (tools/table_error.py, 40 random (a, i)): ring flux <= 2e-4, centroid <= 0.006 bin of the
# 1000-bin grid for r >= 1.5 r_isco; <= 1.6e-3 / 0.034 bin at the ISCO row itself.
DEFAULT = dict(
    spins=np.array([-0.998, -0.6, -0.2, 0.1, 0.3, 0.5, 0.7, 0.8, 0.9, 0.95, 0.975, 0.9875, 0.995, 0.998]),
    thetas=np.deg2rad(np.array([0.0, 10.0, 20.0, 30.0, 40.0, 50.0, 57.5, 65.0, 70.0, 75.0, 77.5, 80.0, 82.5, 85.0])),
    NR=64, NPHI=64, r_in=1.2369)


def _key(cfg):
    h = hashlib.sha1()
    for k in ("spins", "thetas"):
        h.update(np.ascontiguousarray(cfg[k], dtype=np.float64).tobytes())
    h.update(("v2|%d|%d|%.17g" % (cfg["NR"], cfg["NPHI"], cfg["r_in"])).encode())
    return h.hexdigest()[:16]


def cached_default(cfg=None):
    """The cached table array of `cfg` (default: DEFAULT) or None."""
    cfg = DEFAULT if cfg is None else cfg
    path = os.path.join(CACHE_DIR, "kerr_%s.npy" % _key(cfg))
    return np.load(path) if os.path.exists(path) else None


def load_default(ctx, cfg=None, cache=True):
    """Make `cfg` (default: DEFAULT) resident in ctx: from the disk cache if present, else ray-trace on the
    device and store."""
    cfg = DEFAULT if cfg is None else cfg
    path = os.path.join(CACHE_DIR, "kerr_%s.npy" % _key(cfg))
    if cache and os.path.exists(path):
        data = np.load(path)
        ctx.tables_upload(cfg["spins"], cfg["thetas"], cfg["NR"], cfg["NPHI"], cfg["r_in"], data)
        return 0
    bad = ctx.tables_generate(cfg["spins"], cfg["thetas"], cfg["NR"], cfg["NPHI"], cfg["r_in"])
    if cache:
        try:
            os.makedirs(CACHE_DIR, exist_ok=True)
            np.save(path, ctx.tables_download())
        except OSError:
            pass
    return bad
