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

# default node grid: spins denser towards 0.998 (risco changes fastest there), inclinations uniform in angle
DEFAULT = dict(
    spins=np.array([-0.998, -0.5, 0.0, 0.3, 0.5, 0.7, 0.8, 0.9, 0.95, 0.98, 0.998]),
    thetas=np.deg2rad(np.array([3.0, 10.0, 18.0, 26.0, 34.0, 42.0, 50.0, 57.0, 63.0, 69.0, 74.0, 78.0, 82.0, 85.0])),
    NR=64, NPHI=64, dl=3.76 / 63)


def _key(cfg):
    h = hashlib.sha1()
    for k in ("spins", "thetas"):
        h.update(np.ascontiguousarray(cfg[k], dtype=np.float64).tobytes())
    h.update(("%d|%d|%.17g" % (cfg["NR"], cfg["NPHI"], cfg["dl"])).encode())
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
        ctx.tables_upload(cfg["spins"], cfg["thetas"], cfg["NR"], cfg["NPHI"], cfg["dl"], data)
        return 0
    bad = ctx.tables_generate(cfg["spins"], cfg["thetas"], cfg["NR"], cfg["NPHI"], cfg["dl"])
    if cache:
        try:
            os.makedirs(CACHE_DIR, exist_ok=True)
            np.save(path, ctx.tables_download())
        except OSError:
            pass
    return bad
