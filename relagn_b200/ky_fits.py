"""Importer / exporter for Kerr transfer tables in the layout of the KY package's `KBHtablesNN.fits`.

The reference reaches these numbers through `xspec.callModelFunction('kyconv', ...)` (relagn.py:979-1001): HEASOFT's
KYCONV interpolates g (energy shift), the lensing factor l, the cosine of the local emission angle mu_e (and alpha, beta,
the time delay) from a FITS file shipped with HEASOFT.  That file is not in the reference repository and cannot be
installed offline; the layout below is the one the KY documentation describes (Dovciak, Karas & Yaqoob 2004, ApJS 153,
205; `KBHtables80.fits`), restated from memory:

    HDU 1  BINTABLE, one column  r_horizon      horizon radii r_h of the tabulated spins (a = sqrt(1 - (r_h - 1)^2))
    HDU 2  BINTABLE, one column  inclination    observer inclinations, degrees
    HDU 3  BINTABLE, one column  r_vector       radii of the table rows as r - r_h
    HDU 4  BINTABLE, one column  phi_vector     Boyer-Lindquist azimuths of the table columns, radians
    HDU 5+ one BINTABLE per (r_horizon, inclination) pair, inclination fastest, n_r rows, columns with n_phi entries
           each:  alpha, beta, lensing, g-factor, cosine, delay

UNVERIFIED AGAINST A REAL HEASOFT FILE.  What is tested (tests/test_ky_fits.py) is the round trip through this layout
of tables ray-traced by the oracle / the device ray tracer; extension and column names are arguments of the importer,
so that a differing real file can be mapped without touching the code.

Conversion to the library's table format 2 (include/relagn_b200.h):
  * KY's weight per disc-plane area element is  g^2 l mu_e r dr dphi  (times the limb law f(mu_e)); with the identity
    l mu_e r dr dphi = g dalpha dbeta the library's  Jn = dalpha dbeta / (dX dY) = l mu_e / g,
    and a limb-darkening law (kyconv parameter 10: 0 isotropic, -1 Laor 1 + 2.06 mu_e, -2 Haardt ln(1 + 1/mu_e)) is
    baked in as Jn f(mu_e);
  * rows: resampled linearly in log r onto the library's physical radii (uniform in 1/r + 1/sqrt(r) from r_in to 1000);
  * columns: the library measures azimuth from the "front ray" (alpha = 0, beta < 0) of every row; its azimuth phi0(r)
    is found as the zero crossing of alpha along the row, and the columns are resampled (periodic, linear) onto
    phi_t = 2 pi t / NPHI + phi0(r).

No astropy here (not installed): a minimal FITS binary-table reader / writer for exactly this layout.
"""
import numpy as np

_BLOCK = 2880


# ------------------------------------------------------------------------------------------------
# minimal FITS
def _card(key, value=None, comment=""):
    if value is None:
        s = "%-8s" % key
    elif isinstance(value, bool):
        s = "%-8s= %20s" % (key, "T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        s = "%-8s= %20d" % (key, value)
    elif isinstance(value, float):
        s = "%-8s= %20.13E" % (key, value)
    else:
        s = "%-8s= %-20s" % (key, "'%-8s'" % value)
    if comment:
        s += " / " + comment
    return (s + " " * 80)[:80]


def _pad(b, fill=b"\0"):
    r = (-len(b)) % _BLOCK
    return b + fill * r


def _write_header(f, cards):
    f.write(_pad(("".join(cards) + _card("END")).encode("ascii"), b" "))


def _write_bintable(f, name, columns):
    """columns: list of (name, array [nrows, repeat]) -> one BINTABLE HDU, big-endian float32 ('E')."""
    nrows = columns[0][1].shape[0]
    reps = [c[1].reshape(nrows, -1).shape[1] for c in columns]
    width = 4 * sum(reps)
    cards = [_card("XTENSION", "BINTABLE"), _card("BITPIX", 8), _card("NAXIS", 2), _card("NAXIS1", width),
             _card("NAXIS2", nrows), _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(columns)),
             _card("EXTNAME", name)]
    for i, ((cname, _), rep) in enumerate(zip(columns, reps), 1):
        cards += [_card("TTYPE%d" % i, cname), _card("TFORM%d" % i, "%dE" % rep)]
    _write_header(f, cards)
    rec = np.concatenate([c[1].reshape(nrows, -1) for c in columns], axis=1).astype(">f4")
    f.write(_pad(rec.tobytes()))


def _read_header(f):
    cards = {}
    order = []
    while True:
        block = f.read(_BLOCK)
        if len(block) < _BLOCK:
            return None, None
        done = False
        for i in range(0, _BLOCK, 80):
            c = block[i:i + 80].decode("ascii", "replace")
            key = c[:8].strip()
            if key == "END":
                done = True
                break
            if c[8:10] == "= ":
                v = c[10:].split("/")[0].strip()
                if v.startswith("'"):
                    v = v.strip("'").strip()
                elif v in ("T", "F"):
                    v = v == "T"
                else:
                    try:
                        v = int(v)
                    except ValueError:
                        try:
                            v = float(v.replace("D", "E"))
                        except ValueError:
                            pass
                cards[key] = v
                order.append(key)
        if done:
            return cards, order


_TFORM = {"E": (">f4", 4), "D": (">f8", 8), "J": (">i4", 4), "I": (">i2", 2), "K": (">i8", 8)}


def read_fits_tables(path):
    """All BINTABLE HDUs of a FITS file: list of (EXTNAME, {column name: ndarray [nrows, repeat]})."""
    out = []
    with open(path, "rb") as f:
        hdr, _ = _read_header(f)  # primary
        if hdr is None or not hdr.get("SIMPLE", False):
            raise ValueError("%s is not a FITS file" % path)
        nbytes = abs(hdr.get("BITPIX", 8)) // 8
        n = 1 if hdr.get("NAXIS", 0) else 0
        for k in range(1, hdr.get("NAXIS", 0) + 1):
            n *= hdr["NAXIS%d" % k]
        f.seek((n * nbytes + _BLOCK - 1) // _BLOCK * _BLOCK, 1)
        while True:
            hdr, _ = _read_header(f)
            if hdr is None:
                break
            size = hdr.get("NAXIS1", 0) * hdr.get("NAXIS2", 0) + hdr.get("PCOUNT", 0)
            raw = f.read((size + _BLOCK - 1) // _BLOCK * _BLOCK)
            if hdr.get("XTENSION") != "BINTABLE":
                continue
            nrows, width = hdr["NAXIS2"], hdr["NAXIS1"]
            rec = np.frombuffer(raw[:nrows * width], dtype=np.uint8).reshape(nrows, width)
            cols, off = {}, 0
            for i in range(1, hdr["TFIELDS"] + 1):
                tf = str(hdr["TFORM%d" % i]).strip()
                rep = int(tf[:-1]) if tf[:-1] else 1
                dt, sz = _TFORM[tf[-1]]
                a = rec[:, off:off + rep * sz].copy().view(dt).reshape(nrows, rep)
                cols[str(hdr.get("TTYPE%d" % i, "col%d" % i)).strip()] = a.astype(np.float64 if dt[1] == "f" else np.int64)
                off += rep * sz
            out.append((str(hdr.get("EXTNAME", "")).strip(), cols))
    return out


# ------------------------------------------------------------------------------------------------
# KY layout
def write_ky_fits(path, r_horizon, inclination_deg, r_vector, phi_vector, nodes):
    """nodes[(i_h, i_inc)] = dict(alpha, beta, lensing, g, cosine, delay) of arrays [n_r][n_phi]."""
    with open(path, "wb") as f:
        _write_header(f, [_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True)])
        _write_bintable(f, "r_horizon", [("r_horizon", np.asarray(r_horizon, dtype=np.float64)[:, None])])
        _write_bintable(f, "inclination", [("inclination", np.asarray(inclination_deg, dtype=np.float64)[:, None])])
        _write_bintable(f, "r_vector", [("r_vector", np.asarray(r_vector, dtype=np.float64)[:, None])])
        _write_bintable(f, "phi_vector", [("phi_vector", np.asarray(phi_vector, dtype=np.float64)[:, None])])
        for ih in range(len(r_horizon)):
            for ii in range(len(inclination_deg)):
                n = nodes[(ih, ii)]
                zeros = np.zeros_like(n["g"], dtype=np.float64)
                _write_bintable(f, "tables_%d_%d" % (ih, ii),
                                [("alpha", n["alpha"]), ("beta", n["beta"]), ("lensing", n["lensing"]),
                                 ("g-factor", n["g"]), ("cosine", n["cosine"]), ("delay", n.get("delay", zeros))])


def limb_factor(law, mue):
    """kyconv parameter 10 (relagn.py:988 passes 0): 0 isotropic, -1 Laor (1991), -2 Haardt (1993)."""
    mue = np.asarray(mue, dtype=np.float64)
    if law == -1:
        return 1.0 + 2.06 * mue
    if law == -2:
        return np.log(1.0 + 1.0 / np.maximum(mue, 1e-300))
    if law == 0:
        return np.ones_like(mue)
    raise ValueError("limb law %r: 0 (isotropic), -1 (Laor) or -2 (Haardt)" % (law,))


def _row_radii(r_in, NR):
    x = lambda r: 1.0 / r + 1.0 / np.sqrt(r)
    xs = x(r_in) + np.arange(NR) * ((x(1000.0) - x(r_in)) / (NR - 1))
    u = 0.5 * (-1.0 + np.sqrt(1.0 + 4.0 * xs))
    r = 1.0 / (u * u)
    r[0], r[-1] = r_in, 1000.0
    return r


def _r_valid(a):
    return 1.02 * 2 * (1 + np.cos(2.0 / 3.0 * np.arccos(-a)))


def _front_azimuth(phi, alpha_row, beta_row):
    """azimuth of the ray alpha = 0, beta < 0 on one table row: zero crossing of alpha(phi) on the beta < 0 side"""
    n = len(phi)
    best = None
    for t in range(n):
        t1 = (t + 1) % n
        a0, a1 = alpha_row[t], alpha_row[t1]
        if a0 == 0 or a0 * a1 < 0:
            if 0.5 * (beta_row[t] + beta_row[t1]) < 0:
                f = 0.0 if a0 == a1 else a0 / (a0 - a1)
                dphi = (phi[t1] - phi[t]) % (2 * np.pi)
                best = phi[t] + f * dphi
                break
    return 0.0 if best is None else best


def import_ky_tables(path, NR=64, NPHI=64, r_in=1.2369, limb=0, names=None):
    """KY FITS file -> dict(spins, thetas, NR, NPHI, r_in, data [n_spin][n_inc][2][NR][NPHI] float32) for
    Context.tables_upload / BatchEvaluator(tables=...).  names: override the extension / column names, e.g.
    dict(g="g-factor", lensing="lensing", cosine="cosine", alpha="alpha", beta="beta")."""
    nm = dict(r_horizon="r_horizon", inclination="inclination", r_vector="r_vector", phi_vector="phi_vector",
              g="g-factor", lensing="lensing", cosine="cosine", alpha="alpha", beta="beta")
    nm.update(names or {})
    hdus = read_fits_tables(path)

    def vec(key):
        for _, cols in hdus:
            for cname, arr in cols.items():
                if cname.lower() == nm[key].lower() and arr.shape[1] == 1:
                    return arr[:, 0]
        raise ValueError("%s: no one-column table '%s'" % (path, nm[key]))

    rh, inc, rv, phi = vec("r_horizon"), vec("inclination"), vec("r_vector"), vec("phi_vector")
    tabs = [cols for _, cols in hdus if any(c.lower() == nm["g"].lower() for c in cols)]
    if len(tabs) != len(rh) * len(inc):
        raise ValueError("%s: %d transfer-function extensions for %d x %d (r_horizon, inclination) pairs" % (
            path, len(tabs), len(rh), len(inc)))

    def col(cols, key):
        for cname, arr in cols.items():
            if cname.lower() == nm[key].lower():
                return arr
        raise ValueError("%s: column '%s' missing" % (path, nm[key]))

    spins_all = np.sqrt(np.clip(1.0 - (rh - 1.0) ** 2, 0.0, 1.0))  # prograde tables (KY tabulates a >= 0)
    order_a = np.argsort(spins_all)
    order_i = np.argsort(inc)
    rows = _row_radii(r_in, NR)
    data = np.zeros((len(rh), len(inc), 2, NR, NPHI), dtype=np.float32)
    tcol = 2 * np.pi * np.arange(NPHI) / NPHI
    for ma, ih in enumerate(order_a):
        a = spins_all[ih]
        r_src = rh[ih] + rv
        lr_src = np.log(r_src)
        for mi, ii in enumerate(order_i):
            cols = tabs[ih * len(inc) + ii]
            g, ell, mue = col(cols, "g"), col(cols, "lensing"), col(cols, "cosine")
            al, be = col(cols, "alpha"), col(cols, "beta")
            jn = np.where(g > 0, ell * mue / np.where(g > 0, g, 1.0), 0.0) * limb_factor(limb, mue)
            for s, r in enumerate(rows):
                if not (r > _r_valid(a)) or r < r_src[0] or r > r_src[-1] * (1 + 1e-9):
                    continue
                k = int(np.clip(np.searchsorted(lr_src, np.log(r), side="right") - 1, 0, len(r_src) - 2))
                fr = (np.log(r) - lr_src[k]) / (lr_src[k + 1] - lr_src[k])
                fr = min(max(fr, 0.0), 1.0)
                lg_row = np.log(np.maximum(g[k], 1e-300)) * (1 - fr) + np.log(np.maximum(g[k + 1], 1e-300)) * fr
                jn_row = jn[k] * (1 - fr) + jn[k + 1] * fr
                al_row = al[k] * (1 - fr) + al[k + 1] * fr
                be_row = be[k] * (1 - fr) + be[k + 1] * fr
                phi0 = _front_azimuth(phi, al_row, be_row)
                want = (tcol + phi0) % (2 * np.pi)
                ph_ext = np.concatenate([phi % (2 * np.pi), [(phi[0] % (2 * np.pi)) + 2 * np.pi]])
                o = np.argsort(ph_ext[:-1])
                ph_s = np.concatenate([ph_ext[:-1][o], [ph_ext[:-1][o][0] + 2 * np.pi]])

                def per(v):
                    vs = np.concatenate([v[o], [v[o][0]]])
                    w = np.where(want < ph_s[0], want + 2 * np.pi, want)
                    return np.interp(w, ph_s, vs)

                data[ma, mi, 0, s] = np.exp(per(lg_row))
                data[ma, mi, 1, s] = np.maximum(per(jn_row), 0.0)
    return dict(spins=spins_all[order_a], thetas=np.deg2rad(inc[order_i]), NR=NR, NPHI=NPHI, r_in=r_in, data=data)
