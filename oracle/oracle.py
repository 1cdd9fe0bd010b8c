"""ctypes front-end of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(relagn_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

NPAR = 15
NATTR = 24
ATTR_NAMES = ["risco", "r_sg", "eta", "L_edd", "Mdot_edd", "Rg", "r_h", "r_w", "r_out", "hmax", "gamma_h",
              "kTe_h", "kTe_w", "gamma_w", "kT_seed", "Ldiss", "Lseed", "n_disc", "n_warm", "n_hot", "flags",
              "status", "mdot", "inc"]

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle with gcc (idempotent)."""
    srcs = [os.path.join(_HERE, f) for f in ("relagn_oracle.c", "kerr_oracle.c", "observe_oracle.c",
                                             "relagn_oracle.h", "par_for.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    L.ro_risco.restype = C.c_double
    L.ro_risco.argtypes = [C.c_double]
    L.ro_nt_factor.restype = C.c_double
    L.ro_nt_factor.argtypes = [C.c_double] * 3
    L.ro_fcol.restype = C.c_double
    L.ro_fcol.argtypes = [C.c_double]
    L.ro_make_rbins.restype = C.c_int
    L.ro_make_rbins.argtypes = [C.c_double, C.c_double, C.c_double, _dp, C.c_int]
    L.ro_donthcomp.restype = None
    L.ro_donthcomp.argtypes = [_dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp]
    L.ro_thcompton.restype = None
    L.ro_thcompton.argtypes = [C.c_double, C.c_double, C.c_double, _dp, C.POINTER(C.c_int), _dp]
    L.ro_eval.restype = C.c_int
    L.ro_eval.argtypes = [_dp, C.c_int, C.c_double, _dp, C.c_int, C.c_int, C.c_void_p, _dp, _dp]
    L.ro_eval_batch.restype = None
    L.ro_eval_batch.argtypes = [_dp, C.c_int, C.c_int, C.c_double, _dp, C.c_int, C.c_int, C.c_void_p, _dp, _dp, _ip,
                                C.c_int]
    L.ro_hot_shape.restype = C.c_int
    L.ro_hot_shape.argtypes = [_dp, C.c_int, C.c_double, _dp, C.c_int, _dp]
    L.ro_setup.restype = C.c_int
    L.ro_setup.argtypes = [_dp, C.c_int, C.c_double, _dp]
    L.ro_trace_ray.restype = C.c_int
    L.ro_trace_ray.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_double)] * 2
    L.ro_make_node.restype = C.c_int
    L.ro_make_node.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, _fp, _fp, C.c_int]
    L.ro_row_radius.restype = C.c_double
    L.ro_row_radius.argtypes = [C.c_double, C.c_int, C.c_int]
    L.ro_r_valid.restype = C.c_double
    L.ro_r_valid.argtypes = [C.c_double]
    L.ro_set_limb.restype = None
    L.ro_set_limb.argtypes = [C.c_int]
    L.ro_node_extras.restype = None
    L.ro_node_extras.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.ro_make_node_rows.restype = C.c_int
    L.ro_make_node_rows.argtypes = [C.c_double, C.c_double, C.c_int, _dp, C.c_int, _fp, _fp, C.c_int]
    L.ro_tables_finish.restype = None
    L.ro_tables_finish.argtypes = [C.c_void_p]
    L.ro_slice.restype = C.c_int
    L.ro_slice.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, _dp, C.POINTER(C.c_int)]
    L.ro_tables_create.restype = C.c_void_p
    L.ro_tables_create.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_double]
    L.ro_tables_generate.restype = C.c_void_p
    L.ro_tables_generate.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    L.ro_tables_free.restype = None
    L.ro_tables_free.argtypes = [C.c_void_p]
    L.ro_tables_data.restype = C.POINTER(C.c_float)
    L.ro_tables_data.argtypes = [C.c_void_p]
    L.ro_tables_nfloats.restype = C.c_long
    L.ro_tables_nfloats.argtypes = [C.c_void_p]
    L.ro_ky_kernel.restype = C.c_int
    L.ro_ky_kernel.argtypes = [C.c_void_p] + [C.c_double] * 9 + [C.c_int, _dp, C.c_int, C.POINTER(C.c_int),
                                                                  C.POINTER(C.c_int)]
    L.ro_kyconv.restype = C.c_int
    L.ro_kyconv.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp]
    L.ro_constants.restype = None
    L.ro_constants.argtypes = [_dp]
    L.ro_convert.restype = None
    L.ro_convert.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp]
    L.ro_photar.restype = None
    L.ro_photar.argtypes = [_dp, C.c_int, _dp, C.c_double, C.c_double, _dp, C.c_int, _dp]
    L.ro_fold_stat.restype = C.c_double
    L.ro_fold_stat.argtypes = [_dp, C.c_int, _ip, _ip, _dp, C.c_double, _dp, C.c_void_p, C.c_int, C.c_void_p]
    _lib = L
    return L


def constants():
    o = np.zeros(8)
    lib().ro_constants(o)
    return dict(zip(["GMsun", "c", "h", "k_B", "sigma_sb", "keV", "Mpc_cm", "pi"], o))


def make_rbins(logr_in, logr_out, dlr):
    cap = int(abs(logr_out - logr_in) / dlr) + 16
    e = np.zeros(cap)
    n = lib().ro_make_rbins(float(logr_in), float(logr_out), float(dlr), e, cap)
    assert n > 0
    return e[:n].copy()


def donthcomp(ear, gamma, kte, ktbb, z=0.0):
    ear = np.ascontiguousarray(ear, dtype=np.float64)
    out = np.zeros_like(ear)
    lib().ro_donthcomp(ear, ear.size, gamma, kte, ktbb, z, out)
    return out


def thcompton(tempbb, theta, gamma):
    x = np.zeros(900)
    s = np.zeros(900)
    j = C.c_int(0)
    lib().ro_thcompton(tempbb, theta, gamma, x, C.byref(j), s)
    return x, j.value, s


class Tables:
    """Kerr transfer tables (CPU-generated or wrapped around a given array)."""

    def __init__(self, spins, thetas, NR, NPHI, r_in, data=None, threads=0):
        self.spins = np.ascontiguousarray(spins, dtype=np.float64)
        self.thetas = np.ascontiguousarray(thetas, dtype=np.float64)
        self.NR, self.NPHI, self.r_in = int(NR), int(NPHI), float(r_in)
        L = lib()
        if data is None:
            self.ptr = L.ro_tables_generate(self.spins, self.spins.size, self.thetas, self.thetas.size, self.NR,
                                            self.NPHI, self.r_in, threads)
        else:
            self.ptr = L.ro_tables_create(self.spins, self.spins.size, self.thetas, self.thetas.size, self.NR,
                                          self.NPHI, self.r_in)
            self.data[...] = np.asarray(data, dtype=np.float32).reshape(self.data.shape)
            L.ro_tables_finish(self.ptr)

    def rows(self):
        return np.array([lib().ro_row_radius(self.r_in, self.NR, s) for s in range(self.NR)])

    def slice(self, a, theta_o):
        """(ln g [NR][NPHI], Jn [NR][NPHI], first row) of the blended slice at (a, theta_o)."""
        lg = np.zeros((self.NR, self.NPHI))
        jn = np.zeros((self.NR, self.NPHI))
        s_lo = C.c_int(0)
        st = lib().ro_slice(self.ptr, float(a), float(theta_o), lg, jn, C.byref(s_lo))
        if st:
            raise RuntimeError("ro_slice status %d" % st)
        return lg, jn, s_lo.value

    @property
    def data(self):
        L = lib()
        n = L.ro_tables_nfloats(self.ptr)
        arr = np.ctypeslib.as_array(L.ro_tables_data(self.ptr), shape=(n,))
        return arr.reshape(self.spins.size, self.thetas.size, 2, self.NR, self.NPHI)

    def __del__(self):
        try:
            lib().ro_tables_free(self.ptr)
        except Exception:
            pass


def trace_ray(a, theta_o, alpha, beta):
    r = C.c_double(0)
    p = C.c_double(0)
    ok = lib().ro_trace_ray(a, theta_o, alpha, beta, C.byref(r), C.byref(p))
    return ok, r.value, p.value


def make_node(a, theta_o, NR, NPHI, r_in, threads=0):
    g = np.zeros((NR, NPHI), dtype=np.float32)
    jn = np.zeros((NR, NPHI), dtype=np.float32)
    nfail = lib().ro_make_node(a, theta_o, NR, NPHI, r_in, g, jn, threads)
    return g, jn, nfail


def make_node_full(a, theta_o, rows, NPHI, limb=0, absolute_phi=False):
    """One node on explicit row radii with the image-plane coordinates and the emission cosine:
    dict(g, jn, alpha, beta, mue, unresolved).  Runs single-threaded (the extras are per thread)."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    NR = rows.size
    g = np.zeros((NR, NPHI), dtype=np.float32)
    jn = np.zeros((NR, NPHI), dtype=np.float32)
    al, be, mu = np.zeros((NR, NPHI)), np.zeros((NR, NPHI)), np.zeros((NR, NPHI))
    L = lib()
    L.ro_set_limb(int(limb))
    L.ro_node_extras(al.ctypes.data, be.ctypes.data, mu.ctypes.data, int(bool(absolute_phi)))
    try:
        bad = L.ro_make_node_rows(float(a), float(theta_o), NR, rows, int(NPHI), g, jn, 1)
    finally:
        L.ro_node_extras(None, None, None, 0)
        L.ro_set_limb(0)
    return dict(g=g, jn=jn, alpha=al, beta=be, mue=mu, unresolved=bad)


def ky_kernel(tab, a, theta_o, r1, r2, dlnE, alpha=0.0, beta=0.0, rb=1.0, zshift=0.0, nsub=0, cap=65536):
    H = np.zeros(cap)
    kmin = C.c_int(0)
    nk = C.c_int(0)
    st = lib().ro_ky_kernel(tab.ptr, a, theta_o, r1, r2, alpha, beta, rb, zshift, dlnE, nsub, H, cap, C.byref(kmin),
                            C.byref(nk))
    if st:
        raise RuntimeError("ro_ky_kernel status %d" % st)
    return kmin.value, H[:nk.value].copy()


def kyconv(tab, ebins, par12, flux):
    """In-place XSPEC-kyconv seam (returns the mutated copy)."""
    ebins = np.ascontiguousarray(ebins, dtype=np.float64)
    f = np.ascontiguousarray(flux, dtype=np.float64).copy()
    par = np.ascontiguousarray(par12, dtype=np.float64)
    st = lib().ro_kyconv(tab.ptr, ebins, f.size, par, f)
    if st:
        raise RuntimeError("ro_kyconv status %d" % st)
    return f


def setup(params, model=0, dr_dex=50.0):
    p = np.zeros(NPAR)
    p[:len(params)] = params
    at = np.zeros(NATTR)
    st = lib().ro_setup(p, model, dr_dex, at)
    return st, dict(zip(ATTR_NAMES, at))


def hot_shape(params, ebins, model=0, dr_dex=50.0):
    """Fnu_seed_hot (relagn.py:1200-1223) of one parameter set on the grid."""
    p = np.zeros(NPAR)
    p[:len(params)] = params
    ebins = np.ascontiguousarray(ebins, dtype=np.float64)
    F = np.zeros(ebins.size - 1)
    lib().ro_hot_shape(p, model, dr_dex, ebins, ebins.size - 1, F)
    return F


def evaluate(params, ebins, model=0, rel=False, tab=None, dr_dex=50.0):
    """One parameter set -> (status, out[4, nE] (disc, warm, hot, total; W/Hz), attrs dict)."""
    p = np.zeros(NPAR)
    p[:len(params)] = params
    ebins = np.ascontiguousarray(ebins, dtype=np.float64)
    nE = ebins.size - 1
    out = np.zeros((4, nE))
    at = np.zeros(NATTR)
    st = lib().ro_eval(p, model, dr_dex, ebins, nE, int(bool(rel)), tab.ptr if tab is not None else None, out, at)
    return st, out, dict(zip(ATTR_NAMES, at))


def evaluate_batch(params, ebins, model=0, rel=False, tab=None, dr_dex=50.0, threads=0):
    params = np.asarray(params, dtype=np.float64)
    P = params.shape[0]
    pp = np.zeros((P, NPAR))
    pp[:, :params.shape[1]] = params
    ebins = np.ascontiguousarray(ebins, dtype=np.float64)
    nE = ebins.size - 1
    out = np.zeros((P, 4, nE))
    at = np.zeros((P, NATTR))
    st = np.zeros(P, dtype=np.int32)
    lib().ro_eval_batch(pp, P, model, dr_dex, ebins, nE, int(bool(rel)), tab.ptr if tab is not None else None, out,
                        at, st, threads)
    return st, out, at


# ---- observation epilogue (observe_oracle.c) ----
UNITS = {"SI": 0, "cgs": 1, "cgs_wave": 2, "SI_wave": 3, "counts": 4}


def convert(L, ebins, unit, as_flux, dist_mpc, z):
    """relagn.py:487-564 on one spectrum [nE] (W/Hz)."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    eb = np.ascontiguousarray(ebins, dtype=np.float64)
    out = np.zeros_like(L)
    lib().ro_convert(L, L.size, eb, UNITS[unit] if isinstance(unit, str) else int(unit), int(bool(as_flux)),
                     float(dist_mpc), float(z), out)
    return out


def photar(L, ebins, dist_mpc, z, ear):
    """relagn.f:98-112: photons/cm^2/s per bin of the observer's grid `ear` (edges, keV)."""
    L = np.ascontiguousarray(L, dtype=np.float64)
    eb = np.ascontiguousarray(ebins, dtype=np.float64)
    ear = np.ascontiguousarray(ear, dtype=np.float64)
    out = np.zeros(ear.size - 1)
    lib().ro_photar(L, L.size, eb, float(dist_mpc), float(z), ear, ear.size - 1, out)
    return out


def fold_stat(photar_, rowptr, col, val, exposure, counts, sigma=None, kind=0):
    """(statistic, folded model counts [n_chan]); kind 0 chi^2 (needs sigma), 1 Cash."""
    ph = np.ascontiguousarray(photar_, dtype=np.float64)
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    cl = np.ascontiguousarray(col, dtype=np.int32)
    vl = np.ascontiguousarray(val, dtype=np.float64)
    cn = np.ascontiguousarray(counts, dtype=np.float64)
    n_chan = rp.size - 1
    folded = np.zeros(n_chan)
    sg = np.ascontiguousarray(sigma, dtype=np.float64) if sigma is not None else None
    st = lib().ro_fold_stat(ph, n_chan, rp, cl, vl, float(exposure), cn, sg.ctypes.data if sg is not None else None,
                            int(kind), folded.ctypes.data)
    return st, folded
