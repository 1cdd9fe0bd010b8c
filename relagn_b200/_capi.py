"""ctypes binding of the C ABI declared in include/relagn_b200.h.

The shipped library is relagn_b200/_build/librelagn_b200.so (nvcc, sm_100a).
There is no CPU implementation behind this module: if the library is missing
or no CUDA device is usable the calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "librelagn_b200.so")

NPAR = 15
NATTR = 24
ATTR_NAMES = ["risco", "r_sg", "eta", "L_edd", "Mdot_edd", "Rg", "r_h", "r_w", "r_out", "hmax", "gamma_h", "kTe_h",
              "kTe_w", "gamma_w", "kT_seed", "Ldiss", "Lseed", "n_disc", "n_warm", "n_hot", "flags", "status", "mdot",
              "inc"]
A = {n: i for i, n in enumerate(ATTR_NAMES)}

MODEL_RELAGN, MODEL_RELQSO = 0, 1
FLAG_REL, FLAG_COMPONENTS, FLAG_FP32 = 1, 2, 4
S_OK, S_SPIN, S_INC, S_MDOT, S_CAP = 0, 1, 2, 3, 4
F_RW_LT_RH, F_RW_GT_ROUT, F_RH_LT_RISCO, F_RW_LT_RISCO, F_HMAX_GT_RH, F_HOT_SEED_GE_KTE = 1, 2, 4, 8, 16, 32

# every symbol include/relagn_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = ["rg_version", "rg_last_error", "rg_ctx_create", "rg_ctx_destroy", "rg_set_grid", "rg_tables_generate",
           "rg_tables_upload", "rg_tables_download", "rg_tables_info", "rg_kyconv", "rg_ky_kernel", "rg_eval_batch",
           "rg_eval_batch_pieces", "rg_eval_batch_host", "rg_launch_count", "rg_measure_fp64_peak", "rg_set_option", "rg_get_stat",
           "rg_reset_stats", "rg_convert_batch", "rg_set_instrument", "rg_photar_batch", "rg_set_response",
           "rg_set_data", "rg_fitstat_batch", "rg_eval_fitstat_batch", "rg_eval_fitstat_batch_host",
           "rg_xspec_eval", "rg_xspec_bind", "rg_tables_default", "rg_last_hot_shape", "rg_tables_set_limb"]
XSPEC_SYMBOLS = ["relagn_", "relqso_"]  # gfortran names of the XSPEC local models (seam B3)

UNITS = {"SI": 0, "cgs": 1, "cgs_wave": 2, "SI_wave": 3, "counts": 4}  # RG_UNIT_*
FIT_CHI2, FIT_CSTAT = 0, 1

OPT_PROFILE = 1
STATS = {"ms_setup": 1, "ms_nthcomp": 2, "ms_spectrum": 3, "n_setup": 4, "n_nthcomp": 5, "n_spectrum": 6, "sum_W": 7,
         "n_conv_ann": 8, "sum_jmax": 9, "n_sys": 10, "n_sets": 11, "ms_observe": 12, "n_observe": 13, "ms_hist": 14, "n_hist": 15}

PIECE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int)  # rg_piece_fn(user, first_set, n_sets)
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


class RelagnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("relagn_b200 error %d: %s" % (code, msg))
        self.code = code


def bind(lib):
    """Attach argtypes/restypes to a loaded CDLL."""
    lib.rg_version.restype = C.c_int
    lib.rg_version.argtypes = []
    lib.rg_last_error.restype = C.c_char_p
    lib.rg_last_error.argtypes = []
    lib.rg_ctx_create.restype = C.c_int
    lib.rg_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.rg_ctx_destroy.restype = None
    lib.rg_ctx_destroy.argtypes = [C.c_void_p]
    lib.rg_set_grid.restype = C.c_int
    lib.rg_set_grid.argtypes = [C.c_void_p, _dp, C.c_int]
    lib.rg_tables_generate.restype = C.c_int
    lib.rg_tables_generate.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_double,
                                       C.POINTER(C.c_long)]
    lib.rg_tables_upload.restype = C.c_int
    lib.rg_tables_upload.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _fp]
    lib.rg_tables_set_limb.restype = C.c_int
    lib.rg_tables_set_limb.argtypes = [C.c_void_p, C.c_int]
    lib.rg_tables_download.restype = C.c_int
    lib.rg_tables_download.argtypes = [C.c_void_p, _fp, C.c_size_t]
    lib.rg_tables_info.restype = C.c_int
    lib.rg_tables_info.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, _dp, C.POINTER(C.c_size_t)]
    lib.rg_kyconv.restype = C.c_int
    lib.rg_kyconv.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp]
    lib.rg_ky_kernel.restype = C.c_int
    lib.rg_ky_kernel.argtypes = [C.c_void_p] + [C.c_double] * 9 + [_dp, C.c_int, _ip, _ip]
    lib.rg_eval_batch.restype = C.c_int
    lib.rg_eval_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    lib.rg_eval_batch_pieces.restype = C.c_int
    lib.rg_eval_batch_pieces.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_uint, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int, PIECE_FN, C.c_void_p]
    lib.rg_eval_batch_host.restype = C.c_int
    lib.rg_eval_batch_host.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, C.c_double, C.c_uint, _dp, _dp]
    lib.rg_last_hot_shape.restype = C.c_int
    lib.rg_last_hot_shape.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.rg_launch_count.restype = C.c_long
    lib.rg_launch_count.argtypes = [C.c_void_p]
    lib.rg_measure_fp64_peak.restype = C.c_int
    lib.rg_measure_fp64_peak.argtypes = [C.c_void_p, _dp]
    lib.rg_set_option.restype = C.c_int
    lib.rg_set_option.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.rg_get_stat.restype = C.c_int
    lib.rg_get_stat.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.rg_reset_stats.restype = C.c_int
    lib.rg_reset_stats.argtypes = [C.c_void_p]
    vp = C.c_void_p
    lib.rg_convert_batch.restype = C.c_int
    lib.rg_convert_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.rg_set_instrument.restype = C.c_int
    lib.rg_set_instrument.argtypes = [vp, _dp, C.c_int]
    lib.rg_photar_batch.restype = C.c_int
    lib.rg_photar_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp]
    lib.rg_set_response.restype = C.c_int
    lib.rg_set_response.argtypes = [vp, C.c_int, _ip, _ip, _dp, C.c_double]
    lib.rg_set_data.restype = C.c_int
    lib.rg_set_data.argtypes = [vp, _dp, _dp, C.c_int]
    lib.rg_fitstat_batch.restype = C.c_int
    lib.rg_fitstat_batch.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    lib.rg_eval_fitstat_batch.restype = C.c_int
    lib.rg_eval_fitstat_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, C.c_uint, vp, vp, vp]
    lib.rg_eval_fitstat_batch_host.restype = C.c_int
    lib.rg_eval_fitstat_batch_host.argtypes = [vp, _dp, C.c_int, C.c_int, C.c_double, C.c_uint, _dp, _dp]
    lib.rg_xspec_eval.restype = C.c_int
    lib.rg_xspec_eval.argtypes = [C.c_int, C.c_uint, _dp, C.c_int, _dp, _dp]
    lib.rg_xspec_bind.restype = C.c_int
    lib.rg_xspec_bind.argtypes = [vp]
    lib.rg_tables_default.restype = C.c_int
    lib.rg_tables_default.argtypes = [vp, C.c_char_p]
    for name in XSPEC_SYMBOLS:
        f = getattr(lib, name)
        f.restype = None
        f.argtypes = [_fp, _ip, _fp, _ip, _fp, _fp]
    return lib


_lib = None


def load():
    """Load the sm_100a library.  Raises if it has not been built (python -m relagn_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RelagnError(-1, "CUDA library %s not built (run `python -m relagn_b200.build`); "
                                  "relagn_b200 has no CPU fallback" % LIB_PATH)
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


class Context:
    """One device context (rg_ctx).  `lib` is only overridden by the test-only host simulation."""

    def __init__(self, device=-1, lib=None):
        self.lib = lib if lib is not None else load()
        h = C.c_void_p()
        self._check(self.lib.rg_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.nE = 0
        self.ebins = None

    def _check(self, rc):
        if rc != 0:
            raise RelagnError(rc, self.lib.rg_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.rg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- grid ----
    def set_grid(self, ebins):
        eb = np.ascontiguousarray(ebins, dtype=np.float64)
        self._check(self.lib.rg_set_grid(self.h, _d(eb), eb.size - 1))
        self.nE = eb.size - 1
        self.ebins = eb.copy()

    # ---- tables ----
    def tables_upload(self, spins, thetas, NR, NPHI, r_in, data):
        sp = np.ascontiguousarray(spins, dtype=np.float64)
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        dat = np.ascontiguousarray(data, dtype=np.float32)
        assert dat.size == sp.size * th.size * 2 * NR * NPHI
        self._check(self.lib.rg_tables_upload(self.h, _d(sp), sp.size, _d(th), th.size, int(NR), int(NPHI), float(r_in),
                                              dat.ctypes.data_as(_fp)))

    def tables_generate(self, spins, thetas, NR, NPHI, r_in):
        sp = np.ascontiguousarray(spins, dtype=np.float64)
        th = np.ascontiguousarray(thetas, dtype=np.float64)
        bad = C.c_long(0)
        self._check(self.lib.rg_tables_generate(self.h, _d(sp), sp.size, _d(th), th.size, int(NR), int(NPHI),
                                                float(r_in), C.byref(bad)))
        return bad.value

    def tables_set_limb(self, law):
        """kyconv parameter 10: 0 isotropic, -1 Laor, -2 Haardt -- before tables_generate, or after tables_upload."""
        self._check(self.lib.rg_tables_set_limb(self.h, int(law)))

    def tables_info(self):
        ns, ni, nr, nphi = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        r_in = C.c_double()
        nf = C.c_size_t()
        self._check(self.lib.rg_tables_info(self.h, C.byref(ns), C.byref(ni), C.byref(nr), C.byref(nphi), C.byref(r_in),
                                            C.byref(nf)))
        return dict(n_spin=ns.value, n_inc=ni.value, NR=nr.value, NPHI=nphi.value, r_in=r_in.value, n_floats=nf.value)

    def tables_download(self):
        info = self.tables_info()
        out = np.zeros(info["n_floats"], dtype=np.float32)
        self._check(self.lib.rg_tables_download(self.h, out.ctypes.data_as(_fp), out.size))
        return out.reshape(info["n_spin"], info["n_inc"], 2, info["NR"], info["NPHI"])

    # ---- kyconv seam ----
    def kyconv(self, energies, par12, flux):
        """In place on `flux` when it is a C-contiguous float64 array; returns the convolved array."""
        eb = np.ascontiguousarray(energies, dtype=np.float64)
        par = np.ascontiguousarray(par12, dtype=np.float64)
        f = flux if (isinstance(flux, np.ndarray) and flux.dtype == np.float64 and flux.flags.c_contiguous) \
            else np.array(flux, dtype=np.float64)
        assert par.size == 12 and eb.size == f.size + 1
        self._check(self.lib.rg_kyconv(self.h, _d(eb), f.size, _d(par), _d(f)))
        return f

    def ky_kernel(self, a, theta_o, r1, r2, dlnE, alpha=0.0, beta=0.0, rb=1.0, zshift=0.0, cap=65536):
        H = np.zeros(cap)
        kmin, nk = C.c_int(), C.c_int()
        self._check(self.lib.rg_ky_kernel(self.h, a, theta_o, r1, r2, alpha, beta, rb, zshift, dlnE, _d(H), cap,
                                          C.byref(kmin), C.byref(nk)))
        return kmin.value, H[:nk.value].copy()

    # ---- batched evaluation ----
    def eval_batch_host(self, params, model=MODEL_RELAGN, rel=True, components=False, dr_dex=50.0, out=None,
                        attrs=True, fp32=False, attrs_out=None):
        """params: [P, 15] (relqso: first 7 columns used).  Returns (out [P, nE] or [P, 4, nE], attrs [P, 24]).
        out / attrs_out: caller-owned result buffers (e.g. pinned memory) to fill instead of fresh arrays."""
        p = np.asarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p[None, :]
        P = p.shape[0]
        if p.shape[1] != NPAR or not p.flags.c_contiguous:
            pp = np.zeros((P, NPAR))
            pp[:, :p.shape[1]] = p
            p = pp
        flags = (FLAG_REL if rel else 0) | (FLAG_COMPONENTS if components else 0) | (FLAG_FP32 if fp32 else 0)
        shape = (P, 4, self.nE) if components else (P, self.nE)
        if out is None:
            out = np.empty(shape)
        assert out.shape == shape and out.dtype == np.float64 and out.flags.c_contiguous
        at = None
        if attrs:
            at = attrs_out if attrs_out is not None else np.empty((P, NATTR))
            assert at.shape == (P, NATTR) and at.dtype == np.float64 and at.flags.c_contiguous
        self._check(self.lib.rg_eval_batch_host(self.h, _d(p), P, int(model), float(dr_dex), flags, _d(out),
                                                _d(at) if attrs else None))
        return out, at

    def eval_batch_device(self, params_ptr, P, out_ptr, attrs_ptr=0, model=MODEL_RELAGN, rel=True, components=False,
                          dr_dex=50.0, stream=0, fp32=False):
        """Raw device pointers (e.g. torch tensor .data_ptr()); asynchronous on `stream`."""
        flags = (FLAG_REL if rel else 0) | (FLAG_COMPONENTS if components else 0) | (FLAG_FP32 if fp32 else 0)
        self._check(self.lib.rg_eval_batch(self.h, C.c_void_p(params_ptr), int(P), int(model), float(dr_dex), flags,
                                           C.c_void_p(out_ptr), C.c_void_p(attrs_ptr) if attrs_ptr else None,
                                           C.c_void_p(stream) if stream else None))

    def eval_batch_device_pieces(self, params_ptr, P, out_ptr, piece_sets, on_piece, attrs_ptr=0, model=MODEL_RELAGN,
                                 rel=True, components=False, dr_dex=50.0, stream=0, fp32=False):
        """rg_eval_batch_pieces: on_piece(first_set, n_sets) is called on this thread as soon as the kernels that
        complete those rows of the output are enqueued on `stream`."""
        flags = (FLAG_REL if rel else 0) | (FLAG_COMPONENTS if components else 0) | (FLAG_FP32 if fp32 else 0)
        err = []

        def _cb(_user, first, n):
            try:
                on_piece(int(first), int(n))
            except BaseException as e:  # an exception must not unwind through the C frames
                err.append(e)

        cb = PIECE_FN(_cb)
        self._check(self.lib.rg_eval_batch_pieces(self.h, C.c_void_p(params_ptr), int(P), int(model), float(dr_dex),
                                                  flags, C.c_void_p(out_ptr),
                                                  C.c_void_p(attrs_ptr) if attrs_ptr else None,
                                                  C.c_void_p(stream) if stream else None, int(piece_sets), cb, None))
        if err:
            raise err[0]

    # ---- observation epilogue ----
    def convert_device(self, lnu_ptr, params_ptr, P, ncomp, unit, as_flux, out_ptr, model=MODEL_RELAGN, stream=0):
        """relagn.py:487-564 on device buffers [P][ncomp][nE]; out_ptr may equal lnu_ptr."""
        u = UNITS[unit] if isinstance(unit, str) else int(unit)
        self._check(self.lib.rg_convert_batch(self.h, C.c_void_p(lnu_ptr), C.c_void_p(params_ptr) if params_ptr else None,
                                              int(P), int(model), int(ncomp), u, int(bool(as_flux)),
                                              C.c_void_p(out_ptr), C.c_void_p(stream) if stream else None))

    def set_instrument(self, ear):
        e = np.ascontiguousarray(ear, dtype=np.float64)
        self._check(self.lib.rg_set_instrument(self.h, _d(e), e.size - 1))
        self.ne = e.size - 1
        self.n_chan = 0

    def set_response(self, rowptr, col, val, exposure=1.0):
        """CSR over channels: model_c = exposure * sum_k val[k] * photar[col[k]]."""
        rp = np.ascontiguousarray(rowptr, dtype=np.int32)
        cl = np.ascontiguousarray(col, dtype=np.int32)
        vl = np.ascontiguousarray(val, dtype=np.float64)
        assert rp[-1] == cl.size == vl.size
        self._check(self.lib.rg_set_response(self.h, rp.size - 1, rp.ctypes.data_as(_ip), cl.ctypes.data_as(_ip),
                                             _d(vl), float(exposure)))
        self.n_chan = rp.size - 1

    def set_data(self, counts, sigma=None, stat="chi2"):
        kind = {"chi2": FIT_CHI2, "cstat": FIT_CSTAT}[stat]
        cn = np.ascontiguousarray(counts, dtype=np.float64)
        assert cn.size == self.n_chan
        sg = np.ascontiguousarray(sigma, dtype=np.float64) if sigma is not None else None
        self._check(self.lib.rg_set_data(self.h, _d(cn), _d(sg) if sg is not None else None, kind))

    def photar_device(self, lnu_ptr, params_ptr, P, photar_ptr, model=MODEL_RELAGN, stream=0):
        self._check(self.lib.rg_photar_batch(self.h, C.c_void_p(lnu_ptr), C.c_void_p(params_ptr), int(P), int(model),
                                             C.c_void_p(photar_ptr), C.c_void_p(stream) if stream else None))

    def fitstat_device(self, lnu_ptr, params_ptr, P, stat_ptr, folded_ptr=0, model=MODEL_RELAGN, stream=0):
        self._check(self.lib.rg_fitstat_batch(self.h, C.c_void_p(lnu_ptr), C.c_void_p(params_ptr), int(P), int(model),
                                              C.c_void_p(stat_ptr), C.c_void_p(folded_ptr) if folded_ptr else None,
                                              C.c_void_p(stream) if stream else None))

    def eval_fitstat_device(self, params_ptr, P, stat_ptr, attrs_ptr=0, model=MODEL_RELAGN, rel=True, dr_dex=50.0,
                            stream=0, fp32=False):
        flags = (FLAG_REL if rel else 0) | (FLAG_FP32 if fp32 else 0)
        self._check(self.lib.rg_eval_fitstat_batch(self.h, C.c_void_p(params_ptr), int(P), int(model), float(dr_dex),
                                                   flags, C.c_void_p(stat_ptr),
                                                   C.c_void_p(attrs_ptr) if attrs_ptr else None,
                                                   C.c_void_p(stream) if stream else None))

    def eval_fitstat_host(self, params, model=MODEL_RELAGN, rel=True, dr_dex=50.0, attrs=False, fp32=False, out=None):
        """params [P, 15] on the host -> statistic [P] (and attrs [P, 24]); 120 B in, 8 B out per set."""
        p = np.asarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p[None, :]
        P = p.shape[0]
        if p.shape[1] != NPAR or not p.flags.c_contiguous:
            pp = np.zeros((P, NPAR))
            pp[:, :p.shape[1]] = p
            p = pp
        flags = (FLAG_REL if rel else 0) | (FLAG_FP32 if fp32 else 0)
        st = out if out is not None else np.empty(P)
        at = np.zeros((P, NATTR)) if attrs else None
        self._check(self.lib.rg_eval_fitstat_batch_host(self.h, _d(p), P, int(model), float(dr_dex), flags, _d(st),
                                                        _d(at) if attrs else None))
        return (st, at) if attrs else st

    # ---- XSPEC local-model ABI (seam B3) ----
    def xspec_bind(self):
        """Route relagn_ / relqso_ / rg_xspec_eval of this process through this context (and its tables)."""
        self._check(self.lib.rg_xspec_bind(self.h))

    def tables_default(self, npy_path=None):
        self._check(self.lib.rg_tables_default(self.h, npy_path.encode() if npy_path else None))

    def xspec_model(self, name, ear, param):
        """Call the XSPEC entry point `relagn_` / `relqso_` exactly as XSPEC does (real*4 arrays by reference)."""
        e = np.ascontiguousarray(ear, dtype=np.float32)
        p = np.ascontiguousarray(param, dtype=np.float32)
        ne = C.c_int(e.size - 1)
        ifl = C.c_int(1)
        photar = np.zeros(e.size - 1, dtype=np.float32)
        photer = np.zeros(e.size - 1, dtype=np.float32)
        getattr(self.lib, name + "_")(e.ctypes.data_as(_fp), C.byref(ne), p.ctypes.data_as(_fp), C.byref(ifl),
                                      photar.ctypes.data_as(_fp), photer.ctypes.data_as(_fp))
        return photar

    def xspec_eval(self, model, rel, ear, par):
        e = np.ascontiguousarray(ear, dtype=np.float64)
        p = np.zeros(NPAR)
        p[:len(par)] = par
        out = np.zeros(e.size - 1)
        self._check(self.lib.rg_xspec_eval(int(model), FLAG_REL if rel else 0, _d(e), e.size - 1, _d(p), _d(out)))
        return out

    def set_profile(self, on=True):
        self._check(self.lib.rg_set_option(self.h, OPT_PROFILE, 1.0 if on else 0.0))

    def reset_stats(self):
        self._check(self.lib.rg_reset_stats(self.h))

    def stats(self):
        out = {}
        v = C.c_double()
        for name, key in STATS.items():
            self._check(self.lib.rg_get_stat(self.h, key, C.byref(v)))
            out[name] = v.value
        return out

    def last_hot_shape(self, set_index=0):
        """Fnu_seed_hot (relagn.py:1200-1223) of a set of the last evaluated batch."""
        out = np.zeros(self.nE)
        self._check(self.lib.rg_last_hot_shape(self.h, int(set_index), _d(out)))
        return out

    def launch_count(self):
        return int(self.lib.rg_launch_count(self.h))

    def measure_fp64_peak(self):
        v = C.c_double()
        self._check(self.lib.rg_measure_fp64_peak(self.h, C.byref(v)))
        return v.value
