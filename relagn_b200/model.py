"""Host-side mirror of the reference's class API (scotthgn/RELAGN src/python_version/relagn.py):

    relagn(M, dist, log_mdot, a, cos_inc, kTe_hot, kTe_warm, gamma_hot, gamma_warm, r_hot, r_warm, log_rout,
           fcol, h_max, z)                                   relagn.py:232-247
    relqso(M, dist, log_mdot, a, cos_inc, fcol, z)           relagn.py:1767-1774

Same constructor parameters, class-level knobs (Emin, Emax, numE, dr_dex, units, as_flux), getters
(get_totSED / get_DiscComponent / get_WarmComponent / get_HotComponent(rel=...), get_Ledd, get_Rg, get_Mdot),
mutators (set_units, set_flux, set_lum, new_ear, change_rBins), validation behaviour (ValueError for spin,
inclination, relqso mdot; print-WARNING-and-clamp for radii and h_max) and physical attributes.

All spectrum arithmetic -- radial grid, Novikov-Thorne emissivity, local spectra, Kompaneets solve, relativistic
transfer -- runs in the CUDA library (one-set batch through the same rg_eval_batch path BatchEvaluator uses); this
file only does the O(nE) unit epilogue (relagn.py:487-564 stays on the host there too) and attribute bookkeeping.
There is no CPU fallback: constructing a model without a CUDA device raises.
"""
import numpy as np

from . import _capi
from . import tables as _tables
from .host import risco as _risco_host

# CODATA-2018 values as shipped by astropy >= 4 (the reference imports them from astropy.constants, relagn.py:33-39)
_H = 6.62607015e-34
_C = 299792458.0
_KEV = 1.602176634e-16
_KB = 1.380649e-23
_MPC_CM = 3.0856775814913674e+24

_ctx_cache = {}


def _context(device=None):
    """One library context per device, shared by all model objects of the process."""
    import torch
    if not torch.cuda.is_available():
        raise _capi.RelagnError(-1, "no CUDA device: relagn_b200 has no CPU fallback")
    dev = torch.cuda.current_device() if device is None else int(device)
    if dev not in _ctx_cache:
        _ctx_cache[dev] = {"ctx": _capi.Context(dev), "tables": False}
    return _ctx_cache[dev]


def set_tables(tables, device=None):
    """Use explicit Kerr transfer tables (dict spins, thetas, NR, NPHI, r_in, data) instead of the default set."""
    c = _context(device)
    c["ctx"].tables_upload(tables["spins"], tables["thetas"], tables["NR"], tables["NPHI"], tables["r_in"],
                           tables["data"])
    c["tables"] = True


def _make_rbins(logr_in, logr_out, dlog_r):
    """Edges of the reference's radial grid (relagn.py:731-776): built top-down by repeated subtraction, the
    innermost edge snapped to logr_in.  Host copy for the `logr_*_bins` attributes only; the kernels stream the
    same edges on the device (csrc/rg_common.cuh BinIter)."""
    edges = [logr_out]
    x = logr_out
    while x > logr_in:
        x = x - dlog_r
        edges.append(x)
    edges.reverse()
    if edges[0] != logr_in:
        if edges[0] < logr_in and len(edges) > 1:
            edges.pop(0)
        edges[0] = logr_in
    return np.array(edges)


class relagn:
    Emin = 1e-4
    Emax = 1e4
    numE = 2000
    mu = 0.55
    A = 0.3
    dr_dex = 50
    default_units = 'cgs'
    units = 'cgs'
    as_flux = False

    _model = _capi.MODEL_RELAGN

    def __init__(self, M=1e8, dist=100, log_mdot=-1, a=0, cos_inc=0.5, kTe_hot=100, kTe_warm=0.2, gamma_hot=1.7,
                 gamma_warm=2.7, r_hot=10, r_warm=20, log_rout=-1, fcol=1, h_max=10, z=0):
        self._params = [M, dist, log_mdot, a, cos_inc, kTe_hot, kTe_warm, gamma_hot, gamma_warm, r_hot, r_warm,
                        log_rout, fcol, h_max, z]
        self.M = M
        self.D, self.d = dist, dist * _MPC_CM
        self.mdot = 10 ** (log_mdot)
        self.a = np.float64(a)
        self.inc = np.arccos(cos_inc)
        self.cosinc = cos_inc
        self.kTe_h, self.kTe_w = kTe_hot, kTe_warm
        self.gamma_h, self.gamma_w = gamma_hot, gamma_warm
        self.fcol = fcol
        self.z = z
        self._check_spin()
        self._check_inc()
        self._init_grid(np.geomspace(self.Emin, self.Emax, self.numE))
        self._evaluate_setup()

    # ---- validation (relagn.py:372-386) ----
    def _check_spin(self):
        if not (self.a >= -0.998 and self.a <= 0.998):
            raise ValueError('Spin ' + str(self.a) + ' not physical! \n'
                             'Must be within: -0.998 <= a_star <= 0.998')

    def _check_inc(self):
        if not (self.cosinc <= 1 and self.cosinc >= 0.09):
            raise ValueError('Inclination out of bounds - will not work with kyconv! \n'
                             'Require: 0.09 <= cos(inc) <= 0.98 \n'
                             'Translates to: 11.5 <= inc <= 85 deg')

    # ---- grids ----
    def _init_grid(self, ebins):
        self.Ebins = np.asarray(ebins, dtype=np.float64)
        self.dEs = self.Ebins[1:] - self.Ebins[:-1]
        self.Egrid = self.Ebins[:-1] + 0.5 * self.dEs
        self.nu_grid = self.Egrid * _KEV / _H
        self.wave_grid = _C / self.nu_grid * 1e10
        self.nu_obs = self.nu_grid / (1 + self.z)
        self.E_obs = self.Egrid / (1 + self.z)
        self.wave_obs = self.wave_grid * (1 + self.z)

    def _run(self, rel):
        """One-set batch through the CUDA path; returns ([4, nE] W/Hz, attrs)."""
        c = _context()
        ctx = c["ctx"]
        if ctx.ebins is None or ctx.ebins.shape != self.Ebins.shape or not np.array_equal(ctx.ebins, self.Ebins):
            ctx.set_grid(self.Ebins)
        if rel and not c["tables"]:
            _tables.load_default(ctx)
            c["tables"] = True
        p = np.zeros((1, _capi.NPAR))
        p[0, :len(self._params)] = self._params
        out, at = ctx.eval_batch_host(p, model=self._model, rel=rel, components=True, dr_dex=float(self.dr_dex))
        # the hot corona's spectral shape (relagn.py:1200-1223), a public attribute of the reference's class
        self.Fnu_seed_hot = ctx.last_hot_shape(0)
        return out[0], at[0]

    def _evaluate_setup(self):
        """Constructor part: derived scalars from the device setup kernel (+ the non-relativistic spectra, which
        come for free with the same launch)."""
        spec, at = self._run(rel=False)
        A = _capi.A
        st = int(at[A["status"]])
        if st == _capi.S_MDOT:
            raise ValueError('mdot is out of bounds! \n'
                             'Require: -1.65 <= log mdot <= 0.5')
        if st != _capi.S_OK:
            raise _capi.RelagnError(st, "parameter set rejected by the device setup kernel (status %d)" % st)
        fl = int(at[A["flags"]])
        # the reference's print-and-clamp warnings (relagn.py:388-421)
        if fl & _capi.F_RW_LT_RH:
            print('WARNING r_warm < r_hot ---- Setting r_warm = r_hot')
        if fl & _capi.F_RW_GT_ROUT:
            print('WARNING r_warm > r_out ----- Setting r_warm = r_out')
        if fl & _capi.F_RH_LT_RISCO:
            print('WARNING r_hot < r_isco ----- Setting r_hot = r_isco')
        if fl & _capi.F_RW_LT_RISCO:
            print('WARNING! r_warm < r_isco ----- Settin r_warm = r_isco')
        if fl & _capi.F_HMAX_GT_RH:
            print('WARNING! hmax > r_h ------- Setting hmax = r_h')
        for name in ("risco", "r_sg", "eta", "L_edd", "Mdot_edd", "Rg", "r_h", "r_w", "r_out", "hmax", "gamma_h",
                     "kTe_h", "kTe_w", "gamma_w", "Ldiss", "Lseed"):
            setattr(self, name, float(at[A[name]]))
        self._kT_seed = float(at[A["kT_seed"]])
        self._n_ann = (int(at[A["n_disc"]]), int(at[A["n_warm"]]), int(at[A["n_hot"]]))
        self.dlog_r = 1 / self.dr_dex
        self.logr_ad_bins = _make_rbins(np.log10(self.r_w), np.log10(self.r_out), self.dlog_r)
        self.logr_wc_bins = _make_rbins(np.log10(self.r_h), np.log10(self.r_w), self.dlog_r)
        self.logr_hc_bins = _make_rbins(np.log10(self.risco), np.log10(self.r_h), self.dlog_r)
        self._cache = {False: spec}
        self._store(False, spec)

    def _store(self, rel, spec):
        sfx = 'rel' if rel else 'norel'
        self.__dict__['Lnu_disc_' + sfx] = spec[0]
        self.__dict__['Lnu_warm_' + sfx] = spec[1]
        self.__dict__['Lnu_hot_' + sfx] = spec[2]

    def _spectra(self, rel):
        rel = bool(rel)
        if rel not in self._cache:
            spec, _ = self._run(rel)
            self._cache[rel] = spec
            self._store(rel, spec)
        return self._cache[rel]

    # ---- units (relagn.py:437-564) ----
    def set_units(self, new_unit='cgs'):
        unit_lst = ['cgs', 'cgs_wave', 'SI', 'SI_wave', 'counts']
        if new_unit not in unit_lst:
            print('Invalid Unit!!!')
            print(f'Valid options are: {unit_lst}')
            print('Setting as default: cgs')
            new_unit = 'cgs'
        self.units = new_unit

    def set_flux(self):
        self.as_flux = True

    def set_lum(self):
        self.as_flux = False

    def _to_newUnit(self, L, as_spec=True):
        if as_spec:
            if self.units == 'cgs':
                return L * 1e7
            if self.units == 'counts':
                # W/Hz -> keV/s/keV is a division by h (J s); then photons: / E
                Lnew = L / _H
                return Lnew / self.Egrid if np.ndim(L) == 1 else Lnew / self.Egrid[:, np.newaxis]
            if self.units == 'cgs_wave':
                # W/Hz -> erg/s/Angstrom: L_lambda = L_nu nu^2 / c
                return L * self.nu_grid ** 2 / _C * 1e-10 * 1e7
            if self.units == 'SI_wave':
                return L * self.nu_grid ** 2 / _C * 1e-10
            return L
        if self.units in ('cgs', 'counts', 'cgs_wave'):
            return L * 1e7
        return L

    def _to_flux(self, L):
        d = self.d if self.units in ('cgs', 'counts', 'cgs_wave') else self.d / 100
        return L / (4 * np.pi * d ** 2 * (1 + self.z))

    def new_ear(self, new_es):
        """User energy grid of bin EDGES in keV (relagn.py:780-805)."""
        self._init_grid(np.asarray(new_es, dtype=np.float64))
        self.Emin = min(self.Egrid)
        self.Emax = max(self.Egrid)
        self.numE = len(self.Egrid)
        self._evaluate_setup()

    def change_rBins(self, new_drdex):
        """Radial bins per decade (relagn.py:809-831)."""
        self.dr_dex = new_drdex
        self._evaluate_setup()

    # ---- scalar helpers kept for API compatibility (host, O(1)) ----
    def calc_Tnt(self, r):
        """Novikov-Thorne T^4 in K^4 at radius r [Rg] (relagn.py:662-685).  Host evaluation for interactive use;
        the spectrum path evaluates it on the device (csrc/rg_common.cuh rg_tnt4)."""
        r = np.asarray(r, dtype=np.float64)
        a = float(self.a)
        y, yi = np.sqrt(r), np.sqrt(self.risco)
        ac = np.arccos(a)
        y1 = 2 * np.cos(ac / 3 - np.pi / 3)
        y2 = 2 * np.cos(ac / 3 + np.pi / 3)
        y3 = -2 * np.cos(ac / 3)
        B = 1 - 3 / r + 2 * a / r ** 1.5
        C1 = 1 - yi / y - (3 * a / (2 * y)) * np.log(y / yi)
        C2 = 0.0
        for yk, yo1, yo2 in ((y1, y2, y3), (y2, y1, y3), (y3, y1, y2)):
            C2 = C2 + (3 * (yk - a) ** 2 / (y * yk * (yk - yo1) * (yk - yo2))) * np.log((y - yk) / (yi - yk))
        G_Msun = 6.6743e-11 * 1.988409870698051e+30
        sigma = 5.6703744191844314e-08
        fac = 3 * G_Msun * self.M * self.mdot * self.Mdot_edd / (8 * np.pi * sigma * (r * self.Rg) ** 3)
        return fac * (C1 - C2) / B

    def calc_fcol(self, Tm):
        """Done et al. (2012) colour correction (relagn.py:689-719)."""
        if Tm > 1e5:
            return (72 / (_KB * Tm / _KEV)) ** (1 / 9)
        if 3e4 < Tm < 1e5:
            return (Tm / 3e4) ** 0.82
        return 1.0

    # ---- getters (relagn.py:1412-1642) ----
    def _out(self, L):
        L = self._to_newUnit(L, as_spec=True)
        if self.as_flux:
            L = self._to_flux(L)
        return L

    def get_DiscComponent(self, rel=True):
        return self._out(self._spectra(rel)[0].copy())

    def get_WarmComponent(self, rel=True):
        return self._out(self._spectra(rel)[1].copy())

    def get_HotComponent(self, rel=True):
        return self._out(self._spectra(rel)[2].copy())

    def get_totSED(self, rel=True):
        return self.get_DiscComponent(rel=rel) + self.get_WarmComponent(rel=rel) + self.get_HotComponent(rel=rel)

    def get_Ledd(self):
        return self._to_newUnit(self.L_edd, as_spec=False)

    def get_Rg(self):
        # the reference raises AttributeError here outside SI units (it reads self.R_G, relagn.py:1618);
        # we return the documented value (cm) instead
        return self.Rg if self.units in ('SI', 'SI_wave') else self.Rg * 100

    def get_Mdot(self):
        Mdot = self.mdot * self.Mdot_edd
        return Mdot if self.units in ('SI', 'SI_wave') else Mdot * 1e3


class relqso(relagn):
    _model = _capi.MODEL_RELQSO

    def __init__(self, M=1e5, dist=1, log_mdot=-1, a=0, cos_inc=0.5, fcol=1, z=0):
        self._params = [M, dist, log_mdot, a, cos_inc, fcol, z]
        self.M = M
        self.D, self.d = dist, dist * _MPC_CM
        self.mdot = 10 ** (log_mdot)
        self.a = np.float64(a)
        self.inc = np.arccos(cos_inc)
        self.cosinc = cos_inc
        self.fcol = fcol
        self.z = z
        self._check_spin()
        self._check_inc()
        self._check_mdot()
        self._init_grid(np.geomspace(self.Emin, self.Emax, self.numE))
        self._evaluate_setup()

    def _check_mdot(self):
        lmdot = np.log10(self.mdot)
        if not (lmdot >= -1.65 and lmdot <= 0.5):
            raise ValueError('mdot is out of bounds! \n'
                             'Require: -1.65 <= log mdot <= 0.5')
