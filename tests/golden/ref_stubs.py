"""In-memory stand-ins for `xspec` and `astropy.{units,constants}` so that the
UNMODIFIED reference (/root/reference/src/python_version/relagn.py) imports and
runs in a container that has neither HEASOFT nor astropy.

Used only by tests/golden/make_golden.py (fixture generation, build container
only -- /root/reference does not exist on the GPU box).

* astropy: CODATA-2018 constants as shipped by astropy >= 4.0
  and the handful of unit conversions relagn.py performs
  (relagn.py:287,339-342,513-527,708,929-930,970-971,1006,1179,1217-1218).
* xspec: `callModelFunction('kyconv', energies, params, flux)` mutates `flux`
  in place (relagn.py:1000-1003).  The stub forwards to a user-supplied
  callable (the CPU oracle's kyconv) and records every call.
"""
import sys
import types

import numpy as np

# ----------------------------------------------------------------------------
# constants (SI)
G = 6.6743e-11
M_sun = 1.988409870698051e+30
c = 299792458.0
h = 6.62607015e-34
k_B = 1.380649e-23
sigma_sb = 5.6703744191844314e-08
sigma_T = 6.6524587321e-29
m_p = 1.67262192369e-27
keV_J = 1.602176634e-16
Mpc_m = 3.0856775814913674e+22


class _Const:
    def __init__(self, value):
        self.value = value

    def __mul__(self, other):
        return _Const(self.value * (other.value if isinstance(other, _Const) else other))


# ----------------------------------------------------------------------------
# units: scale * kg^a m^b s^c
class Unit:
    __array_ufunc__ = None  # make ndarray * Unit defer to Unit.__rmul__

    def __init__(self, scale, dims):
        self.scale = float(scale)
        self.dims = tuple(dims)

    def __mul__(self, o):
        if isinstance(o, Unit):
            return Unit(self.scale * o.scale, tuple(a + b for a, b in zip(self.dims, o.dims)))
        return Quantity(o, self)

    def __rmul__(self, o):
        return Quantity(o, self)

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return Unit(self.scale / o.scale, tuple(a - b for a, b in zip(self.dims, o.dims)))
        raise TypeError

    def __rtruediv__(self, o):
        return Quantity(o, Unit(1.0 / self.scale, tuple(-a for a in self.dims)))


class _Equiv:
    def __init__(self, kind, nu=None):
        self.kind = kind
        self.nu = nu


def spectral():
    return _Equiv("spectral")


def spectral_density(nu):
    return _Equiv("spectral_density", nu)


_D_ENERGY = (1, 2, -2)
_D_FREQ = (0, 0, -1)
_D_LEN = (0, 1, 0)
_D_H = (1, 2, -1)  # J s


class Quantity:
    __array_ufunc__ = None

    def __init__(self, value, unit):
        self._v = np.asarray(value, dtype=np.float64) if np.ndim(value) else float(value)
        self.unit = unit

    @property
    def value(self):
        return self._v

    def __mul__(self, o):
        if isinstance(o, Unit):
            return Quantity(self._v, self.unit * o)
        return Quantity(self._v * o, self.unit)

    def __truediv__(self, o):
        if isinstance(o, Unit):
            return Quantity(self._v, self.unit / o)
        return Quantity(self._v / o, self.unit)

    def to(self, unit, equivalencies=None):
        f, t = self.unit, unit
        if f.dims == t.dims:
            return Quantity(self._v * (f.scale / t.scale), t)
        if equivalencies is None:
            raise ValueError("incompatible units")
        if equivalencies.kind == "spectral":
            si = self._v * f.scale
            if f.dims == _D_ENERGY and t.dims == _D_FREQ:
                return Quantity(si / h / t.scale, t)
            if f.dims == _D_ENERGY and t.dims == _D_LEN:
                return Quantity(h * c / si / t.scale, t)
            # spectral densities: the two units differ by a power of (J s)
            diff = tuple(a - b for a, b in zip(f.dims, t.dims))
            if diff == _D_H:  # e.g. W/Hz -> keV/s/keV : divide by h
                return Quantity(si / h / t.scale, t)
            if diff == tuple(-a for a in _D_H):  # e.g. W/keV -> W/Hz : multiply by h
                return Quantity(si * h / t.scale, t)
            raise ValueError("unsupported spectral conversion")
        if equivalencies.kind == "spectral_density":
            nu_q = equivalencies.nu
            nu = nu_q._v * nu_q.unit.scale
            si = self._v * f.scale
            diff = tuple(a - b for a, b in zip(f.dims, t.dims))
            if diff == (0, 1, 1):  # per Hz -> per length: L_lambda = L_nu nu^2 / c
                return Quantity(si * nu ** 2 / c / t.scale, t)
            raise ValueError("unsupported spectral_density conversion")
        raise ValueError


def install(kyconv_callable=None):
    """Insert the stub modules into sys.modules.  Returns the xspec stub (with
    `.calls`, a list of (energies, params) recorded at the kyconv seam)."""
    const = types.ModuleType("astropy.constants")
    const.G, const.M_sun, const.c, const.h = _Const(G), _Const(M_sun), _Const(c), _Const(h)
    const.sigma_T, const.sigma_sb, const.m_p, const.k_B = _Const(sigma_T), _Const(sigma_sb), _Const(m_p), _Const(k_B)
    u = types.ModuleType("astropy.units")
    u.m = Unit(1, (0, 1, 0))
    u.cm = Unit(1e-2, (0, 1, 0))
    u.AA = Unit(1e-10, (0, 1, 0))
    u.Mpc = Unit(Mpc_m, (0, 1, 0))
    u.s = Unit(1, (0, 0, 1))
    u.Hz = Unit(1, (0, 0, -1))
    u.J = Unit(1, _D_ENERGY)
    u.erg = Unit(1e-7, _D_ENERGY)
    u.keV = Unit(keV_J, _D_ENERGY)
    u.W = Unit(1, (1, 2, -3))
    u.spectral = spectral
    u.spectral_density = spectral_density
    astropy = types.ModuleType("astropy")
    astropy.units, astropy.constants = u, const
    xs = types.ModuleType("xspec")
    xs.calls = []
    xs.kyconv = kyconv_callable

    def callModelFunction(name, energies, params, flux, *a, **k):
        assert name == "kyconv"
        xs.calls.append((len(energies), list(params)))
        if xs.kyconv is None:
            raise RuntimeError("kyconv seam reached but no implementation installed")
        out = xs.kyconv(np.asarray(energies, dtype=np.float64), np.asarray(params, dtype=np.float64),
                        np.asarray(flux, dtype=np.float64))
        flux[:] = list(out)

    xs.callModelFunction = callModelFunction
    sys.modules["astropy"] = astropy
    sys.modules["astropy.units"] = u
    sys.modules["astropy.constants"] = const
    sys.modules["xspec"] = xs
    return xs
